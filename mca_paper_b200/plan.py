"""Static (construction-time) tables of the MCA/MMA model, built on the host with numpy.

Reference: MCA.__init__ model.py:312-329 (return token types), :355/:383-390 (token_types), :392-398 (MMA "zorro"
mask), :408-430 (MCA fusion-channel mask), :400-406/:432-446 (pooling masks), model.py:11-12 (combination order);
MCAPretrainingLoss.__init__/forward model.py:151-168,198-220 (pair list, names, presence-mask rules).

Beyond the dense boolean masks the reference registers as buffers (kept for state_dict compatibility), this module
derives what the kernels consume: a per-key group id and a per-query-row bitmask of allowed key groups (the masks
are exactly block-structured — SURVEY.md Appendix B — which is verified bit-for-bit here), a 128-row tile schedule
that lists, for every query tile, only the key tiles containing at least one allowed pair, its transpose for the
backward pass, and the contrastive-pair plan.
"""
from __future__ import annotations

from itertools import chain, combinations
from typing import Dict, List

import numpy as np

FUSION_TOKEN = -1
GLOBAL_TOKEN = -2
TILE = 128


def fusion_channel_sets(n_modalities: int, cardinalities) -> List[frozenset]:
    return [frozenset(c) for c in chain.from_iterable(combinations(range(n_modalities), k) for k in cardinalities)]


class StaticPlan:
    def __init__(self, encoder_configs: dict, num_fusion_tokens: int, fusion_combos, fcl: bool, zorro: bool,
                 no_fusion: bool, bimodal_contrastive: bool, non_fusion_fcl: bool):
        self.names = list(encoder_configs.keys())
        self.n_mod = len(self.names)
        if self.n_mod > 8:
            raise AssertionError("at most 8 modalities are supported")
        self.lengths = [int(encoder_configs[k]["max_tokens"]) for k in self.names]
        self.offsets = [int(x) for x in np.cumsum([0] + self.lengths[:-1])]
        self.eao = False
        self.mask_src = list(range(self.n_mod))      # (virtual) modality blocks of the packed sequence: which
        self.mask_lengths = list(self.lengths)        # modality's attention_mask pads them, and their lengths
        self.zorro, self.fcl, self.no_fusion = bool(zorro), bool(fcl), bool(no_fusion)
        self.do_fcl = self.fcl and not self.zorro
        self.F = 0 if no_fusion else int(num_fusion_tokens)
        self.n_tok = int(sum(self.lengths))
        self.N = self.n_tok + self.F
        self.combos = fusion_channel_sets(self.n_mod, fusion_combos)

        # ---- pooled ("return") token types
        if no_fusion:
            rtt = list(range(self.n_mod)) + [GLOBAL_TOKEN]
        elif (not fcl) or zorro:
            rtt = list(range(self.n_mod)) + [FUSION_TOKEN, GLOBAL_TOKEN]
        else:
            rtt = list(range(self.n_mod)) + [FUSION_TOKEN] * len(self.combos) + [GLOBAL_TOKEN]
        self.return_token_types = rtt
        self.R = len(rtt)

        # ---- token types and key groups
        tt = np.concatenate([np.full(n, i, dtype=np.int64) for i, n in enumerate(self.lengths)]
                            + [np.full(self.F, FUSION_TOKEN, dtype=np.int64)])
        self.token_types = tt
        keygrp = np.zeros(self.N, dtype=np.uint8)
        for i, (o, n) in enumerate(zip(self.offsets, self.lengths)):
            keygrp[o:o + n] = i
        if self.F:
            if zorro:
                keygrp[self.n_tok:] = self.n_mod
                self.n_groups = self.n_mod + 1
            else:
                if self.F % len(self.combos) != 0:
                    raise AssertionError(
                        f"Number of fusion tokens {self.F} must be divisible by the number of combinations {len(self.combos)}")
                self.nsub = self.F // len(self.combos)
                keygrp[self.n_tok:] = self.n_mod + np.arange(self.F) // self.nsub
                self.n_groups = self.n_mod + len(self.combos)
        else:
            self.n_groups = self.n_mod
        if self.n_groups > 32:
            raise AssertionError("more than 32 key groups")
        self.keygrp = keygrp

        # ---- per-row allowed-group bitmasks (self-attention and pooling)
        rowbits = np.zeros(self.N, dtype=np.uint32)
        for i, (o, n) in enumerate(zip(self.offsets, self.lengths)):
            rowbits[o:o + n] = np.uint32(1 << i)  # a modality token sees its own modality only
        if self.F:
            if zorro:
                rowbits[self.n_tok:] = np.uint32((1 << self.n_groups) - 1)  # fusion sees everything
            else:
                for c, combo in enumerate(self.combos):
                    bits = sum(1 << m for m in combo) | (1 << (self.n_mod + c))
                    rowbits[self.n_tok + c * self.nsub:self.n_tok + (c + 1) * self.nsub] = np.uint32(bits)
        self.rowbits = rowbits
        pool_bits = np.zeros(self.R, dtype=np.uint32)
        fusion_groups = sum(1 << g for g in range(self.n_mod, self.n_groups))
        f_seen = 0
        for r, t in enumerate(rtt):
            if t >= 0:
                pool_bits[r] = 1 << t
            elif t == GLOBAL_TOKEN:
                pool_bits[r] = (1 << self.n_groups) - 1
            else:  # fusion row
                if self.do_fcl:
                    pool_bits[r] = 1 << (self.n_mod + f_seen)
                    f_seen += 1
                else:
                    pool_bits[r] = fusion_groups
        self.pool_rowbits = pool_bits

        # ---- dense masks (True = may not attend), as the reference registers them
        self.attn_mask = ~(((rowbits[:, None] >> keygrp[None, :].astype(np.uint32)) & 1).astype(bool))
        self.pool_mask = ~(((pool_bits[:, None] >> keygrp[None, :].astype(np.uint32)) & 1).astype(bool))

        self._build_tiles()
        self._build_loss_plan(bimodal_contrastive, non_fusion_fcl)

    # ------------------------------------------------------------------------------------------------ tiles
    def _segments(self):
        segs = [(o, n) for o, n in zip(self.offsets, self.lengths)]
        if self.F:
            segs.append((self.n_tok, self.F))
        return segs

    def _build_tiles(self):
        segs = self._segments()
        tiles = []
        for o, n in segs:
            for s in range(o, o + n, TILE):
                tiles.append((s, min(TILE, o + n - s)))
        self.tiles = np.asarray(tiles, dtype=np.int32).reshape(-1, 2)
        allowed = ~self.attn_mask
        fwd_refs, q_rows = [], []
        pair = {}
        for qi, (qs, ql) in enumerate(tiles):
            off = len(fwd_refs)
            for ki, (ks, kl) in enumerate(tiles):
                blk = allowed[qs:qs + ql, ks:ks + kl]
                if blk.any():
                    flag = 0 if blk.all() else 1
                    fwd_refs.append((ki, flag))
                    pair[(qi, ki)] = flag
            q_rows.append((qs, ql, off, len(fwd_refs) - off))
        self.q_tiles = np.asarray(q_rows, dtype=np.int32).reshape(-1, 4)
        # launch order of the forward kernel: heaviest query tiles first (longest-processing-time-first balancing)
        order = np.argsort(-self.q_tiles[:, 3], kind="stable")
        self.q_tiles_sorted = np.ascontiguousarray(self.q_tiles[order])
        # key group shared by every key of a tile, 255 when the tile mixes groups (the fusion sub-blocks)
        tg = []
        for ks, kl in tiles:
            g = np.unique(self.keygrp[ks:ks + kl])
            tg.append(int(g[0]) if len(g) == 1 else 255)
        self.tile_grp = np.asarray(tg, dtype=np.uint8)
        self.kt_list = np.asarray(fwd_refs, dtype=np.int32).reshape(-1, 2)
        bwd_refs, k_rows = [], []
        for ki, (ks, kl) in enumerate(tiles):
            off = len(bwd_refs)
            for qi in range(len(tiles)):
                if (qi, ki) in pair:
                    bwd_refs.append((qi, pair[(qi, ki)]))
            k_rows.append((ks, kl, off, len(bwd_refs) - off))
        self.k_tiles_q = np.asarray(k_rows, dtype=np.int32).reshape(-1, 4)
        self.qt_list = np.asarray(bwd_refs, dtype=np.int32).reshape(-1, 2)
        self.n_tile_pairs = len(fwd_refs)
        self.allowed_pairs = int(allowed.sum())

    # ------------------------------------------------------------------------------------------------ losses
    def _build_loss_plan(self, bimodal: bool, non_fusion_fcl: bool):
        names, n = self.names, self.n_mod
        row: Dict[object, int] = {m: i for i, m in enumerate(names)}
        self.output_rows = [(m, i) for i, m in enumerate(names)]  # (key, pooled row) in the reference's dict order
        if self.do_fcl:
            for i, c in enumerate(self.combos):
                row[c] = n + i
                self.output_rows.append((c, n + i))
            if not self.no_fusion:
                row["fusion"] = row[self.combos[0]]  # fcl_root is always the first (all-modality) combo, model.py:151
                self.output_rows.append(("fusion", row["fusion"]))
        elif not self.no_fusion:
            row["fusion"] = n
            self.output_rows.append(("fusion", n))
        if self.no_fusion:
            pairs = list(combinations(names, 2))
        elif bimodal:
            pairs = list(combinations(names + ["fusion"], 2))
        else:
            pairs = [(m, "fusion") for m in names]
        plan = []
        for a, b in pairs:
            need = sum(1 << names.index(m) for m in (a, b) if m != "fusion")
            plan.append(("_".join(sorted((a, b))), row[a], row[b], need, 0, 0))
        if self.do_fcl:
            for c in self.combos[1:]:
                cname = "_".join(sorted(names[i] for i in c))
                any_bits = sum(1 << i for i in c)
                if not self.no_fusion:
                    plan.append((f"fcl_fusion|{cname}", row["fusion"], row[c], 0, any_bits, 1))
                if non_fusion_fcl:
                    for m in names:
                        plan.append((f"fcl_{m}|{cname}", row[m], row[c], 1 << names.index(m), any_bits, 1))
        self.loss_names = [p[0] for p in plan]
        self.loss_is_fcl = np.asarray([p[5] for p in plan], dtype=bool)
        # struct mca_loss_pair {int a_row, b_row; uint32 all_mask, any_mask; int is_fcl;}
        arr = np.zeros(len(plan), dtype=np.dtype([("a", "<i4"), ("b", "<i4"), ("all", "<u4"), ("any", "<u4"), ("fcl", "<i4")]))
        for i, p in enumerate(plan):
            arr[i] = (p[1], p[2], p[3], p[4], p[5])
        self.loss_plan = arr
        self.n_pairs = len(plan)


class EAOPlan(StaticPlan):
    """Static tables of the `EAO` ("everything at once") baseline, model.py:481-596: the SAME layers are run once per
    modality and once per modality combination, each pass over the packed tokens of its modalities with key padding
    as the only mask (model.py:544-545,583-587), then mean-pooled (MeanTokenProjectionPool, model.py:257-280).

    Here all passes of a sample are laid out back to back as ONE sequence of N = sum(pass lengths) tokens with a
    block-diagonal static mask (a token attends exactly the tokens of its own pass), which is what the block-sparse
    attention kernels and the token-parallel GEMM / LayerNorm kernels already execute: the single-modality passes come
    first, so the first n_tok tokens are the encoder outputs in modality order and every later block is a replica of one
    of them (`replicas`: (dst offset, src offset, length)).  Key group = block index (<= 32 blocks); `mask_src` lists,
    per block, the modality whose attention_mask pads it.  A row whose pass has no live key at all is never read by
    the pooling (its pooled token is zeros, model.py:270-271), so the fully-masked-row rule needs no per-pass variant.
    Pooled rows are [modalities..., combinations...] = the row map of MCAPretrainingLoss with no_fusion (model.py:181-186).
    """

    def __init__(self, encoder_configs: dict, fusion_combos, fcl: bool, zorro: bool, no_fusion: bool,
                 bimodal_contrastive: bool, non_fusion_fcl: bool):
        self.names = list(encoder_configs.keys())
        self.n_mod = len(self.names)
        if self.n_mod > 8:
            raise AssertionError("at most 8 modalities are supported")
        if not no_fusion:
            # EAO pools no fusion row, so MCAPretrainingLoss would index a pooled token that does not exist (model.py:190)
            raise NotImplementedError("EAO is only defined with no_fusion=True (every shipped *_EAO config)")
        self.eao = True
        self.lengths = [int(encoder_configs[k]["max_tokens"]) for k in self.names]
        self.offsets = [int(x) for x in np.cumsum([0] + self.lengths[:-1])]
        self.zorro, self.fcl, self.no_fusion = bool(zorro), bool(fcl), True
        self.do_fcl = self.fcl and not self.zorro
        self.F = 0
        self.n_tok = int(sum(self.lengths))
        self.combos = fusion_channel_sets(self.n_mod, fusion_combos)
        self.passes = [[i] for i in range(self.n_mod)] + [sorted(c) for c in self.combos]   # model.py:583
        self.R = len(self.passes)
        self.return_token_types = list(range(self.n_mod))                                   # model.py:512
        self.token_types = np.concatenate([np.full(n, i, dtype=np.int64) for i, n in enumerate(self.lengths)])

        mask_src, mask_lengths, blk_off, pass_of_blk, pass_start = [], [], [], [], [0]
        off = 0
        for pi, members in enumerate(self.passes):
            for m in members:
                mask_src.append(m)
                mask_lengths.append(self.lengths[m])
                blk_off.append(off)
                pass_of_blk.append(pi)
                off += self.lengths[m]
            pass_start.append(off)
        self.mask_src, self.mask_lengths, self.block_offsets = mask_src, mask_lengths, blk_off
        self.pass_start = np.asarray(pass_start, dtype=np.int32)
        self.N = off
        self.n_groups = len(mask_src)
        if self.n_groups > 32:
            raise AssertionError("more than 32 key groups (modality blocks over all passes)")
        self.replicas = [(blk_off[g], self.offsets[mask_src[g]], mask_lengths[g]) for g in range(self.n_mod, self.n_groups)]

        keygrp = np.zeros(self.N, dtype=np.uint8)
        rowbits = np.zeros(self.N, dtype=np.uint32)
        tok_pass = np.zeros(self.N, dtype=np.int32)
        pass_bits = [0] * self.R
        for g, pi in enumerate(pass_of_blk):
            pass_bits[pi] |= 1 << g
        for g, (o, n, pi) in enumerate(zip(blk_off, mask_lengths, pass_of_blk)):
            keygrp[o:o + n] = g
            rowbits[o:o + n] = np.uint32(pass_bits[pi])
            tok_pass[o:o + n] = pi
        self.keygrp, self.rowbits, self.tok_pass = keygrp, rowbits, tok_pass
        self.pool_rowbits = np.asarray(pass_bits, dtype=np.uint32)      # pooled row p reads the tokens of pass p
        self.attn_mask = tok_pass[:, None] != tok_pass[None, :]          # True = may not attend (never registered as a buffer)
        self.pool_mask = np.arange(self.R)[:, None] != tok_pass[None, :]
        self._build_tiles()
        self._build_loss_plan(bimodal_contrastive, non_fusion_fcl)

    def _segments(self):
        return list(zip(self.block_offsets, self.mask_lengths))


class MaskPlan(StaticPlan):
    """Schedule of a free-standing `Attention(x, attn_mask=...)` call (model.py:73-105, standalone.attention) derived
    from the dense boolean mask itself (True = may not attend, model.py:90-91): keys whose mask COLUMNS are identical
    form one key group, a query row's bits say which groups it sees.  Every mask the reference builds is of this form
    with at most 23 groups; an arbitrary mask is accepted as long as it has at most 32 distinct
    columns.  Tiles follow the runs of equal key group (runs shorter than a tile are merged into mixed tiles, which the
    kernels mask per key)."""

    def __init__(self, attn_mask, n: int):
        self.eao = False
        self.N = int(n)
        if attn_mask is None:
            attn_mask = np.zeros((self.N, self.N), dtype=bool)
        attn_mask = np.ascontiguousarray(attn_mask, dtype=bool)
        if attn_mask.shape != (self.N, self.N):
            raise AssertionError(f"attn_mask {attn_mask.shape} != {(self.N, self.N)}")
        allowed = ~attn_mask
        cols, first, inv = np.unique(allowed.T, axis=0, return_index=True, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        # number the groups in order of first appearance (keeps the reference's modality order for its own masks)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        keygrp = rank[inv]
        self.n_groups = int(len(cols))
        if self.n_groups > 32:
            raise NotImplementedError(f"attn_mask has {self.n_groups} distinct key columns; the block-sparse attention "
                                      "kernels support masks with at most 32 key groups")
        self.keygrp = keygrp.astype(np.uint8)
        rep = np.asarray(first)[order]                         # one representative key per group
        bits = (allowed[:, rep].astype(np.uint64) << np.arange(self.n_groups, dtype=np.uint64)[None, :]).sum(axis=1)
        self.rowbits = bits.astype(np.uint32)
        self.attn_mask = ~(((self.rowbits[:, None] >> self.keygrp[None, :].astype(np.uint32)) & 1).astype(bool))
        if not np.array_equal(self.attn_mask, attn_mask):
            raise AssertionError("internal error: key-group factorisation does not reproduce attn_mask")
        # segments: maximal runs of one key group; neighbouring runs shorter than a tile merge into one mixed segment
        runs, s = [], 0
        for i in range(1, self.N + 1):
            if i == self.N or self.keygrp[i] != self.keygrp[s]:
                runs.append((s, i - s))
                s = i
        segs = []
        for o, ln in runs:
            if segs and ln < TILE and segs[-1][2]:
                segs[-1] = (segs[-1][0], segs[-1][1] + ln, True)
            else:
                segs.append((o, ln, ln < TILE))
        self._segs = [(o, ln) for o, ln, _ in segs]
        self._build_tiles()

    def _segments(self):
        return self._segs
