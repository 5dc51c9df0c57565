"""Kernel orchestration for the MCA/MMA hot path: owns the flat parameter/gradient buffers, the bf16 operand arena,
the activation workspaces and the launch order of forward, backward and the fused optimiser step.

Everything numerical happens in libmca_b200.so (include/mca_b200.h); this file only sequences launches on torch's
current stream, so a whole step can be captured into a CUDA graph (no host synchronisation, no allocation after the
first call).  Reference call stack being replaced: MCA.forward model.py:448-478 -> MCALayer.forward :117-122 ->
Attention.forward :73-105 / FeedForward :35-54 -> MCAPretrainingLoss.forward :175-233, its autograd backward, and
train_accel_gpu.py:112-119 (zero_grad, backward, clip_grad_norm_, AdamW.step, scheduler.step).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, ops
from .ops import P, S, call
from .plan import StaticPlan

D = 512
DH = 64
ALIGN = 64  # elements; keeps every parameter 256-byte aligned inside the flat buffers


# transformers.get_scheduler names (train_accel_gpu.py:81-86 passes config.lr_scheduler_type) -> device lr_mode
LR_MODES = {"constant": 0, "cosine": 1, "constant_with_warmup": 2, "linear": 3}


def _round_up(x, m):
    return (x + m - 1) // m * m


class Engine:
    def __init__(self, model, plan: StaticPlan, depth: int, heads: int, ff_inner: int, batch_size: int,
                 trunk: bool = True):
        self.model = model
        self.trunk = bool(trunk)   # False: encoders only (free-standing encoder(batch) calls, standalone.py)
        self.plan = plan
        self.depth, self.H, self.I = depth, heads, ff_inner
        if heads * DH != D:
            raise AssertionError("this build supports dim = heads*dim_head = 512 with dim_head = 64")
        self.IP = _round_up(ff_inner, 128)
        self.B = batch_size
        self.N = plan.N
        self.M = self.B * self.N
        self.R = plan.R
        self.eao = bool(getattr(plan, "eao", False))   # EAO baseline: passes stacked in one sequence, mean pooling
        self.device = None
        self.world, self.rank, self.group = 1, 0, None
        self._p2p = None
        self._ws_ready = False
        self._flat_ptrs = None
        self.check_finite = True
        import os
        self.parallel_encoders = os.environ.get("MCA_PARALLEL_ENCODERS", "1") != "0"
        self.precision = "bf16"
        self._xws = None
        # varlen query skipping in the attention kernels (include/mca_b200.h, mca_query_skip_flags): "exact" (default) skips
        # all-padded query tiles only where no result can change, "fast" skips them in every sample, "off" never
        self.varlen = os.environ.get("MCA_VARLEN", "exact").lower()
        if self.varlen not in ("exact", "fast", "off") or self.eao:
            self.varlen = "off"   # (EAO replicates modalities per pass: its masks are per block, not per modality)
        if os.environ.get("MCA_PRECISION", "bf16").lower() in ("fp32", "f32", "float32"):
            self.precision = "fp32"

    # ------------------------------------------------------------------------------------------ parameters
    def _param_list(self):
        return list(self.model.named_parameters())

    def ensure_flat(self):
        """(Re)build the flat fp32 parameter buffer when parameters are not (or no longer) views of it —
        e.g. after model.to(device) or load_state_dict(assign=True)."""
        params = self._param_list()
        dev = params[0][1].device
        if dev.type != "cuda":
            raise _lib.MCAKernelError("mca_paper_b200 runs on CUDA only: move the model to a B200 (no CPU fallback)")
        ptrs = tuple(p.data_ptr() for _, p in params)
        if self._flat_ptrs == ptrs and self.device == dev:
            return
        offs, total = {}, 0
        for name, p in params:
            offs[name] = total
            total += _round_up(p.numel(), ALIGN)
        self.device = dev
        flat, self._flat_peers = self._alloc_flat(total)
        self._flat_mc = self._mc_last
        for name, p in params:
            o = offs[name]
            flat[o:o + p.numel()].copy_(p.detach().reshape(-1).to(torch.float32))
            p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat, self.offs, self.n_flat = flat, offs, total
        self.flat_grad, self._grad_peers = self._alloc_flat(total)
        self._grad_mc = self._mc_last
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.total_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.device = dev
        self._flat_ptrs = tuple(p.data_ptr() for _, p in params)
        self._build_static(dev)
        self._build_pack_descs(dev)
        self._alloc_workspace(dev)

    def _alloc_flat(self, total):
        """fp32 [total] zeros.  Under peer-memory data parallelism the buffer comes from P2P-mapped symmetric memory
        (collective call: every rank allocates in the same order) and the peers' addresses are returned with it."""
        dev = self.device
        if not getattr(self, "_want_symm", False):
            self._mc_last = 0
            return torch.zeros(total, device=dev, dtype=torch.float32), None
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(total, dtype=torch.float32, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        off = int(buf.data_ptr()) - int(hdl.buffer_ptrs[self.rank])
        if off < 0 or off + 4 * total > int(hdl.buffer_size):
            raise _lib.MCAKernelError(f"symmetric-memory tensor outside its allocation (offset {off})")
        self._symm_keep = getattr(self, "_symm_keep", []) + [(buf, hdl)]
        peers = torch.tensor([int(p) + off for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        # NVSwitch multicast mapping of the same allocation (NVLS), 0 when the fabric / driver has none
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self._mc_last = mc + off if mc else 0
        return buf, peers

    def pview(self, name):  # fp32 view of a parameter inside the flat buffer
        p = dict(self._param_list())[name]
        o = self.offs[name]
        return self.flat[o:o + p.numel()].view(p.shape)

    def gview(self, name):
        p = dict(self._param_list())[name]
        o = self.offs[name]
        return self.flat_grad[o:o + p.numel()].view(p.shape)

    # ------------------------------------------------------------------------------------------ static tables
    def _build_static(self, dev):
        pl = self.plan
        t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.keygrp = t(pl.keygrp)
        self.rowbits = t(pl.rowbits.view(np.int32))
        self.pool_rowbits = t(pl.pool_rowbits.view(np.int32))
        self.q_tiles = t(pl.q_tiles_sorted)
        self.tile_grp = t(pl.tile_grp)
        self.kt_list = t(pl.kt_list)
        self.k_tiles = t(pl.tiles)
        self.k_tiles_q = t(pl.k_tiles_q)
        self.qt_list = t(pl.qt_list)
        self.kt_start = t(pl.tiles[:, 0].copy())
        self.kt_len = t(pl.tiles[:, 1].copy())
        self.n_kt = int(pl.tiles.shape[0])
        self.loss_plan = torch.from_numpy(pl.loss_plan.view(np.uint8).copy()).to(dev)
        if self.eao:
            self.pass_start, self.tok_pass = t(pl.pass_start), t(pl.tok_pass)

    # ------------------------------------------------------------------------------------------ weight layouts
    def _build_pack_descs(self, dev):
        """bf16 operand arena (kernel layouts) and the fp32 weight-gradient arena: ONE kernel-layout slab per matrix
        that every k-split of the dW GEMM reduce-adds into (MCA_EPI_F32_ACC); trunk_backward zeroes it first."""
        descs = np.zeros(0, dtype=ops.PACK_DESC_DTYPE)
        rows: List[tuple] = []
        self.w = {}    # name -> (arena offset, rows, ld)
        self.gw = {}   # name -> (partial arena offset, splits, slab elements)
        a_off, g_off = 0, 0

        def add_matrix(key, shape, parts, tokens):
            nonlocal a_off, g_off
            r, c = shape
            tiles = (r // 128) * ((c + 127) // 128)
            splits = ops.effective_splits(tokens, max(1, min(16, 296 // max(1, tiles))))
            self.w[key] = (a_off, r, c)
            self.gw[key] = (g_off, splits, r * c)
            for (pname, prow, pcol, row0, mode, half, scale) in parts:
                rows.append((self.offs[pname], a_off, prow, pcol, c, row0, mode, half, scale, g_off, 1, r * c))
            a_off += _round_up(r * c, 512)
            g_off += _round_up(r * c, 512)

        I, IP, M = self.I, self.IP, self.M
        for l in range(self.depth):
            p = f"layers.{l}."
            add_matrix(p + "qkv", (3 * D, D), [(p + "attn.to_q.weight", D, D, 0, 0, 0, DH ** -0.5),
                                               (p + "attn.to_kv.weight", 2 * D, D, D, 0, 0, 1.0)], M)
            add_matrix(p + "out", (D, D), [(p + "attn.to_out.weight", D, D, 0, 0, 0, 1.0)], M)
            add_matrix(p + "ff1", (2 * IP, D), [(p + "ff.feedforward.0.weight", 2 * I, D, 0, 1, I, 1.0)], M)
            add_matrix(p + "ff2", (D, IP), [(p + "ff.feedforward.2.weight", D, I, 0, 0, 0, 1.0)], M)
        if not self.eao and self.trunk:
            add_matrix("attn_pool.kv", (2 * D, D), [("attn_pool.to_kv.weight", 2 * D, D, 0, 0, 0, 1.0)], M)
        self.enc_kpad = {}
        for name, enc in zip(self.plan.names, self.model.encoder_specs):
            pre = f"encoders.{name}."
            L = enc["max_tokens"]
            if enc["type"] == "EmbeddedSequenceEncoder":
                kin = enc["input_size"]
                kp = _round_up(kin, 64)
                self.enc_kpad[name] = kp
                add_matrix(pre + "proj", (D, kp), [(pre + "token_encoder.1.weight", D, kin, 0, 0, 0, 1.0)], self.B * L)
            elif enc["type"] in ("TabularEncoder", "SparseTabularEncoder"):
                add_matrix(pre + "proj", (D, D), [(pre + "value_encoder.linear2.weight", D, D, 0, 0, 0, 1.0)], self.B * L)
            elif enc["type"] == "PatchEncoder":
                kin = int(np.prod(enc.get("patch_size", (16, 16))))
                kp = _round_up(kin, 64)
                self.enc_kpad[name] = kp
                add_matrix(pre + "proj", (D, kp), [(pre + "batch_to_tokens.2.weight", D, kin, 0, 0, 0, 1.0)], self.B * L)
        self.arena = torch.zeros(a_off, device=dev, dtype=torch.bfloat16)
        self.garena = torch.zeros(g_off, device=dev, dtype=torch.float32)
        pd = np.zeros(len(rows), dtype=ops.PACK_DESC_DTYPE)
        ud = np.zeros(len(rows), dtype=ops.PACK_DESC_DTYPE)
        for i, (src, aoff, prow, pcol, ld, row0, mode, half, scale, goff, splits, slab) in enumerate(rows):
            pd[i] = (src, aoff, prow, pcol, ld, row0, mode, half, scale, 1, 0)
            ud[i] = (src, goff, prow, pcol, ld, row0, mode, half, scale, splits, slab)
        self.n_desc = len(rows)
        self.pack_descs = torch.from_numpy(pd.view(np.uint8).copy()).to(dev)
        self.unpack_descs = torch.from_numpy(ud.view(np.uint8).copy()).to(dev)

    def W(self, key):  # bf16 kernel-layout weight
        o, r, c = self.w[key]
        return self.arena[o:o + r * c].view(r, c)

    def GW(self, key):  # fp32 kernel-layout gradient slab + the number of k-splits that reduce into it
        o, s, slab = self.gw[key]
        r, c = self.w[key][1], self.w[key][2]
        return self.garena[o:o + slab].view(1, r, c), s

    def pack_weights(self):
        call("mca_pack_weights", P(self.flat), P(self.arena), P(self.pack_descs), self.n_desc, S())
        if self.precision == "fp32":
            call("mca_x_pack_weights_split", P(self.flat), P(self._exact_ws()["arena3"]), P(self.pack_descs), self.n_desc, S())

    # ------------------------------------------------------------------------------------------ fp32-parity mode
    def set_precision(self, mode: str):
        """"bf16" (default): bf16 tensor-core operands, fp32 accumulation — within 2e-2 of the fp32 reference.
        "fp32": the fp32-parity FORWARD (csrc/exact.cu): every contraction as a 3-term bf16 split product on the same
        tcgen05 GEMM (16 significand bits per operand), fp32 attention core, GEGLU and pooling — loss and embeddings
        within 1e-3 of the reference (train_accel_gpu.py:21: fp32, no autocast).  The backward stays on the bf16 kernels
        and consumes the bf16 copies this forward saves (north_star sets no 1e-3 bar for gradients).  Also selectable
        with MCA_PRECISION=fp32."""
        if mode not in ("bf16", "fp32"):
            raise ValueError(f"precision {mode!r}: expected 'bf16' or 'fp32'")
        if mode == "fp32" and self.eao:
            raise NotImplementedError("fp32-parity mode covers the MCA / MMA models (the EAO baseline pools bf16 tokens)")
        self.precision = mode
        if self._ws_ready:
            self.pack_weights()

    def _exact_ws(self):
        if self._xws is None:
            dev, M, IP = self.device, self.M, self.IP
            f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            x = {"arena3": torch.zeros(3 * self.arena.numel(), device=dev, dtype=torch.bfloat16),
                 "x1_32": f32(M, D), "qkv32": f32(M, 3 * D), "ao32": f32(M, D), "x3_32": f32(M, D), "u32": f32(M, 2 * IP),
                 "xf32": f32(M, D), "kv32": f32(M, 2 * D),
                 "a3": torch.empty(M * 3 * max(IP, D), device=dev, dtype=torch.bfloat16), "enc": {}}
            for name, enc in zip(self.plan.names, self.model.encoder_specs):
                rows = self.B * enc["max_tokens"]
                if enc["type"] in ("EmbeddedSequenceEncoder", "PatchEncoder"):
                    x["enc"][name] = torch.empty(rows, 3 * self.enc_kpad[name], device=dev, dtype=torch.bfloat16)
                elif enc["type"] in ("TabularEncoder", "SparseTabularEncoder"):
                    x["enc"][name] = torch.empty(rows, 3 * D, device=dev, dtype=torch.bfloat16)
            self._xws = x
        return self._xws

    def W3(self, key):  # [W_hi | W_lo | W_hi] kernel-layout weight of the split product
        o, r, c = self.w[key]
        return self._exact_ws()["arena3"][3 * o:3 * o + 3 * r * c].view(r, 3 * c)

    def _split(self, src32, rows, cols):
        """fp32 [rows, cols] -> the [hi | hi | lo] operand [rows, 3*cols] in the shared scratch."""
        a3 = self._exact_ws()["a3"][:rows * 3 * cols].view(rows, 3 * cols)
        call("mca_x_split_f32", P(src32), src32.stride(0), P(a3), rows, cols, cols, 0, S())
        return a3

    def _proj(self, name, enc_key, a3, kp, bias, out):
        """Encoder projection of the split operand a3 [rows, 3*kp] (or the bf16 one) into z."""
        rows = a3.shape[0]
        ops.gemm(a3, 0, self.W3(enc_key), 0, rows, D, 3 * kp, _lib.EPI_F32, out, bias=bias)

    def trunk_forward_exact(self, batch):
        """trunk_forward in fp32-parity mode (set_precision): same buffers for the backward, fp32 values in between."""
        ws, X, M, IP = self.ws, self._exact_ws(), self.M, self.IP
        ws["nonfinite"].zero_()
        ws["drop_ctr"].add_(1)
        self.build_offsets(batch)
        self.encode(batch)
        for l in range(self.depth):
            p = f"layers.{l}."
            gamma, beta = self.pview(p + "norm.gamma"), self.model.layers[l].norm.beta
            # x1 = LN(x)                                                     (model.py:118, shared LayerNorm: Q1)
            ops.layernorm512_fwd(ws["xa"][l], gamma, beta, X["x1_32"], ws["x1_16"][l], ws["st1"][l], M)
            ops.gemm(self._split(X["x1_32"], M, D), 0, self.W3(p + "qkv"), 0, M, 3 * D, 3 * D, _lib.EPI_F32, X["qkv32"])
            call("mca_cast_f32_bf16", P(X["qkv32"]), 3 * D, P(ws["qkv"][l]), 3 * D, M, 3 * D, S())
            call("mca_x_attn_fwd_f32", P(X["qkv32"]), P(self.rowbits), P(self.keygrp), P(ws["padding"]), P(ws["vmean"]),
                 P(X["ao32"]), P(ws["ao"][l]), P(ws["lse"][l]), self.B, self.N, self.H, S())
            # x2 = attn(x1) Wo^T + x1                                        (model.py:119)
            ops.gemm(self._split(X["ao32"], M, D), 0, self.W3(p + "out"), 0, M, D, 3 * D, _lib.EPI_RESID, ws["x2"][l],
                     aux0=X["x1_32"], ldaux=D)
            # x3 = LN(x2);  x' = geglu(x3 W1^T) W2^T + x3                    (model.py:120-122, 35-54)
            ops.layernorm512_fwd(ws["x2"][l], gamma, beta, X["x3_32"], ws["x3_16"][l], ws["st2"][l], M)
            ops.gemm(self._split(X["x3_32"], M, D), 0, self.W3(p + "ff1"), 0, M, 2 * IP, 3 * D, _lib.EPI_F32, X["u32"])
            h3 = X["a3"][:M * 3 * IP].view(M, 3 * IP)
            call("mca_x_geglu_f32", P(X["u32"]), P(h3), P(ws["h"][l]), P(ws["u"][l]), M, IP, S())
            ops.gemm(h3, 0, self.W3(p + "ff2"), 0, M, D, 3 * IP, _lib.EPI_RESID, ws["xa"][l + 1], aux0=X["x3_32"], ldaux=D)
        ops.layernorm512_fwd(ws["xa"][self.depth], self.pview("norm.gamma"), self.model.norm.beta, X["xf32"], ws["xf_16"],
                             ws["stF"], M)
        # attention pooling on the final-normed tokens (model.py:470-473)
        ops.gemm(self._split(X["xf32"], M, D), 0, self.W3("attn_pool.kv"), 0, M, 2 * D, 3 * D, _lib.EPI_F32, X["kv32"])
        call("mca_cast_f32_bf16", P(X["kv32"]), 2 * D, P(ws["kvp"]), 2 * D, M, 2 * D, S())
        rt, wq, wo = self.pview("return_tokens"), self.pview("attn_pool.to_q.weight"), self.pview("attn_pool.to_out.weight")
        ops.small_gemm(rt, D, 1, wq, D, 1, ws["qp"], D, self.R, D, D, alpha=DH ** -0.5)
        call("mca_x_pool_attn_fwd_f32", P(ws["qp"]), P(X["kv32"]), P(ws["padding"]), P(self.keygrp), P(self.pool_rowbits),
             P(ws["probs"]), P(ws["fm"]), P(ws["po"]), self.B, self.H, self.R, self.N, S())
        ops.small_gemm(ws["po"].view(self.B * self.R, D), D, 1, wo, D, 1, ws["pooled"].view(self.B * self.R, D), D,
                       self.B * self.R, D, D, add=rt, ldadd=D, add_rows=self.R)
        return ws["pooled"]

    # ------------------------------------------------------------------------------------------ workspaces
    def _alloc_workspace(self, dev):
        M, N, B, H, R, IP, L = self.M, self.N, self.B, self.H, self.R, self.IP, self.depth
        f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        b16 = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
        u8 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.uint8)
        i32 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.int32)
        ws = {}
        ws["xa"] = [f32(M, D) for _ in range(L + 1)]
        ws["st1"] = [f32(M, 2) for _ in range(L)]
        ws["st2"] = [f32(M, 2) for _ in range(L)]
        ws["x1_16"] = [b16(M, D) for _ in range(L)]
        ws["qkv"] = [b16(M, 3 * D) for _ in range(L)]
        ws["ao"] = [b16(M, D) for _ in range(L)]
        ws["lse"] = [f32(B, H, N) for _ in range(L)]
        ws["x2"] = [f32(M, D) for _ in range(L)]
        ws["x3_16"] = [b16(M, D) for _ in range(L)]
        ws["u"] = [b16(M, 2 * IP) for _ in range(L)]
        ws["h"] = [b16(M, IP) for _ in range(L)]
        n_blk, blk_tok = len(self.plan.mask_src), int(sum(self.plan.mask_lengths))
        ws["y16"] = b16(M, D)    # branch output of out-proj / FF2 (consumed at once by the fused add + LayerNorm)
        ws["dy16"] = b16(M, D)   # branch gradient of the dX GEMMs (added inside the LayerNorm backward)
        ws["stF"], ws["xf_16"] = f32(M, 2), b16(M, D)
        ws["pooled"] = f32(B, R, D)
        if self.eao:
            ws["pool_cnt"] = f32(B, R)
            ws["pool_scratch"] = f32(int(ops.fn("mca_mean_pool_scratch_floats")(B, R)))
        elif self.trunk:
            ws["kvp"] = b16(M, 2 * D)
            ws["qp"], ws["probs"], ws["fm"] = f32(R, D), f32(B, H, R, N), u8(B, R)
            ws["po"] = f32(B, R, D)
        ws["vmean"] = torch.zeros(B, D, device=dev, dtype=torch.float32)
        # offsets / masks
        ws["padding"], ws["pad_mod"] = u8(B, N), u8(B * blk_tok)
        ws["present"], ws["live_count"] = u8(B, n_blk), i32(B, n_blk)   # the first n_mod columns are the modalities
        ws["live_idx"], ws["cu_live"] = i32(B, N), i32(B * n_blk + 1)
        ws["kt_class"], ws["any_absent"], ws["nonfinite"] = u8(B, self.n_kt), i32(1), i32(1)
        ws["kt_live"] = i32(B, self.n_kt, 4)
        ws["skip_ok"] = u8(B)     # varlen query skipping flags (mca_query_skip_flags)
        # encoders
        ws["enc"] = {}
        for name, enc in zip(self.plan.names, self.model.encoder_specs):
            rows = B * enc["max_tokens"]
            kind = enc["type"]
            if kind == "SequenceEncoder":  # a table lookup: no projection workspace
                ws["enc"][name] = {"flags": u8(int(enc.get("num_embeddings", 36602)))}
                continue
            e = {"z": f32(rows, D), "st_out": f32(rows, 2)}
            if kind in ("EmbeddedSequenceEncoder", "PatchEncoder"):
                kp = self.enc_kpad[name]
                e.update(y=b16(rows, kp), st_in=f32(rows, 2), dz16=b16(rows, D), dz32=f32(rows, D), dy=f32(rows, kp))
                if kind == "PatchEncoder":
                    e.update(ptok=f32(rows, int(np.prod(enc.get("patch_size", (16, 16))))), mask=u8(B, enc["max_tokens"]))
            else:
                e.update(h1=b16(rows, D), vpad=u8(rows), dz16=b16(rows, D), dz32=f32(rows, D), dh1=f32(rows, D))
                if kind == "SparseTabularEncoder":
                    e.update(flags=u8(int(enc.get("num_embeddings", 36602))))
            ws["enc"][name] = e
        ws["drop_ctr"] = torch.zeros(1, device=dev, dtype=torch.int64)  # dropout counter: one tick per forward
        # loss
        nP = self.plan.n_pairs
        ws["losses"], ws["summary"], ws["w_default"] = f32(nP), f32(4), f32(nP)
        # backward scratch
        ws["dx_a"], ws["dx_b"], ws["d16"] = f32(M, D), f32(M, D), b16(M, D)
        ws["du"], ws["dattn"], ws["dqkv"] = b16(M, 2 * IP), b16(M, D), b16(M, 3 * D)
        ws["dq_acc"], ws["delta"], ws["ucorr"] = f32(M, D), f32(B, H, N), f32(B, D)
        ws["dxf"] = f32(M, D)
        if not self.eao and self.trunk:
            ws["dkvp"] = b16(M, 2 * D)
            ws["pool_ds"], ws["dqp"], ws["dpo"] = f32(B, H, R, N), f32(R, D), f32(B, R, D)
        self.ws = ws
        self._gather_ws = None
        self._ws_ready = True

    def set_distributed(self, world: int, rank: int, group=None, p2p: Optional[bool] = None):
        """Data parallel over `group`.  p2p (default: env MCA_P2P != "0"): keep the gathered pooled block, its gathered
        gradient and the barrier flags in P2P-mapped symmetric memory and exchange them with our own push / pull
        kernels over NVLink (no NCCL all_gather / reduce_scatter in the step, the whole forward + loss + backward is
        one CUDA graph); otherwise the NCCL collectives are used between graph segments."""
        import os
        self.world, self.rank, self.group = world, rank, group
        self._gather_ws = None
        self._p2p = None
        want = (os.environ.get("MCA_P2P", "1") != "0") if p2p is None else p2p
        if world > 1 and want:
            # rebuild the flat parameter / gradient buffers in symmetric memory (peers pull gradient shards from them and
            # push updated parameter shards into them), then the small exchange buffers
            err = None
            try:
                self._want_symm = True
                self._flat_ptrs = None
                self.ensure_flat()
                self._setup_p2p()
            except Exception as exc:  # no P2P mapping between these GPUs / symmetric memory unavailable
                err = exc
            # the choice of exchange protocol is collective: one rank on the NCCL form while the others spin on peer
            # flags would deadlock, so every rank learns whether ALL ranks succeeded before committing to either path
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device or "cuda")
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                import warnings
                why = f"{type(err).__name__}: {err}" if err is not None else "a peer rank could not set it up"
                warnings.warn(f"peer-memory data parallelism unavailable ({why}); "
                              "using the NCCL all_gather / reduce_scatter / all_reduce form on every rank")
                self._want_symm, self._p2p, self._flat_ptrs = False, None, None
                self.ensure_flat()

    def _setup_p2p(self):
        """One symmetric allocation per rank: [pooled_all G*B*R*D | dpooled_all G*B*R*D | flags (G uint32, padded to 64
        words) | gradient sum-of-squares slots (G doubles, padded to 32)]."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        dev, G = self.device, self.world
        n_p = n_d = G * self.B * self.R * D
        total = n_p + n_d + 64 + 64
        buf = symm.empty(total, dtype=torch.float32, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        # buffer_ptrs are the bases of the symmetric ALLOCATIONS; the tensor may sit at an offset inside its own (the
        # same offset on every rank, the allocation is symmetric)
        off = int(buf.data_ptr()) - int(hdl.buffer_ptrs[self.rank])
        if off < 0 or off + 4 * total > int(hdl.buffer_size):
            raise _lib.MCAKernelError(f"symmetric-memory tensor outside its allocation (offset {off})")
        base = [int(p) + off for p in hdl.buffer_ptrs]
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        self._p2p = {
            "buf": buf, "handle": hdl,
            "pooled_all": buf[:n_p].view(G * self.B, self.R, D),
            "dpooled_all": buf[n_p:n_p + n_d].view(G * self.B, self.R, D),
            "pooled_peers": i64(base),
            "dall_peers": i64([b + 4 * n_p for b in base]),
            "flags_peers": i64([b + 4 * (n_p + n_d) for b in base]),
            "slots": buf[n_p + n_d + 64:n_p + n_d + 128].view(torch.float64),
            "slots_peers": i64([b + 4 * (n_p + n_d + 64) for b in base]),
            "sumsq_local": torch.zeros(1, dtype=torch.float64, device=dev),
            "grad_peers": self._grad_peers, "param_peers": self._flat_peers,
            # in-switch reduction / replication of the optimiser exchange (multimem.ld_reduce / multimem.st) when every
            # rank got a multicast mapping of both flat buffers
            "grad_mc": self._grad_mc, "param_mc": self._flat_mc, "multimem": False,
            "epoch": torch.zeros(1, dtype=torch.int32, device=dev),
            # pinned host memory (device-addressable under UVA): still readable after the barrier kernel trapped
            "err": torch.zeros(1, dtype=torch.int32).pin_memory(),
        }
        import os
        want_mc = os.environ.get("MCA_MULTIMEM", "1") != "0" and self._grad_mc != 0 and self._flat_mc != 0
        ok = torch.tensor([1 if want_mc else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # collective choice, like the P2P-or-NCCL one
        self._p2p["multimem"] = bool(int(ok.item()))
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)  # every rank's flags are zero before anybody raises one

    def xgpu_barrier(self, payload=None):
        """Flag barrier over peer memory; `payload` (a device double) is delivered to slot `rank` of every rank's slots."""
        p = self._p2p
        call("mca_xgpu_barrier", P(p["flags_peers"]), self.world, self.rank, P(p["epoch"]), P(p["err"]), P(payload),
             P(p["slots_peers"]) if payload is not None else None, S())

    def shard(self):
        """Flat-buffer range [off, off + n) this rank owns in the sharded optimiser step."""
        per = Engine.shard_size(self.n_flat, self.world)
        off = min(self.rank * per, self.n_flat)
        return off, max(0, min(per, self.n_flat - off))

    @staticmethod
    def shard_size(n_flat: int, world: int) -> int:
        """Elements per rank of the sharded optimiser step (float4 units; the last shard may be shorter)."""
        return ((n_flat + world - 1) // world + 3) // 4 * 4

    def check_p2p(self):
        """Raises if a peer missed a flag barrier.  The flag lives in pinned host memory, so this is a plain host read
        (no device synchronisation): Trainer calls it every step and checkpoint.save_state before writing."""
        if self._p2p is not None and int(self._p2p["err"][0]) != 0:
            raise _lib.MCAKernelError(f"rank {int(self._p2p['err'][0]) - 1} did not reach mca_xgpu_barrier within 10 s "
                                      f"(seen by rank {self.rank}); the step was aborted")

    def _gather_buffers(self):
        if self._gather_ws is None:
            GB = self.world * self.B
            dev = self.device
            self._gather_ws = {
                "pooled_all": torch.zeros(GB, self.R, D, device=dev, dtype=torch.float32),
                "dpooled_all": torch.zeros(GB, self.R, D, device=dev, dtype=torch.float32),
                "dpooled": torch.zeros(self.B, self.R, D, device=dev, dtype=torch.float32),
                "dscale": torch.zeros(1, device=dev, dtype=torch.float32),
            }
        return self._gather_ws

    # ------------------------------------------------------------------------------------------ forward
    def build_offsets(self, batch):
        pl, ws = self.plan, self.ws
        masks = []
        for n, enc in zip(pl.names, self.model.encoder_specs):
            if enc["type"] == "PatchEncoder":  # the encoder derives its own pad mask (encoders.py:273)
                masks.append(self._patchify(n, enc, batch[n]["values"]))
            else:
                masks.append(batch[n]["attention_mask"])
        for m, L in zip(masks, pl.lengths):
            if m.shape != (self.B, L):
                raise AssertionError(f"attention_mask shape {tuple(m.shape)} != {(self.B, L)} (batch must equal batch_size, model.py:454)")
        self._mask_keepalive = [m if m.is_contiguous() else m.contiguous() for m in masks]
        # one entry per modality block of the packed sequence (EAO: a modality pads every pass it takes part in)
        blocks = [self._mask_keepalive[src] for src in pl.mask_src]
        n = len(blocks)
        ptrs = (ctypes.c_void_p * n)(*[m.data_ptr() for m in blocks])
        es = (ctypes.c_int * n)(*[m.element_size() for m in blocks])
        lens = (ctypes.c_int * n)(*pl.mask_lengths)
        call("mca_build_offsets", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(es, ctypes.c_void_p),
             ctypes.cast(lens, ctypes.c_void_p), n, self.B, self.N, P(self.kt_start), P(self.kt_len), self.n_kt,
             P(ws["padding"]), P(ws["pad_mod"]), P(ws["present"]), P(ws["live_count"]), P(ws["live_idx"]),
             P(ws["cu_live"]), P(ws["kt_class"]), P(ws["kt_live"]), P(ws["any_absent"]), S())
        if self.varlen != "off":
            call("mca_query_skip_flags", P(ws["present"]), n, len(pl.names), self.B, 1 if self.varlen == "exact" else 2,
                 P(ws["skip_ok"]), S())

    def _patchify(self, name, enc, values):
        """values [B,H,W] -> e['ptok'] [B*L, p1*p2] and the all-pad mask e['mask'] [B,L] (encoders.py:243-246,273)."""
        e = self.ws["enc"][name]
        if values.dtype != torch.float32 or not values.is_contiguous():
            values = values.to(torch.float32).contiguous()
        e["values"] = values
        p1, p2 = (int(x) for x in enc.get("patch_size", (16, 16)))
        Bv, Hh, Ww = values.shape
        if Bv != self.B or (Hh // p1) * (Ww // p2) != enc["max_tokens"]:
            raise AssertionError(f"{(Hh // p1) * (Ww // p2)} - {enc['max_tokens']}")  # encoders.py:270
        call("mca_patchify", P(values), self.B, Hh, Ww, p1, p2, -10000.0, P(e["ptok"]), P(e["mask"]), S())
        return e["mask"]

    def _idx(self, t):
        return t if (t.dtype == torch.int64 and t.is_contiguous()) else t.to(torch.int64).contiguous()

    def _pad_mod(self, i):
        pl = self.plan
        o = pl.offsets[i] * self.B
        return self.ws["pad_mod"][o:o + self.B * pl.lengths[i]]

    # The modality encoders are independent chains of small kernels (a [400..12000]-row LayerNorm / GEMM / LayerNorm does
    # not fill 148 SMs and costs ~10 us of launch + pipeline latency per kernel whatever its size): forward and backward
    # run them on parallel streams — graph branches once the step is captured — forked from and joined to the step's stream.
    def _branches(self, n):
        """[(stream, is_main)] for n independent chains: the largest first on the current stream, the others on side streams
        that wait for everything enqueued so far."""
        main = torch.cuda.current_stream()
        if n <= 1 or not self.parallel_encoders:
            return [main] * n, main
        if getattr(self, "_side_streams", None) is None or len(self._side_streams) < n - 1:
            self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(n - 1)]
            self._fork_ev = torch.cuda.Event()
            self._join_ev = [torch.cuda.Event() for _ in range(n - 1)]
        self._fork_ev.record(main)
        streams = [main] + self._side_streams[:n - 1]
        for st in streams[1:]:
            st.wait_event(self._fork_ev)
        return streams, main

    def _join(self, streams, main):
        k = 0
        for st in streams:
            if st is not main:
                self._join_ev[k].record(st)
                main.wait_event(self._join_ev[k])
                k += 1

    def _encoder_order(self):
        """Encoder indices, largest token count first (it takes the main stream)."""
        specs = self.model.encoder_specs
        return sorted(range(len(specs)), key=lambda i: -int(specs[i]["max_tokens"]) * int(specs[i].get("input_size", 512)))

    def encode(self, batch):
        """Encoders write straight into the packed token buffer xa[0] (no `pack` copy, model.py:464)."""
        pl = self.plan
        order = self._encoder_order()
        streams, main = self._branches(len(order))
        for st, i in zip(streams, order):
            with torch.cuda.stream(st):
                self._encode_one(i, batch)
        self._join(streams, main)
        x0 = self.ws["xa"][0]
        if pl.F:
            call("mca_broadcast_rows", P(self.pview("fusion_tokens")), P(x0), pl.F, D, self.B, self.N, pl.n_tok, S())
        if self.eao:
            # every modality is encoded once (model.py:576-578); the passes it takes part in read replicas (model.py:584)
            xv = x0.view(self.B, self.N, D)
            for dst, src, L in pl.replicas:
                xv[:, dst:dst + L].copy_(xv[:, src:src + L])

    def _encode_one(self, i, batch):
        pl, ws = self.plan, self.ws
        x0 = ws["xa"][0]
        name, enc = pl.names[i], self.model.encoder_specs[i]
        if True:
            pre = f"encoders.{name}."
            L = enc["max_tokens"]
            rows = self.B * L
            e = ws["enc"][name]
            pad = self._pad_mod(i)
            if enc["type"] == "EmbeddedSequenceEncoder":
                tok = batch[name]["tokens"]
                if tok.dtype != torch.float32 or not tok.is_contiguous():
                    tok = tok.to(torch.float32).contiguous()
                e["tokens"] = tok
                kin, kp = enc["input_size"], self.enc_kpad[name]
                call("mca_layernorm_in_fwd", P(tok), P(self.pview(pre + "token_encoder.0.weight")),
                     P(self.pview(pre + "token_encoder.0.bias")), P(pad), P(e["y"]), P(e["st_in"]), kin, kp, rows,
                     P(ws["nonfinite"]), S())
                if self.precision == "fp32":
                    y3 = self._exact_ws()["enc"][name]
                    call("mca_x_layernorm_in_split", P(tok), P(self.pview(pre + "token_encoder.0.weight")),
                         P(self.pview(pre + "token_encoder.0.bias")), P(pad), P(y3), kin, kp, rows, S())
                    self._proj(name, pre + "proj", y3, kp, self.pview(pre + "token_encoder.1.bias"), e["z"])
                else:
                    ops.gemm(e["y"], 0, self.W(pre + "proj"), 0, rows, D, kp, _lib.EPI_F32, e["z"],
                             bias=self.pview(pre + "token_encoder.1.bias"))
                ops.layernorm512_fwd(e["z"], self.pview(pre + "token_encoder.2.weight"),
                                     self.pview(pre + "token_encoder.2.bias"), x0, None, e["st_out"], rows, pad=pad,
                                     pe=self.model.encoders[name].positional_encoder.pe, seg_len=L,
                                     out_rows_per_b=self.N, out_row_off=pl.offsets[i])
            elif enc["type"] == "TabularEncoder":
                vals = batch[name]["values"]
                if vals.dtype != torch.float32 or not vals.is_contiguous():
                    vals = vals.to(torch.float32).contiguous()
                e["values"] = vals
                emb = self.pview(pre + "token_encoder.embedding.weight")
                call("mca_embedding_renorm", P(emb), L, D, 1.0, S())  # nn.Embedding(max_norm=1.0), in place
                call("mca_tabular_fwd", P(vals), P(self.pview(pre + "value_encoder.linear1.weight")),
                     P(self.pview(pre + "value_encoder.linear1.bias")), P(e["h1"]), P(e["vpad"]),
                     float(enc.get("max_value", 10000)), float(enc.get("padding_idx", -1)), D, rows, S())
                self._tabular_proj(name, pre, enc, e, vals, rows)
                ops.layernorm512_fwd(e["z"], self.pview(pre + "value_encoder.norm.weight"),
                                     self.pview(pre + "value_encoder.norm.bias"), x0, None, e["st_out"], rows,
                                     pad=e["vpad"], pe=emb, seg_len=L, out_rows_per_b=self.N, out_row_off=pl.offsets[i])
            elif enc["type"] == "SequenceEncoder":
                # encoders.py:161-166: Embedding(tokens) (looked-up rows renormalised in place) + sinusoidal PE
                tok = e["idx"] = self._idx(batch[name]["tokens"])
                emb = self.pview(pre + "token_encoder.embedding.weight")
                V = emb.shape[0]
                call("mca_embedding_renorm_indexed", P(emb), P(tok), rows, V, D, 1.0, P(e["flags"]), P(ws["nonfinite"]), S())
                call("mca_embedding_gather", P(emb), P(tok), V, self.B, L, D,
                     P(self.model.encoders[name].positional_encoder.pe), P(x0), self.N, pl.offsets[i], 0, S())
            elif enc["type"] == "SparseTabularEncoder":
                # encoders.py:114-120: Embedding(indices) + ContinuousValueEncoder(data)
                idx = e["idx"] = self._idx(batch[name]["indices"])
                vals = batch[name]["data"]
                if vals.dtype != torch.float32 or not vals.is_contiguous():
                    vals = vals.to(torch.float32).contiguous()
                e["values"] = vals
                emb = self.pview(pre + "token_encoder.embedding.weight")
                V = emb.shape[0]
                call("mca_embedding_renorm_indexed", P(emb), P(idx), rows, V, D, 1.0, P(e["flags"]), P(ws["nonfinite"]), S())
                call("mca_tabular_fwd", P(vals), P(self.pview(pre + "value_encoder.linear1.weight")),
                     P(self.pview(pre + "value_encoder.linear1.bias")), P(e["h1"]), P(e["vpad"]),
                     float(enc.get("max_value", 10000)), float(enc.get("padding_idx", 0)), D, rows, S())
                self._tabular_proj(name, pre, enc, e, vals, rows)
                ops.layernorm512_fwd(e["z"], self.pview(pre + "value_encoder.norm.weight"),
                                     self.pview(pre + "value_encoder.norm.bias"), x0, None, e["st_out"], rows,
                                     pad=e["vpad"], seg_len=L, out_rows_per_b=self.N, out_row_off=pl.offsets[i])
                call("mca_embedding_gather", P(emb), P(idx), V, self.B, L, D, None, P(x0), self.N, pl.offsets[i], 1, S())
            elif enc["type"] == "PatchEncoder":
                # encoders.py:268-274: LN(in) -> Linear -> LN(512) of every patch (nothing is zeroed for padded patches),
                # + learned position embedding, dropout in training mode; patches / mask were made by _patchify
                kin, kp = e["ptok"].shape[1], self.enc_kpad[name]
                call("mca_layernorm_in_fwd", P(e["ptok"]), P(self.pview(pre + "batch_to_tokens.1.weight")),
                     P(self.pview(pre + "batch_to_tokens.1.bias")), None, P(e["y"]), P(e["st_in"]), kin, kp, rows,
                     P(ws["nonfinite"]), S())
                if self.precision == "fp32":
                    y3 = self._exact_ws()["enc"][name]
                    call("mca_x_layernorm_in_split", P(e["ptok"]), P(self.pview(pre + "batch_to_tokens.1.weight")),
                         P(self.pview(pre + "batch_to_tokens.1.bias")), None, P(y3), kin, kp, rows, S())
                    self._proj(name, pre + "proj", y3, kp, self.pview(pre + "batch_to_tokens.2.bias"), e["z"])
                else:
                    ops.gemm(e["y"], 0, self.W(pre + "proj"), 0, rows, D, kp, _lib.EPI_F32, e["z"],
                             bias=self.pview(pre + "batch_to_tokens.2.bias"))
                ops.layernorm512_fwd(e["z"], self.pview(pre + "batch_to_tokens.3.weight"),
                                     self.pview(pre + "batch_to_tokens.3.bias"), x0, None, e["st_out"], rows,
                                     pe=self.pview(pre + "embedding.weight"), seg_len=L, out_rows_per_b=self.N,
                                     out_row_off=pl.offsets[i])
                self._dropout(i, enc, x0)
            else:
                raise NotImplementedError(f"unknown encoder type {enc['type']}")

    def _tabular_proj(self, name, pre, enc, e, vals, rows):
        """linear2 of ContinuousValueEncoder (encoders.py:60-75) on h1; fp32-parity mode recomputes h1 as a split operand."""
        if self.precision == "fp32":
            h3 = self._exact_ws()["enc"][name]
            call("mca_x_tabular_split", P(vals), P(self.pview(pre + "value_encoder.linear1.weight")),
                 P(self.pview(pre + "value_encoder.linear1.bias")), P(h3), float(enc.get("max_value", 10000)), D, rows, S())
            self._proj(name, pre + "proj", h3, D, self.pview(pre + "value_encoder.linear2.bias"), e["z"])
        else:
            ops.gemm(e["h1"], 0, self.W(pre + "proj"), 0, rows, D, D, _lib.EPI_F32, e["z"],
                     bias=self.pview(pre + "value_encoder.linear2.bias"))

    def _dropout(self, i, enc, rows32):
        """nn.Dropout of PatchEncoder (encoders.py:274) on the modality's rows of `rows32` (tokens in the forward, their
        gradient in the backward: same counter -> same mask).  Inactive in eval mode, like the module."""
        p = float(enc.get("dropout", 0.1))
        if p > 0.0 and self.model.training:
            call("mca_dropout_rows", P(rows32), self.B, enc["max_tokens"], D, self.N, self.plan.offsets[i], p,
                 (torch.initial_seed() + 977 * i) & 0xFFFFFFFFFFFFFFFF, P(self.ws["drop_ctr"]), S())

    def _skip_ok(self):
        return self.ws["skip_ok"] if self.varlen != "off" else None

    def set_varlen(self, mode: str):
        """"exact" / "fast" / "off": which all-padded query tiles the attention kernels leave out (mca_query_skip_flags)."""
        if mode not in ("exact", "fast", "off"):
            raise ValueError(f"varlen mode {mode!r}: expected 'exact', 'fast' or 'off'")
        if self.eao and mode != "off":
            raise NotImplementedError("varlen query skipping covers the MCA / MMA models")
        self.varlen = mode

    def attention_fwd(self, qkv, out, lse):
        ws = self.ws
        call("mca_attn_fwd", P(qkv), P(self.q_tiles), int(self.q_tiles.shape[0]), P(self.kt_list), P(self.k_tiles),
             self.n_kt, P(self.rowbits), P(self.keygrp), P(self.tile_grp), P(ws["kt_class"]), P(ws["kt_live"]),
             P(ws["padding"]), P(self._skip_ok()), P(ws["any_absent"]), P(ws["vmean"]), P(out), P(lse), self.B, self.N, self.H, S())

    def trunk_forward(self, batch):
        """encoders -> depth x [LN, QKV, attention, out-proj(+res), LN, FF1(GEGLU), FF2(+res)] -> LN -> pooling."""
        if self.precision == "fp32":
            return self.trunk_forward_exact(batch)
        ws, M, IP = self.ws, self.M, self.IP
        ws["nonfinite"].zero_()
        ws["drop_ctr"].add_(1)
        self.build_offsets(batch)
        self.encode(batch)
        # Residual wiring of model.py:117-122 (quirk Q1): x1 = LN(x); x2 = attn(x1) + x1; x3 = LN(x2); x' = ff(x3) + x3,
        # with ONE LayerNorm per layer.  The branch GEMMs write bf16; the residual add is fused with the NEXT LayerNorm,
        # which recomputes the normed residual from (x, stats) instead of reading an fp32 copy of it.
        gamma0, beta0 = self.pview("layers.0.norm.gamma"), self.model.layers[0].norm.beta
        ops.layernorm512_fwd(ws["xa"][0], gamma0, beta0, None, ws["x1_16"][0], ws["st1"][0], M)
        for l in range(self.depth):
            p = f"layers.{l}."
            gamma, beta = self.pview(p + "norm.gamma"), self.model.layers[l].norm.beta
            ops.gemm(ws["x1_16"][l], 0, self.W(p + "qkv"), 0, M, 3 * D, D, _lib.EPI_BF16, ws["qkv"][l])
            self.attention_fwd(ws["qkv"][l], ws["ao"][l], ws["lse"][l])
            ops.gemm(ws["ao"][l], 0, self.W(p + "out"), 0, M, D, D, _lib.EPI_BF16, ws["y16"])
            # x2 = LN(x) + y ; x3_16 = LN(x2)
            ops.add_layernorm512_fwd(ws["xa"][l], ws["st1"][l], gamma, beta, ws["y16"], ws["x2"][l], gamma, beta,
                                     ws["x3_16"][l], ws["st2"][l], M)
            ops.gemm(ws["x3_16"][l], 0, self.W(p + "ff1"), 0, M, 2 * IP, D, _lib.EPI_GEGLU, ws["h"][l], ld0=IP,
                     out1=ws["u"][l], ld1=2 * IP)
            ops.gemm(ws["h"][l], 0, self.W(p + "ff2"), 0, M, D, IP, _lib.EPI_BF16, ws["y16"])
            # x' = LN(x2) + y ; next operand = LN_next(x') (next layer's norm, or the final norm)
            if l + 1 < self.depth:
                gn, bn = self.pview(f"layers.{l + 1}.norm.gamma"), self.model.layers[l + 1].norm.beta
                ops.add_layernorm512_fwd(ws["x2"][l], ws["st2"][l], gamma, beta, ws["y16"], ws["xa"][l + 1], gn, bn,
                                         ws["x1_16"][l + 1], ws["st1"][l + 1], M)
            else:
                ops.add_layernorm512_fwd(ws["x2"][l], ws["st2"][l], gamma, beta, ws["y16"], ws["xa"][l + 1],
                                         self.pview("norm.gamma"), self.model.norm.beta, ws["xf_16"], ws["stF"], M)
        if self.eao:
            # mean pooling of every pass over its live tokens (model.py:562-563, 257-280)
            call("mca_mean_pool_fwd", P(ws["xf_16"]), P(ws["padding"]), P(self.pass_start), self.B, self.N, self.R, D,
                 P(ws["pooled"]), P(ws["pool_cnt"]), P(ws["pool_scratch"]), S())
            return ws["pooled"]
        # attention pooling on the final-normed tokens (model.py:470-473)
        rt, wq, wo = self.pview("return_tokens"), self.pview("attn_pool.to_q.weight"), self.pview("attn_pool.to_out.weight")
        main, side = self._pool_side()
        with torch.cuda.stream(side):   # the pooled queries only depend on parameters: under the K/V projection
            ops.small_gemm(rt, D, 1, wq, D, 1, ws["qp"], D, self.R, D, D, alpha=DH ** -0.5)
        ops.gemm(ws["xf_16"], 0, self.W("attn_pool.kv"), 0, M, 2 * D, D, _lib.EPI_BF16, ws["kvp"])
        if side is not main:
            self._pool_ev[3].record(side)
            main.wait_event(self._pool_ev[3])
        call("mca_pool_attn_fwd", P(ws["qp"]), P(ws["kvp"]), P(ws["padding"]), P(self.keygrp), P(self.pool_rowbits),
             P(ws["probs"]), P(ws["fm"]), P(ws["po"]), self.B, self.H, self.R, self.N, S())
        # pooled[b, r] = po[b, r] Wo^T + return_tokens[r]
        ops.small_gemm(ws["po"].view(self.B * self.R, D), D, 1, wo, D, 1, ws["pooled"].view(self.B * self.R, D), D,
                       self.B * self.R, D, D, add=rt, ldadd=D, add_rows=self.R)
        return ws["pooled"]

    def loss_forward(self, pooled):
        """All-gather of the pooled block (one collective instead of 2 per pair) + fused all-pairs InfoNCE."""
        ws, g = self.ws, self._gather_buffers()
        s = self.pview("loss.loss_fn.logit_scale")
        lf = self.model.loss.loss_fn
        lo = float(lf.logit_scale_min if lf.logit_scale_min is not None else -1e30)
        hi = float(lf.logit_scale_max if lf.logit_scale_max is not None else 1e30)
        if self.world > 1 and self._p2p is not None:
            # peer-memory all-gather: push this rank's block into slot `rank` of every rank's gathered buffer, barrier
            n = self.B * self.R * D
            call("mca_p2p_push_rows", P(pooled), P(self._p2p["pooled_peers"]), self.rank * n, n, self.world, S())
            self.xgpu_barrier()
            pooled_all = self._p2p["pooled_all"]
        elif self.world > 1:
            torch.distributed.all_gather_into_tensor(g["pooled_all"], pooled.contiguous(), group=self.group)
            pooled_all = g["pooled_all"]
        else:
            pooled_all = pooled
        self._pooled_all = pooled_all
        call("mca_contrastive_allpairs_fwd", P(pooled_all), P(ws["present"]), P(self.loss_plan), self.plan.n_pairs, P(s),
             self.B, self.world * self.B, self.R, D, len(self.plan.mask_src), self.rank, lo, hi,
             P(ws["losses"]), P(ws["summary"]), P(ws["w_default"]), S())
        return ws["losses"], ws["summary"]

    # ------------------------------------------------------------------------------------------ backward
    def loss_backward(self, w):
        """w[p] = dL/d losses[p].  Returns dL/d pooled [B,R,D] (after the reduce-scatter of gathered gradients)."""
        ws, g = self.ws, self._gather_buffers()
        GB = self.world * self.B
        if self.world > 1 and self._p2p is not None:
            p = self._p2p
            # the peers finished pulling last step's gradient slices before they reached this step's first barrier
            p["dpooled_all"].zero_()
            g["dscale"].zero_()
            s = self.pview("loss.loss_fn.logit_scale")
            call("mca_contrastive_allpairs_bwd", P(self._pooled_all), P(ws["present"]), P(self.loss_plan),
                 self.plan.n_pairs, P(s), self.B, GB, self.R, D, len(self.plan.mask_src), self.rank, P(w), P(p["dpooled_all"]),
                 P(g["dscale"]), S())
            self.xgpu_barrier()
            n = self.B * self.R * D
            call("mca_p2p_reduce_rows", P(p["dall_peers"]), self.rank * n, P(g["dpooled"]), n, self.world, S())
            self.gview("loss.loss_fn.logit_scale").add_(g["dscale"].view(()))
            return g["dpooled"]
        dall = g["dpooled_all"] if self.world > 1 else g["dpooled"]
        dall.zero_()
        g["dscale"].zero_()
        s = self.pview("loss.loss_fn.logit_scale")
        call("mca_contrastive_allpairs_bwd", P(self._pooled_all), P(ws["present"]), P(self.loss_plan), self.plan.n_pairs,
             P(s), self.B, GB, self.R, D, len(self.plan.mask_src), self.rank, P(w), P(dall), P(g["dscale"]), S())
        if self.world > 1:
            torch.distributed.reduce_scatter_tensor(g["dpooled"], dall, op=torch.distributed.ReduceOp.SUM, group=self.group)
        self.gview("loss.loss_fn.logit_scale").add_(g["dscale"].view(()))
        return g["dpooled"]

    def _dw(self, key, dY, X, rows_w, cols_w, tokens):
        """dW[rows_w, cols_w] += dY^T X over `tokens` rows, both operands consumed MN-major; the k-splits reduce-add
        into one slab (zeroed at the start of trunk_backward)."""
        part, splits = self.GW(key)
        ops.gemm(dY, 1, X, 1, rows_w, cols_w, tokens, _lib.EPI_F32_ACC, part, ld0=cols_w, k_splits=splits)

    def trunk_backward(self, dpooled):
        """Reverse of trunk_forward; parameter gradients land in self.flat_grad (state_dict layout)."""
        ws, pl, M, IP, B, R, H, N = self.ws, self.plan, self.M, self.IP, self.B, self.R, self.H, self.N
        self.garena.zero_()  # every dW GEMM reduce-adds its k-splits into this arena
        if self.eao:
            call("mca_mean_pool_bwd", P(dpooled), P(ws["padding"]), P(self.tok_pass), P(ws["pool_cnt"]), B, N, R, D,
                 P(ws["dxf"]), S())
        else:
            self._pool_backward(dpooled)
        dx, dx_alt = ws["dx_a"], ws["dx_b"]
        ops.layernorm512_bwd(ws["dxf"], ws["xa"][self.depth], ws["stF"], self.pview("norm.gamma"), dx, ws["d16"],
                             self.gview("norm.gamma"), None, M)
        self._layers_backward(dx, dx_alt)

    def _pool_backward(self, dpooled):
        """Attention pooling backward (model.py:470-473): gradients of return_tokens / attn_pool.* and dxf."""
        ws, M, B, R, H, N = self.ws, self.M, self.B, self.R, self.H, self.N
        rt, wq, wo = self.pview("return_tokens"), self.pview("attn_pool.to_q.weight"), self.pview("attn_pool.to_out.weight")
        dp2 = dpooled.reshape(B * R, D)
        po2 = ws["po"].view(B * R, D)
        # pooled = po Wo^T + rt.  The R-row products are latency-bound (13 us each whatever their size): the ones off the
        # critical path dpo -> pool backward -> dxf run on a side stream (a graph branch), joined before mca_unpack_grads.
        main, side = self._pool_side()
        sc = DH ** -0.5
        with torch.cuda.stream(side):
            call("mca_batchsum_rows", P(dp2), P(self.gview("return_tokens")), R, D, B, R, 0, 1, S())
            ops.small_gemm(dp2, 1, D, po2, 1, D, self.gview("attn_pool.to_out.weight"), D, D, D, B * R)   # dWo = dp^T po
        ops.small_gemm(dp2, D, 1, wo, 1, D, ws["dpo"].view(B * R, D), D, B * R, D, D)                        # dpo = dp Wo
        call("mca_pool_attn_bwd", P(ws["dpo"]), P(ws["qp"]), P(ws["kvp"]), P(ws["probs"]), P(ws["fm"]), P(ws["pool_ds"]),
             P(ws["dkvp"]), P(ws["dqp"]), B, H, R, N, S())
        if side is not main:
            self._pool_ev[1].record(main)
            side.wait_event(self._pool_ev[1])
        with torch.cuda.stream(side):
            ops.small_gemm(ws["dqp"], 1, D, rt, 1, D, self.gview("attn_pool.to_q.weight"), D, D, D, R, alpha=sc)  # dWq
            ops.small_gemm(ws["dqp"], D, 1, wq, 1, D, self.gview("return_tokens"), D, R, D, D, alpha=sc, accumulate=True)
        if side is not main:
            self._pool_ev[2].record(side)
            self._pool_pending = True
        # through the K/V projection and the final LayerNorm
        ops.gemm(ws["dkvp"], 0, self.W("attn_pool.kv"), 1, M, D, 2 * D, _lib.EPI_F32, ws["dxf"])
        self._dw("attn_pool.kv", ws["dkvp"], ws["xf_16"], 2 * D, D, M)

    def _pool_side(self):
        """(current stream, side stream forked from it) for the small pooling products; the same stream twice when
        MCA_PARALLEL_ENCODERS=0."""
        main = torch.cuda.current_stream()
        if not self.parallel_encoders:
            return main, main
        if getattr(self, "_pool_stream", None) is None:
            self._pool_stream = torch.cuda.Stream(device=self.device)
            self._pool_ev = [torch.cuda.Event() for _ in range(4)]
            self._pool_pending = False
        self._pool_ev[0].record(main)
        self._pool_stream.wait_event(self._pool_ev[0])
        return main, self._pool_stream

    def _pool_join(self):
        if getattr(self, "_pool_pending", False):
            torch.cuda.current_stream().wait_event(self._pool_ev[2])
            self._pool_pending = False

    def _layers_backward(self, dx, dx_alt):
        ws, M, IP = self.ws, self.M, self.IP
        for l in reversed(range(self.depth)):
            p = f"layers.{l}."
            gamma, dgamma = self.pview(p + "norm.gamma"), self.gview(p + "norm.gamma")
            # x4 = h W2^T + x3 ; h = geglu(u) ; u = x3 W1^T
            ops.gemm(ws["d16"], 0, self.W(p + "ff2"), 1, M, IP, D, _lib.EPI_GEGLU_BWD, ws["du"], ld0=2 * IP,
                     aux0=ws["u"][l], ldaux=2 * IP)
            self._dw(p + "ff2", ws["d16"], ws["h"][l], D, IP, M)
            ops.gemm(ws["du"], 0, self.W(p + "ff1"), 1, M, D, 2 * IP, _lib.EPI_BF16, ws["dy16"])   # branch part of dx3
            self._dw(p + "ff1", ws["du"], ws["x3_16"][l], 2 * IP, D, M)
            # dx3 = dx (residual path) + dy16 (branch); LayerNorm backward -> dx2
            ops.layernorm512_bwd(dx, ws["x2"][l], ws["st2"][l], gamma, dx_alt, ws["d16"], dgamma, None, M, dy_delta=ws["dy16"])
            dx, dx_alt = dx_alt, dx
            # x2 = ao Wo^T + x1
            ops.gemm(ws["d16"], 0, self.W(p + "out"), 1, M, D, D, _lib.EPI_BF16, ws["dattn"])
            self._dw(p + "out", ws["d16"], ws["ao"][l], D, D, M)
            self.attention_bwd(l)
            ops.gemm(ws["dqkv"], 0, self.W(p + "qkv"), 1, M, D, 3 * D, _lib.EPI_BF16, ws["dy16"])      # branch part of dx1
            self._dw(p + "qkv", ws["dqkv"], ws["x1_16"][l], 3 * D, D, M)
            ops.layernorm512_bwd(dx, ws["xa"][l], ws["st1"][l], gamma, dx_alt, ws["d16"], dgamma, None, M, dy_delta=ws["dy16"])
            dx, dx_alt = dx_alt, dx
        if self.eao:
            # the gradient of a modality's tokens is the sum over every pass that read them
            dv = dx.view(self.B, self.N, D)
            for dst, src, L in self.plan.replicas:
                dv[:, src:src + L].add_(dv[:, dst:dst + L])
        self.encode_backward(dx)
        self._pool_join()
        call("mca_unpack_grads", P(self.flat_grad), P(self.garena), P(self.unpack_descs), self.n_desc, S())

    def attention_bwd(self, l):
        ws = self.ws
        call("mca_attn_bwd", P(ws["qkv"][l]), P(ws["ao"][l]), P(ws["dattn"]), P(ws["lse"][l]), P(self.k_tiles_q),
             self.n_kt, P(self.qt_list), P(self.k_tiles), int(self.q_tiles.shape[0]), P(self.rowbits), P(self.keygrp),
             P(self.tile_grp), P(ws["padding"]), P(ws["kt_class"]), P(self._skip_ok()), P(ws["delta"]), P(ws["ucorr"]), P(ws["dq_acc"]),
             P(ws["dqkv"]),
             self.B, self.N, self.H, S())

    def encode_backward(self, dx0):
        pl = self.plan
        order = self._encoder_order()
        streams, main = self._branches(len(order))
        if pl.F:
            call("mca_batchsum_rows", P(dx0), P(self.gview("fusion_tokens")), pl.F, D, self.B, self.N, pl.n_tok, 1, S())
        for st, i in zip(streams, order):
            with torch.cuda.stream(st):
                self._encode_backward_one(i, dx0)
        self._join(streams, main)

    def _encode_backward_one(self, i, dx0):
        pl, ws = self.plan, self.ws
        name, enc = pl.names[i], self.model.encoder_specs[i]
        if True:
            pre = f"encoders.{name}."
            L = enc["max_tokens"]
            rows = self.B * L
            e = ws["enc"][name]
            if enc["type"] == "EmbeddedSequenceEncoder":
                kin, kp = enc["input_size"], self.enc_kpad[name]
                pad = self._pad_mod(i)
                ops.layernorm512_bwd(dx0, e["z"], e["st_out"], self.pview(pre + "token_encoder.2.weight"), e["dz32"],
                                     e["dz16"], self.gview(pre + "token_encoder.2.weight"),
                                     self.gview(pre + "token_encoder.2.bias"), rows, pad=pad, seg_len=L,
                                     out_rows_per_b=self.N, out_row_off=pl.offsets[i])
                call("mca_colsum", P(e["dz32"]), D, P(self.gview(pre + "token_encoder.1.bias")), D, rows, S())
                self._dw(pre + "proj", e["dz16"], e["y"], D, kp, rows)
                ops.gemm(e["dz16"], 0, self.W(pre + "proj"), 1, rows, kp, D, _lib.EPI_F32, e["dy"])
                call("mca_layernorm_in_param_bwd", P(e["dy"]), kp, P(e["tokens"]), P(e["st_in"]), P(pad),
                     P(self.gview(pre + "token_encoder.0.weight")), P(self.gview(pre + "token_encoder.0.bias")), kin,
                     rows, S())
            elif enc["type"] == "SequenceEncoder":
                emb = self.pview(pre + "token_encoder.embedding.weight")
                call("mca_embedding_scatter_add", P(dx0), P(e["idx"]), emb.shape[0], self.B, L, D, self.N, pl.offsets[i],
                     int(enc.get("padding_idx", 0)) % emb.shape[0], P(self.gview(pre + "token_encoder.embedding.weight")), S())
            elif enc["type"] == "PatchEncoder":
                kin, kp = e["ptok"].shape[1], self.enc_kpad[name]
                self._dropout(i, enc, dx0)
                call("mca_batchsum_rows", P(dx0), P(self.gview(pre + "embedding.weight")), L, D, self.B, self.N,
                     pl.offsets[i], 1, S())
                ops.layernorm512_bwd(dx0, e["z"], e["st_out"], self.pview(pre + "batch_to_tokens.3.weight"), e["dz32"],
                                     e["dz16"], self.gview(pre + "batch_to_tokens.3.weight"),
                                     self.gview(pre + "batch_to_tokens.3.bias"), rows, seg_len=L,
                                     out_rows_per_b=self.N, out_row_off=pl.offsets[i])
                call("mca_colsum", P(e["dz32"]), D, P(self.gview(pre + "batch_to_tokens.2.bias")), D, rows, S())
                self._dw(pre + "proj", e["dz16"], e["y"], D, kp, rows)
                ops.gemm(e["dz16"], 0, self.W(pre + "proj"), 1, rows, kp, D, _lib.EPI_F32, e["dy"])
                call("mca_layernorm_in_param_bwd", P(e["dy"]), kp, P(e["ptok"]), P(e["st_in"]), None,
                     P(self.gview(pre + "batch_to_tokens.1.weight")), P(self.gview(pre + "batch_to_tokens.1.bias")), kin,
                     rows, S())
            elif enc["type"] in ("TabularEncoder", "SparseTabularEncoder"):
                sparse = enc["type"] == "SparseTabularEncoder"
                ops.layernorm512_bwd(dx0, e["z"], e["st_out"], self.pview(pre + "value_encoder.norm.weight"), e["dz32"],
                                     e["dz16"], self.gview(pre + "value_encoder.norm.weight"),
                                     self.gview(pre + "value_encoder.norm.bias"), rows, pad=e["vpad"], seg_len=L,
                                     out_rows_per_b=self.N, out_row_off=pl.offsets[i])
                call("mca_colsum", P(e["dz32"]), D, P(self.gview(pre + "value_encoder.linear2.bias")), D, rows, S())
                self._dw(pre + "proj", e["dz16"], e["h1"], D, D, rows)
                ops.gemm(e["dz16"], 0, self.W(pre + "proj"), 1, rows, D, D, _lib.EPI_F32, e["dh1"])
                call("mca_tabular_bwd", P(e["dh1"]), P(e["values"]), P(self.pview(pre + "value_encoder.linear1.weight")),
                     P(self.pview(pre + "value_encoder.linear1.bias")),
                     P(self.gview(pre + "value_encoder.linear1.weight")), P(self.gview(pre + "value_encoder.linear1.bias")),
                     None, None, float(enc.get("max_value", 10000)), float(enc.get("padding_idx", -1)), D, rows, S())
                gemb = self.gview(pre + "token_encoder.embedding.weight")
                if sparse:
                    call("mca_embedding_scatter_add", P(dx0), P(e["idx"]), gemb.shape[0], self.B, L, D, self.N,
                         pl.offsets[i], int(enc.get("padding_idx", 0)) % gemb.shape[0], P(gemb), S())
                else:
                    call("mca_batchsum_rows", P(dx0), P(gemb), L, D, self.B, self.N, pl.offsets[i], 1, S())
                    gemb[int(enc.get("padding_idx", -1)) % L].zero_()  # nn.Embedding padding_idx row gets no gradient

    # ------------------------------------------------------------------------------------------ optimiser
    def configure_optimizer(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=2.0,
                            schedule="constant", warmup_steps=0, total_steps=1, scheduler_stride=1):
        """scheduler_stride: scheduler.step() calls per optimiser step.  The reference prepares its scheduler with
        accelerate (train_accel_gpu.py:93), whose wrapper advances it num_processes times per step, so a run on G GPUs
        walks the cosine G times faster than its step count; pass G to reproduce that (Trainer does)."""
        if schedule not in LR_MODES:
            raise ValueError(f"unknown lr schedule {schedule!r}: expected one of {sorted(LR_MODES)} "
                             "(transformers.get_scheduler names, train_accel_gpu.py:81-86)")
        self.adamw_cfg = ops.AdamWCfg(lr, betas[0], betas[1], eps, weight_decay, max_norm,
                                      LR_MODES[schedule], warmup_steps, total_steps,
                                      max(1, int(scheduler_stride)))

    def lr_at(self, step: int) -> float:
        """Host mirror of the device schedule (optim.cu scheduled_lr): the learning rate the `step`-th optimiser step
        (1-based) uses == transformers.get_cosine_schedule_with_warmup after (step-1)*stride scheduler steps."""
        import math
        c = self.adamw_cfg
        if c.lr_mode == 0:
            return float(c.lr)
        cur = float(step - 1) * max(1, c.sched_stride)
        if cur < c.warmup_steps:
            return float(c.lr) * cur / max(1.0, float(c.warmup_steps))
        if c.lr_mode == 2:
            return float(c.lr)
        if c.lr_mode == 3:
            return float(c.lr) * max(0.0, (c.total_steps - cur) / max(1.0, float(c.total_steps - c.warmup_steps)))
        prog = (cur - c.warmup_steps) / max(1.0, float(c.total_steps - c.warmup_steps))
        return float(c.lr) * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))

    def snapshot_state(self):
        """Device copies of everything an optimiser step mutates (weights, moments, step, dropout counter)."""
        snap = {k: getattr(self, k).clone() for k in ("flat", "exp_avg", "exp_avg_sq", "step_dev")}
        snap["drop_ctr"] = self.ws["drop_ctr"].clone()
        return snap

    def restore_state(self, snap):
        for k in ("flat", "exp_avg", "exp_avg_sq", "step_dev"):
            getattr(self, k).copy_(snap[k])
        self.ws["drop_ctr"].copy_(snap["drop_ctr"])
        self.pack_weights()

    def optimizer_state_dict(self):
        """AdamW state in `torch.optim.AdamW.state_dict()` layout ({'state': {i: {step, exp_avg, exp_avg_sq}},
        'param_groups': [...]}, parameters numbered in `named_parameters()` order) so that a checkpoint written through
        accelerate's `save_state` (train_accel_gpu.py:122-123,133-134) can be produced / consumed.  Under peer-memory
        data parallelism every rank only holds the moments of its own shard: they are gathered here (one all-gather of
        the flat buffers over the process group; checkpoints are rare, the step itself never needs the full moments)."""
        self.ensure_flat()
        m, v = self.exp_avg, self.exp_avg_sq
        if self.world > 1 and self._p2p is not None:
            off, n = self.shard()
            per = Engine.shard_size(self.n_flat, self.world)
            full = []
            for buf in (m, v):
                mine = torch.zeros(per, device=self.device, dtype=torch.float32)
                mine[:n] = buf[off:off + n]
                out = torch.empty(per * self.world, device=self.device, dtype=torch.float32)
                torch.distributed.all_gather_into_tensor(out, mine, group=self.group)
                full.append(out[:self.n_flat])
            m, v = full
        step = int(self.step_dev.item())
        state = {}
        for i, (name, p) in enumerate(self._param_list()):
            o = self.offs[name]
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m[o:o + p.numel()].view(p.shape).clone(),
                        "exp_avg_sq": v[o:o + p.numel()].view(p.shape).clone()}
        c = self.adamw_cfg
        group = {"lr": c.lr, "betas": (c.beta1, c.beta2), "eps": c.eps, "weight_decay": c.weight_decay, "amsgrad": False,
                 "params": list(range(len(state)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict (also accepts the dict of a torch.optim.AdamW over the same parameters)."""
        self.ensure_flat()
        step = 0
        for i, (name, p) in enumerate(self._param_list()):
            st = sd["state"].get(i)
            if st is None:
                continue
            o = self.offs[name]
            self.exp_avg[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1).to(self.device, torch.float32))
            self.exp_avg_sq[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1).to(self.device, torch.float32))
            step = max(step, int(float(st["step"])))
        self.step_dev.fill_(step)

    def optimizer_step(self):
        """Gradient all-reduce (data parallel mean, train_accel_gpu.py:93,115) + clip + AdamW + weight re-pack.
        Peer-memory form: reduce-scatter by pulling gradient shards, AdamW on the own shard, all-gather by pushing the
        updated parameters — three flag barriers, no NCCL call, capturable in the step's CUDA graph."""
        if self.world > 1 and self._p2p is not None:
            p = self._p2p
            off, n = self.shard()
            self.xgpu_barrier()                                  # every rank's gradient buffer is complete
            if p["multimem"]:
                call("mca_dp_reduce_shard_mc", p["grad_mc"], P(self.flat_grad), off, n, self.world, P(p["sumsq_local"]), S())
            else:
                call("mca_dp_reduce_shard", P(p["grad_peers"]), P(self.flat_grad), off, n, self.world, P(p["sumsq_local"]), S())
            self.xgpu_barrier(payload=p["sumsq_local"])          # + the G partial sums of squares, everywhere
            if p["multimem"]:
                call("mca_dp_adamw_shard_mc", P(p["param_peers"]), p["param_mc"], self.world, self.rank, P(self.flat_grad),
                     P(self.exp_avg), P(self.exp_avg_sq), off, n, P(p["slots"]), P(self.step_dev), P(self.total_norm),
                     1.0 / self.world, ctypes.addressof(self.adamw_cfg), S())
            else:
                call("mca_dp_adamw_shard", P(p["param_peers"]), self.world, self.rank, P(self.flat_grad), P(self.exp_avg),
                     P(self.exp_avg_sq), off, n, P(p["slots"]), P(self.step_dev), P(self.total_norm), 1.0 / self.world,
                     ctypes.addressof(self.adamw_cfg), S())
            self.xgpu_barrier()                                  # every parameter shard has landed
            self.pack_weights()
            return
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_grad, op=torch.distributed.ReduceOp.SUM, group=self.group)
        call("mca_clip_adamw_step", P(self.flat), P(self.flat_grad), P(self.exp_avg), P(self.exp_avg_sq), self.n_flat,
             P(self.sumsq), P(self.step_dev), P(self.total_norm), 1.0 / self.world, ctypes.addressof(self.adamw_cfg), S())
        self.pack_weights()

    def forward_backward(self, batch):
        """One fused training pass without autograd: forward, loss, backward.  Returns the loss summary tensor."""
        pooled = self.trunk_forward(batch)
        _, summary = self.loss_forward(pooled)
        self.flat_grad.zero_()
        dpooled = self.loss_backward(self.ws["w_default"])
        self.trunk_backward(dpooled)
        return summary
