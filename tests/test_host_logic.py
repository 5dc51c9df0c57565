"""CPU: host-side logic of the product — static plan (bit-exact vs oracle and vs reference hashes), tile schedule
properties, config loader, state_dict schema, synthetic batches, C-ABI symbol export, loud failure without CUDA."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from mca_paper_b200 import _lib, config as C, ops, synthetic as S
from mca_paper_b200.model import MCA
from mca_paper_b200.plan import StaticPlan
from oracle import mca_oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_plan(kw):
    return StaticPlan(kw["encoder_configs"], kw["num_fusion_tokens"], kw["fusion_combos"], kw["fcl"], kw["zorro"],
                      kw["no_fusion"], kw["bimodal_contrastive"], kw["non_fusion_fcl"])


@pytest.mark.parametrize("name", ["CMU_config1", "CMU_config1_z", "CMU_config1_d40", "TCGA_config1"])
def test_plan_masks_bit_exact(name):
    kw = C.get_model_config(C.named_config(name))
    p, t = make_plan(kw), O.static_tables(kw)
    assert np.array_equal(p.token_types, t["token_types"].numpy())
    assert np.array_equal(p.attn_mask, t["attn_mask"].numpy())
    assert np.array_equal(p.pool_mask, t["pool_mask"].numpy())
    assert p.return_token_types == t["return_token_types"] and p.combos == t["combos"]
    plan, _ = O.loss_plan(kw, t)
    assert [x["name"] for x in plan] == p.loss_names
    for x, y in zip(plan, p.loss_plan):
        assert (x["a"], x["b"]) == (y["a"], y["b"])
        assert sum(1 << i for i in x["all"]) == y["all"] and sum(1 << i for i in x["any"]) == y["any"]
        assert bool(x["fcl"]) == bool(y["fcl"])


def test_plan_matches_reference_hashes():
    static = torch.load(H.GOLDEN_DIR + "/static_tables.pt", weights_only=False)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for name, g in static.items():
        p = make_plan(C.get_model_config(C.named_config(name)))
        assert sha(p.attn_mask) == g["attn_mask_sha256"] and sha(p.pool_mask) == g["pool_mask_sha256"]
        assert sha(p.token_types) == g["token_types_sha256"]
        assert p.allowed_pairs == g["attn_allowed_pairs"] and p.loss_names == g["loss_names"]


@pytest.mark.parametrize("name", ["CMU_config1", "CMU_config1_z", "TCGA_config1"])
def test_tile_schedule_covers_exactly_the_allowed_pairs(name):
    p = make_plan(C.get_model_config(C.named_config(name)))
    allowed = ~p.attn_mask
    covered = np.zeros_like(allowed)
    for qs, ql, off, cnt in p.q_tiles:
        for ki, flag in p.kt_list[off:off + cnt]:
            ks, kl = p.tiles[ki]
            blk = allowed[qs:qs + ql, ks:ks + kl]
            assert blk.any()
            assert (flag == 0) == bool(blk.all())
            assert not covered[qs:qs + ql, ks:ks + kl].any()
            covered[qs:qs + ql, ks:ks + kl] = True
    assert not (allowed & ~covered).any()          # nothing allowed is skipped
    # tiles partition [0, N) and never straddle a modality boundary
    assert p.tiles[0, 0] == 0 and (p.tiles[:, 0] + p.tiles[:, 1])[-1] == p.N
    assert np.array_equal(p.tiles[1:, 0], (p.tiles[:, 0] + p.tiles[:, 1])[:-1]) and p.tiles[:, 1].max() <= 128
    for s, l in p.tiles:
        assert len(set(p.keygrp[s:s + l].tolist()) - set(range(p.n_mod))) <= max(0, p.n_groups - p.n_mod)
    # transposed (backward) schedule lists the same pairs
    fwd = {(qi, int(ki)) for qi, (_, _, off, cnt) in enumerate(p.q_tiles) for ki, _ in p.kt_list[off:off + cnt]}
    bwd = {(int(qi), ki) for ki, (_, _, off, cnt) in enumerate(p.k_tiles_q) for qi, _ in p.qt_list[off:off + cnt]}
    assert fwd == bwd


def test_group_bitmasks_reconstruct_masks():
    p = make_plan(C.get_model_config(C.named_config("CMU_config1")))
    rec = ~(((p.rowbits[:, None] >> p.keygrp[None, :].astype(np.uint32)) & 1).astype(bool))
    assert np.array_equal(rec, p.attn_mask)


@pytest.mark.parametrize("name", ["CMU_config1", "CMU_config1_z", "TCGA_config1"])
def test_state_dict_schema_matches_reference(name):
    static = torch.load(H.GOLDEN_DIR + "/static_tables.pt", weights_only=False)[name]
    model = MCA(**C.get_model_config(C.named_config(name)))
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert mine == static["state_dict_schema"]
    assert [k for k, _ in model.named_parameters()] == static["param_names"]
    assert sum(p.numel() for p in model.parameters()) == static["n_params"]


def test_config_defaults_and_model_kwargs():
    cfg = C.training_config({"layers": 5})
    assert cfg.hidden_size == 512 and cfg.heads == 8 and cfg.dim_head == 64 and cfg.ff_mult == 4   # utils/config.py:40-44
    kw = C.get_model_config(C.named_config("CMU_config1"))
    assert set(kw) == {"dim", "depth", "heads", "dim_head", "ff_mult", "num_fusion_tokens", "encoder_configs", "batch_size",
                       "fcl", "fcl_root", "bimodal_contrastive", "non_fusion_fcl", "fusion_combos", "zorro", "eao",
                       "no_fusion", "mean_pool"}
    MCA(**kw)  # swallows `eao`


def test_bad_fusion_token_count_asserts():
    kw = C.get_model_config(C.named_config("CMU_config1"))
    kw["num_fusion_tokens"] = 90   # not divisible by 11 combos (model.py:416-417)
    with pytest.raises(AssertionError):
        MCA(**kw)


def test_synthetic_batches():
    cfg = C.named_config("CMU_config1_d40")
    b = S.make_batch(cfg, seed=3, variant="dropout_ragged")
    for name, enc in cfg["encoder_configs"].items():
        t, m = b[name]["tokens"], b[name]["attention_mask"]
        assert t.shape == (8, enc["max_tokens"], enc["input_size"]) and m.dtype == torch.bool
        assert (t[m] == 0).all()
        # suffix padding only (encoders.py:339-341)
        assert (m[:, 1:].int() - m[:, :-1].int() >= 0).all()
    tb = S.make_batch(C.named_config("TCGA_config1"), seed=1, variant="tcga")
    assert tb["protein"]["attention_mask"].dtype == torch.long
    assert torch.equal(tb["protein"]["attention_mask"].bool(), tb["protein"]["values"] == -10000)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mca_b200.h")).read()
    declared = set(re.findall(r"^int (mca_\w+)\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(ops.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.mca_version() >= 100


def test_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg = C.tiny_config("cmu")
    model = MCA(**C.get_model_config(cfg))
    with pytest.raises(_lib.MCAKernelError):
        model(S.make_batch(cfg))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mca_paper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_optimizer_shards_partition_the_flat_buffer():
    """Peer-memory data parallelism: rank r owns flat elements [off, off + n); the shards must tile [0, n_flat)
    exactly, in float4 units, for every world size."""
    from types import SimpleNamespace
    from mca_paper_b200.engine import Engine
    for n_flat in (64, 4096, 17_413_888, 19_149_376):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for rank in range(world):
                off, n = Engine.shard(SimpleNamespace(n_flat=n_flat, world=world, rank=rank))
                assert off % 4 == 0 and n % 4 == 0 and n >= 0
                assert off == min(pos, n_flat)
                pos = off + n
            assert pos == n_flat


def _eao_plan(kw):
    from mca_paper_b200.plan import EAOPlan
    return EAOPlan(kw["encoder_configs"], kw["fusion_combos"], kw["fcl"], kw["zorro"], kw["no_fusion"],
                   kw["bimodal_contrastive"], kw["non_fusion_fcl"])


def test_eao_plan_stacks_passes_block_diagonally():
    """EAO (model.py:583-587): one pass per modality, one per combination, each over the packed tokens of its
    modalities with key padding only.  The plan lays the passes back to back: check the layout against the oracle's
    pass list, the bitmask form against the dense block-diagonal mask, the tile schedule against the dense mask, and
    the loss plan against the oracle's (= MCAPretrainingLoss with no_fusion)."""
    kw = C.get_model_config(C.tiny_config("cmu", fcl=True, bimodal=True, non_fusion_fcl=True, eao=True))
    p = _eao_plan(kw)
    passes = O.eao_passes(kw)
    assert p.passes == passes and p.R == len(passes) == 10
    lengths = [e["max_tokens"] for e in kw["encoder_configs"].values()]
    pid, src = [], []                                    # per stacked token: its pass and its source token
    for pi, members in enumerate(passes):
        for m in members:
            pid += [pi] * lengths[m]
            src += list(range(sum(lengths[:m]), sum(lengths[:m]) + lengths[m]))
    pid, src = np.asarray(pid), np.asarray(src)
    assert p.N == len(pid) and np.array_equal(p.tok_pass, pid)
    assert np.array_equal(np.asarray(p.pass_start), np.concatenate([[0], np.cumsum(np.bincount(pid))]))
    assert np.array_equal(src[:p.n_tok], np.arange(p.n_tok))               # the single passes ARE the encoder outputs
    got = np.arange(p.N)
    for dst, s, L in p.replicas:
        got[dst:dst + L] = np.arange(s, s + L)
    assert np.array_equal(got, src)                                         # every later block replicates a modality
    assert [sum(lengths[:m]) for m in p.mask_src] == [int(src[o]) for o in p.block_offsets]
    dense = pid[:, None] != pid[None, :]
    assert np.array_equal(p.attn_mask, dense)
    allowed = ((p.rowbits[:, None] >> p.keygrp[None, :].astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(allowed, ~dense)
    cover = np.zeros((p.N, p.N), dtype=bool)
    for qs, ql, off, cnt in p.q_tiles:
        for ki, flag in p.kt_list[off:off + cnt]:
            ks, kl = p.tiles[ki]
            assert flag == 0 and not dense[qs:qs + ql, ks:ks + kl].any()    # block structure: every listed tile is full
            cover[qs:qs + ql, ks:ks + kl] = True
    assert np.array_equal(cover, ~dense) and p.allowed_pairs == int((~dense).sum())
    plan, row = O.loss_plan(kw)
    assert [x["name"] for x in plan] == p.loss_names and len(plan) == 26
    for x, y in zip(plan, p.loss_plan):
        assert (x["a"], x["b"]) == (y["a"], y["b"])
        assert sum(1 << i for i in x["all"]) == y["all"] and sum(1 << i for i in x["any"]) == y["any"]
    assert [k for k, _ in p.output_rows] == list(kw["encoder_configs"].keys()) + list(p.combos)


def test_eao_plan_full_size_and_limits():
    """CMU_config1_EAO (configs/CMU_config1_EAO.yaml): 4 + 6 passes, 9800 stacked tokens per sample, 16 key groups."""
    cfg = C.named_config("CMU_config1")
    enc = cfg["encoder_configs"]
    from mca_paper_b200.plan import EAOPlan
    p = EAOPlan(enc, [2], True, False, True, True, True)
    assert (p.N, p.R, p.n_groups, p.n_tok) == (9800, 10, 16, 2450)
    assert p.allowed_pairs == sum(int(n) ** 2 for n in np.diff(p.pass_start))
    with pytest.raises(NotImplementedError):
        EAOPlan(enc, [2], True, False, False, True, True)          # a fusion row EAO never pools (model.py:190)
    assert EAOPlan(enc, [4, 3, 2], True, False, True, True, True).n_groups == 32     # 4 + 4 + 12 + 12 blocks: the limit
    enc5 = dict(enc, extra={"type": "EmbeddedSequenceEncoder", "input_size": 16, "max_tokens": 40})
    with pytest.raises(AssertionError):
        EAOPlan(enc5, [3, 2], True, False, True, True, True)       # 5 + 30 + 20 blocks > 32 key groups


def test_eao_model_schema_and_cpu_refusal():
    from mca_paper_b200.model import EAO
    kw = C.get_model_config(C.tiny_config("cmu", fcl=True, bimodal=True, non_fusion_fcl=True, eao=True))
    m = EAO(**kw)
    keys = list(m.state_dict().keys())
    assert "return_tokens" not in keys and "fusion_tokens" not in keys and not any(k.startswith("attn_pool") for k in keys)
    assert "token_types" in keys and m.state_dict()["token_types"].shape == (285,)
    with pytest.raises(NotImplementedError):
        EAO(**dict(kw, mean_pool=False))
    batch = S.make_batch(C.tiny_config("cmu", eao=True), seed=1, variant="full")
    with pytest.raises(_lib.MCAKernelError):
        m(batch)


def _random_geometry_strategy():
    from hypothesis import strategies as st
    return st.tuples(
        st.lists(st.integers(min_value=1, max_value=300), min_size=2, max_size=4),          # modality lengths
        st.sampled_from([[2], [3, 2], [4, 3, 2], [2, 3]]),                                  # fusion_combos powers
        st.integers(min_value=1, max_value=9),                                              # fusion tokens per combination
        st.sampled_from(["mca_fcl", "mca", "mma", "bimodal_fcl", "eao"]))


def test_plans_match_oracle_tables_on_random_geometries():
    """Property test over random modality counts / lengths (tiles of 128 that never straddle a block, ragged tails),
    combination sets and model families: the plan's bitmask form, dense masks, tile schedule and loss plan agree with
    the oracle's dense restatement of model.py:383-446 / 151-220 (EAO: the block-diagonal pass layout)."""
    from hypothesis import HealthCheck, given, settings

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(_random_geometry_strategy())
    def check(geo):
        lengths, powers, nsub, family = geo
        n = len(lengths)
        powers = [k for k in powers if k <= n] or [2]
        enc = {f"m{i}": {"type": "EmbeddedSequenceEncoder", "input_size": 8, "max_tokens": L} for i, L in enumerate(lengths)}
        n_combos = len(O.modality_combos(n, powers))
        kw = dict(encoder_configs=enc, num_fusion_tokens=nsub * n_combos, fusion_combos=powers, fcl=family in ("mca_fcl", "bimodal_fcl", "eao"),
                  zorro=family == "mma", no_fusion=family == "eao", bimodal_contrastive=family in ("bimodal_fcl", "eao"),
                  non_fusion_fcl=family in ("bimodal_fcl", "eao"), eao=family == "eao", batch_size=2, depth=1, heads=8)
        if family == "eao":
            if n + sum(len(c) for c in O.modality_combos(n, powers)) > 32:
                return
            p = _eao_plan(kw)
            pid = np.concatenate([np.full(sum(lengths[m] for m in members), pi) for pi, members in enumerate(O.eao_passes(kw))])
            dense = pid[:, None] != pid[None, :]
        else:
            p = make_plan(kw)
            t = O.static_tables(kw)
            dense = t["attn_mask"].numpy()
            assert np.array_equal(p.pool_mask, t["pool_mask"].numpy()) and np.array_equal(p.token_types, t["token_types"].numpy())
        assert np.array_equal(p.attn_mask, dense)
        rec = ~(((p.rowbits[:, None] >> p.keygrp[None, :].astype(np.uint32)) & 1).astype(bool))
        assert np.array_equal(rec, dense)
        covered = np.zeros_like(dense)
        for qs, ql, off, cnt in p.q_tiles:
            for ki, flag in p.kt_list[off:off + cnt]:
                ks, kl = p.tiles[ki]
                blk = ~dense[qs:qs + ql, ks:ks + kl]
                assert blk.any() and (flag == 0) == bool(blk.all())
                covered[qs:qs + ql, ks:ks + kl] = True
        assert not (~dense & ~covered).any()
        assert p.tiles[:, 1].max() <= 128 and (p.tiles[:, 0] + p.tiles[:, 1])[-1] == p.N
        fwd = {(qi, int(ki)) for qi, (_, _, off, cnt) in enumerate(p.q_tiles) for ki, _ in p.kt_list[off:off + cnt]}
        bwd = {(int(qi), ki) for ki, (_, _, off, cnt) in enumerate(p.k_tiles_q) for qi, _ in p.qt_list[off:off + cnt]}
        assert fwd == bwd
        plan, _ = O.loss_plan(kw)
        assert [x["name"] for x in plan] == p.loss_names
        for x, y in zip(plan, p.loss_plan):
            assert (x["a"], x["b"]) == (y["a"], y["b"])
            assert sum(1 << i for i in x["all"]) == y["all"] and sum(1 << i for i in x["any"]) == y["any"]

    check()


def test_ctypes_signatures_match_the_header():
    """Every entry point's ctypes argument list (mca_paper_b200/ops.py) against its declaration in include/mca_b200.h:
    same arity, and the same class per argument (pointer / int / long long / float / unsigned long long) — a drifted
    binding would pass garbage to a kernel instead of failing."""
    header = open(os.path.join(ROOT, "include", "mca_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    decls = dict(re.findall(r"^int (mca_\w+)\((.*?)\);", header, flags=re.M | re.S))
    assert set(decls) == set(ops._SIGS)

    def cls(param: str):
        p = " ".join(param.split())
        if p == "void":
            return None
        if "*" in p:
            return ctypes.c_void_p
        base = " ".join(p.split(" ")[:-1]).replace("const ", "")
        return {"int": ctypes.c_int32, "long long": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
                "unsigned long long": ctypes.c_uint64, "unsigned int": ctypes.c_uint32}[base]

    for name, params in decls.items():
        want = [c for c in (cls(x) for x in params.split(",")) if c is not None]
        got = list(ops._SIGS[name])
        assert len(want) == len(got), (name, len(want), len(got))
        for i, (w, g) in enumerate(zip(want, got)):
            is_f = lambda c: c in (ctypes.c_float, ctypes.c_double)
            same = ctypes.sizeof(w) == ctypes.sizeof(g) and (w is ctypes.c_void_p) == (g is ctypes.c_void_p) and is_f(w) == is_f(g)
            assert same, (name, i, w, g)


@pytest.mark.skipif(not os.path.isdir("/root/reference/configs"), reason="reference YAMLs only exist in the build container")
@pytest.mark.parametrize("name", ["CMU_config1", "CMU_config1_z", "CMU_config1_d40", "TCGA_config1", "CMU_config1_EAO",
                                  "CMU_config1_z_12i"])
def test_named_configs_equal_the_reference_yaml(name):
    """The in-repo copies of the benchmark configurations give the model the same kwargs as the reference's YAML run
    through its own defaults (utils/config.py:9-57 defaults, :96-117 get_model_config)."""
    mine = C.get_model_config(C.named_config(name))
    theirs = C.get_model_config(C.training_config(f"/root/reference/configs/{name}.yaml"))
    assert set(mine) == set(theirs)
    for k in mine:
        a, b = mine[k], theirs[k]
        if k == "encoder_configs":
            assert list(a.keys()) == list(b.keys())
            for m in a:
                assert dict(a[m]) == dict(b[m]), (m, a[m], b[m])
        else:
            assert (list(a) if isinstance(a, (list, tuple)) else a) == (list(b) if isinstance(b, (list, tuple)) else b), k


def test_mask_plan_factorises_dense_masks_into_key_groups():
    """plan.MaskPlan (free-standing Attention calls): key groups / row bits reproduce the dense mask bit for bit, equal
    the model's own tables on the reference masks, and the tile schedule covers exactly the tiles holding allowed pairs."""
    import numpy as np
    from mca_paper_b200.plan import MaskPlan, StaticPlan, TILE

    for kind, kw in (("cmu", dict(fcl=True)), ("cmu", dict(zorro=True, fcl=False)), ("tcga", dict(fcl=True))):
        k = C.get_model_config(C.tiny_config(kind, **kw))
        sp = StaticPlan(k["encoder_configs"], k["num_fusion_tokens"], list(k["fusion_combos"]), k["fcl"],
                        k.get("zorro", False), k.get("no_fusion", False), False, False)
        mp = MaskPlan(sp.attn_mask, sp.N)
        assert np.array_equal(mp.attn_mask, sp.attn_mask)
        assert np.array_equal(mp.keygrp, sp.keygrp) and np.array_equal(mp.rowbits, sp.rowbits)
    rng = np.random.default_rng(3)
    for n, ng in ((1, 1), (129, 2), (700, 6), (1000, 32)):
        grp = np.sort(rng.integers(0, ng, size=n))
        vis = rng.random((ng, ng)) < 0.4
        mask = ~vis[grp][:, grp]
        mp = MaskPlan(mask, n)
        assert np.array_equal(mp.attn_mask, mask) and mp.n_groups <= ng + 1
        covered = np.zeros((n, n), dtype=bool)
        assert int(mp.tiles[:, 1].sum()) == n and (mp.tiles[:, 1] <= TILE).all()
        for qs, ql, off, cnt in mp.q_tiles:
            for t, flag in mp.kt_list[off:off + cnt]:
                ks, kl = mp.tiles[t]
                blk = ~mask[qs:qs + ql, ks:ks + kl]
                assert blk.any() and (flag == 0) == bool(blk.all())
                covered[qs:qs + ql, ks:ks + kl] = True
        assert not (~mask & ~covered).any()
    with pytest.raises(NotImplementedError):
        MaskPlan(rng.random((64, 64)) < 0.5, 64)
    assert MaskPlan(None, 300).n_groups == 1
