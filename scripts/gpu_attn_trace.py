"""Debug timelines of the attention kernels (library built with -DMCA_TRACE, scripts/build_trace_lib.sh):
one CTA's per-role clock64 stamps (forward and backward) and the (start, end, SM) of every CTA of a launch."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S, _lib
from mca_paper_b200.model import MCA

cfg = C.named_config("CMU_config1")
model = MCA(**C.get_model_config(cfg)).to("cuda")
eng = model.engine
eng.ensure_flat()
eng.build_offsets(S.batch_to(S.make_batch(cfg, seed=1, variant="full"), "cuda"))
ws = eng.ws
ws["qkv"][0].copy_((torch.randn(eng.M, 1536, device="cuda") * 0.5).bfloat16())
ws["dattn"].copy_(torch.randn(eng.M, 512, device="cuda").bfloat16())
lib = _lib.lib()


def cta_summary(tag, a):
    """a [n, 4] = (t_start ns, t_end ns, smid, aux)."""
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    st, en, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2]
    dur = en - st
    print(f"[{tag}] CTAs {len(a)}  span {en.max() / 1e3:.1f} us  mean CTA {dur.mean() / 1e3:.2f} us  min {dur.min() / 1e3:.2f}  max {dur.max() / 1e3:.2f}")
    busy = np.zeros(148 + 64)
    last = np.zeros(148 + 64)
    for s in np.unique(sm):
        m = sm == s
        busy[int(s)] = dur[m].sum()
        last[int(s)] = en[m].max()
    used = busy[busy > 0]
    print(f"[{tag}] per-SM sum of CTA durations: mean {used.mean() / 1e3:.1f} us  min {used.min() / 1e3:.1f}  max {used.max() / 1e3:.1f};"
          f" last-CTA end per SM: min {last[last > 0].min() / 1e3:.1f}  max {last.max() / 1e3:.1f}")
    # duration by aux (number of tiles) when provided
    aux = a[:, 3]
    for v in np.unique(aux):
        m = aux == v
        print(f"[{tag}]   aux={int(v):3d}: n={m.sum():4d} mean dur {dur[m].mean() / 1e3:7.2f} us  start range {st[m].min() / 1e3:6.1f}..{st[m].max() / 1e3:6.1f}")
    # timeline of one SM
    s0 = int(sm[len(sm) // 2])
    m = sm == s0
    order = np.argsort(st[m])
    print(f"[{tag}] SM {s0}: " + " ".join(f"[{st[m][i] / 1e3:.1f}-{en[m][i] / 1e3:.1f}|{int(aux[m][i])}]" for i in order))


# ------------------------------------------------------------------ forward
for _ in range(3):
    eng.attention_fwd(ws["qkv"][0], ws["ao"][0], ws["lse"][0])
torch.cuda.synchronize()
NR = 10
n_tr, n_cta = NR * 24 * 8 + 8, 4096 * 4
tr = (ctypes.c_longlong * n_tr)()
ct = (ctypes.c_longlong * n_cta)()
rc = lib.mca_debug_read_trace_fwd(tr, n_tr, ct, n_cta)
a = np.array(tr[:], dtype=np.int64)
t0 = a[NR * 24 * 8]
print(f"FWD traced CTA: start=0 end={int(a[NR * 24 * 8 + 1] - t0)} cycles")
ev = a[:NR * 24 * 8].reshape(NR, 24, 8)
names = {i: f"SW{i} " for i in range(8)}
names.update({8: "MMA ", 9: "TMA "})
print("SWn: top s_full ld_done max_done exp_done pv_done/rescale p_full | MMA: top k_full s_empty S_issued v_full p_full PV_issued | TMA: top k_empty v_empty")
for t in range(24):
    for role in (8, 0, 1, 2, 3, 4, 5, 6, 7, 9):
        row = ev[role, t]
        print(f"t={t:2d} {names[role]}: " + " ".join(f"{int(x - t0):7d}" if x else "      -" for x in row))
c = np.array(ct[:], dtype=np.int64).reshape(4096, 4)   # per work item: (start, end, smid, n_it)
cta_summary("fwd", c)

# ------------------------------------------------------------------ backward
for _ in range(3):
    eng.attention_bwd(0)
torch.cuda.synchronize()
n = 4 * 16 * 16 + 8
buf = (ctypes.c_longlong * n)()
rc = lib.mca_debug_read_trace(buf, n)
a = np.array(buf[:], dtype=np.int64)
g = a[4 * 16 * 16:]
t0 = g[0]
print("BWD global: setup_done=0 kv_full=%d q0_full=%d epi_wg0=%d epi_wg1=%d end_wg0=%d end_wg1=%d cta_end=%d" % tuple(int(x - t0) for x in g[1:8]))
ev = a[:4 * 16 * 16].reshape(4, 16, 16)[:, :, :16]
names = {0: "WG0", 1: "WG1", 2: "MMA", 3: "MMB"}
for t in range(13):
    for role in (2, 3, 0, 1):
        row = ev[role, t]
        print(f"t={t:2d} {names[role]}: " + " ".join(f"{int(x - t0):7d}" if x else "      -" for x in row))
cb = (ctypes.c_longlong * n_cta)()
lib.mca_debug_read_cta_bwd(cb, n_cta)
cta_summary("bwd", np.array(cb[:], dtype=np.int64).reshape(4096, 4))
