"""GPU bring-up check: fused CUDA path vs the CPU oracle, stage by stage (tiny configs, a few seconds of CPU)."""
import sys, time
import torch
sys.path.insert(0, ".")
from mca_paper_b200 import config as C, synthetic as S
from mca_paper_b200.model import MCA
from oracle import mca_oracle as O

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def run(cfg, variant, tag, check_bwd=True, conditioned=False):
    kw = C.get_model_config(cfg)
    torch.manual_seed(0)
    model = MCA(**kw)
    if conditioned:  # small pooled embeddings and T = 1: forward rounding is no longer amplified by a sharp softmax
        with torch.no_grad():
            model.attn_pool.to_out.weight.mul_(0.05); model.return_tokens.mul_(0.02); model.loss.loss_fn.logit_scale.fill_(0.0)
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    batch = S.make_batch(cfg, seed=1, variant=variant)
    t = O.static_tables(kw)
    params = {k: v.clone().requires_grad_(True) for k, v in sd_cpu.items() if k in dict(model.named_parameters())}
    sd2 = dict(sd_cpu); sd2.update(params)
    ref = O.mca_forward(sd2, kw, batch, tables=t)
    # reference intermediates
    with torch.no_grad():
        pooled_ref, smask, xf_ref = O.trunk(sd_cpu, kw, batch, t)
    out = model(S.batch_to(batch, dev))
    torch.cuda.synchronize()
    eng = model.engine
    print(f"== {tag}: N={eng.N} R={eng.R} pairs={eng.plan.n_pairs}")
    # encoder output
    toks = []
    for name in t["names"]:
        x, m = O.encode_modality(sd_cpu, name, kw["encoder_configs"][name], batch[name]); toks.append(x)
    toks.append(sd_cpu["fusion_tokens"].unsqueeze(0).expand(eng.B, -1, -1))
    x0_ref = torch.cat(toks, 1).reshape(eng.M, 512)
    print("  x0 (encoders)      rel", rel(eng.ws["xa"][0], x0_ref))
    padding_ref = torch.cat([batch[n]["attention_mask"].bool() for n in t["names"]] + [torch.zeros(eng.B, eng.plan.F, dtype=torch.bool)], 1)
    print("  padding bit-exact  ", bool((eng.ws["padding"].cpu().bool() == padding_ref).all()))
    print("  final tokens (bf16 copy of LN) rel", rel(eng.ws["xf_16"], xf_ref.reshape(eng.M, 512)))
    print("  pooled             rel", rel(eng.ws["pooled"], pooled_ref))
    worst_emb = 0
    for k in ref:
        if k in ("losses", "modality_sample_mask"): continue
        if isinstance(k, str) and "loss" in k: print(f"  out[{k}] rel", rel(out[k], ref[k]))
        else: worst_emb = max(worst_emb, rel(out[k], ref[k]))
    print("  worst embedding rel", worst_emb)
    worst = 0
    for k, v in ref["losses"].items():
        a = out["losses"][k]
        if torch.isnan(v):
            assert torch.isnan(a), k
        else:
            worst = max(worst, abs(a.item() - v.item()) / abs(v.item()))
    print("  worst per-pair loss rel", worst, " loss", out["loss"].item(), "ref", ref["loss"].item())
    for k in ref["modality_sample_mask"]:
        assert torch.equal(out["modality_sample_mask"][k].cpu(), ref["modality_sample_mask"][k]), k
    if not check_bwd:
        return
    ref["loss"].backward()
    out["loss"].backward()
    torch.cuda.synchronize()
    bad = []
    for k, p in model.named_parameters():
        g_ref = params[k].grad
        g = p.grad
        if g_ref is None or g_ref.abs().max() == 0:
            continue
        r = rel(g, g_ref)
        bad.append((r, k))
    bad.sort(reverse=True)
    print("  worst grads:", [(round(r, 4), k) for r, k in bad[:8]])
    print("  median grad rel", sorted(r for r, _ in bad)[len(bad) // 2])


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cmu"):
    run(C.tiny_config("cmu", fcl=True), "full", "tiny CMU MCA-fcl, full length")
if which in ("all", "cond"):
    run(C.tiny_config("cmu", fcl=True), "full", "tiny CMU MCA-fcl, full length, well-conditioned loss", conditioned=True)
    run(C.tiny_config("cmu", fcl=True), "dropout_ragged", "tiny CMU MCA-fcl, ragged, well-conditioned loss", conditioned=True)
    run(C.tiny_config("tcga", fcl=True, bimodal=True, non_fusion_fcl=True), "tcga", "tiny TCGA, well-conditioned", conditioned=True)
if which in ("all", "ragged"):
    run(C.tiny_config("cmu", fcl=True), "dropout_ragged", "tiny CMU MCA-fcl, absent + ragged")
if which in ("all", "mma"):
    run(C.tiny_config("cmu", zorro=True, fcl=False), "dropout_full", "tiny CMU MMA, absent modalities")
if which in ("all", "tcga"):
    run(C.tiny_config("tcga", fcl=True, bimodal=True, non_fusion_fcl=True), "tcga", "tiny TCGA, scattered pads")
print("DONE")
