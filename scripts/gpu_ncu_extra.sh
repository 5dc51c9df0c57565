#!/bin/bash
# ncu --set full of the kernels outside the default bf16 step: the fp32-parity forward (exact.cu) and the linear probe.
OUT=gpurun_out; mkdir -p $OUT
STEP="python bench.py --profile-step --no-graphs --warmup 3 --no-cpu-baseline"
MCA_PRECISION=fp32 $STEP > $OUT/plain_x.log 2>&1 &&
MCA_PRECISION=fp32 ncu --set full --clock-control none --profile-from-start off -k "regex:attn_fwd_f32_kernel|split_kernel|geglu_f32_kernel|pool_fwd_kernel|pack_weights_split" -c 14 -f -o $OUT/prof_x $STEP > $OUT/ncu_x.log 2>&1
echo "ncu exact rc=$?"
ncu -i $OUT/prof_x.ncu-rep --page raw --csv > $OUT/prof_x_raw.csv 2>/dev/null; rm -f $OUT/prof_x.ncu-rep
cat > /tmp/probe_run.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from mca_paper_b200.linear_probe import FineTuneDataset, LinearProbe
g = torch.Generator().manual_seed(0)
n = 16384
e = {"fusion": torch.randn(n, 512, generator=g)}
s = torch.randn(n, 7, generator=g)
torch.manual_seed(42)
p = LinearProbe(FineTuneDataset(e, s, index=0), FineTuneDataset(e, s, index=0), batch_size=1024, epochs=4, lr=1e-3, num_warmup_steps=4)
p.fit(); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); p.train_epoch(); ev1.record(); torch.cuda.synchronize()
print(f"probe epoch (16384 rows, batch 1024: 16 steps + eval + PCC, host order included): {ev0.elapsed_time(ev1):.3f} ms")
PY
python /tmp/probe_run.py 2>&1 | grep -v Warn | tee $OUT/probe_time.log
ncu --set full --clock-control none -k "regex:probe_epoch_kernel" -c 2 -f -o $OUT/prof_probe python /tmp/probe_run.py > $OUT/ncu_probe.log 2>&1; echo "ncu probe rc=$?"
ncu -i $OUT/prof_probe.ncu-rep --page raw --csv > $OUT/prof_probe_raw.csv 2>/dev/null; rm -f $OUT/prof_probe.ncu-rep
