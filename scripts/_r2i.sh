OUT=gpurun_out; mkdir -p $OUT
python scripts/gpu_determinism.py 1 dropout_ragged 2>&1 | grep -v Warning | tee $OUT/determinism_r2i_l1.log
python scripts/gpu_determinism.py 1 full 2>&1 | grep -v Warning | tee $OUT/determinism_r2i_l1_full.log
