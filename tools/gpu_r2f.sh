#!/bin/bash
# ncu session (1 GPU): source-level capture of the low-rank orthonormalisation, traffic / stall picture of the
# fused-put kernels with W = 2 / 8 virtual ranks, the INT4 pipeline.  Output: gpurun_out/<tag>_*.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
echo "== ncu: k_lr_orth"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lr_orth' -s 3 -c 3 -f -o $OUT/${TAG}_orth \
  python tools/kernel_times.py lowrank --rank 32 --reps 1 > $OUT/${TAG}_ncu_orth.log 2>&1; tail -2 $OUT/${TAG}_ncu_orth.log
for W in 2 8; do
  echo "== ncu: fused put with $W virtual ranks"
  timeout 400 ncu --set full --clock-control none -k regex:'k_delta_stats|k_finalize|k_apply|k_publish' -s $((W * 3 * 4)) -c $((W * 4 + 4)) \
    -f -o $OUT/${TAG}_vr$W python tools/virtual_ranks_step.py --world $W --layers 3 --steps 2 > $OUT/${TAG}_ncu_vr$W.log 2>&1; tail -2 $OUT/${TAG}_ncu_vr$W.log
done
echo "== ncu: INT4 pipeline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_minmax|k_int4' -s 6 -c 4 -f -o $OUT/${TAG}_int4 \
  python tools/kernel_times.py codec --codec int4 --reps 1 > $OUT/${TAG}_ncu_int4.log 2>&1; tail -2 $OUT/${TAG}_ncu_int4.log
ls -la $OUT | grep ${TAG}_
