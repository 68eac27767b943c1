#!/bin/bash
# ncu session (1 GPU): source-level capture of the low-rank orthonormalisation, traffic / stall picture of the
# fused-put kernels with W = 2 / 8 virtual ranks, the INT4 pipeline.  The .ncu-rep files are summarised ON the box
# (tools/ncu_summary.py, tools/ncu_hot_lines.py) and deleted: gpurun brings back at most 64 MiB.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
summ() {  # summ <name> [traffic suffix]
  python tools/ncu_summary.py full $OUT/${TAG}_$1.ncu-rep $OUT/${TAG}_ncu_$1.md ${2:+--traffic $2} 2>&1 | tail -2
  python tools/ncu_hot_lines.py $OUT/${TAG}_$1.ncu-rep $OUT/${TAG}_hot_$1.md 40 2>&1 | tail -2
  cp profiles/traffic.json $OUT/${TAG}_traffic.json 2>/dev/null
  rm -f $OUT/${TAG}_$1.ncu-rep
}
echo "== ncu: k_lr_orth"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lr_orth' -s 3 -c 3 -f -o $OUT/${TAG}_orth \
  python tools/kernel_times.py lowrank --rank 32 --reps 1 > $OUT/${TAG}_log_orth.log 2>&1; tail -1 $OUT/${TAG}_log_orth.log
summ orth
for W in 2 8; do
  echo "== ncu: fused put with $W virtual ranks"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_delta_stats|k_finalize|k_apply|k_publish' -s $((W * 3 * 4)) -c $((W * 4 + 4)) \
    -f -o $OUT/${TAG}_vr$W python tools/virtual_ranks_step.py --world $W --layers 3 --steps 2 > $OUT/${TAG}_log_vr$W.log 2>&1; tail -1 $OUT/${TAG}_log_vr$W.log
  summ vr$W "binary|n$((4608 / W))|w$W"
done
echo "== ncu: INT4 pipeline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_minmax|k_int4' -s 6 -c 4 -f -o $OUT/${TAG}_int4 \
  python tools/kernel_times.py codec --codec int4 --reps 1 > $OUT/${TAG}_log_int4.log 2>&1; tail -1 $OUT/${TAG}_log_int4.log
summ int4
du -sh $OUT; ls -la $OUT | grep ${TAG}_
