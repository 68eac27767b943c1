#!/bin/bash
# ncu source-level capture of k_lr_orth only (summarised on the box; the report is deleted).
TAG=${1:-r2v}; RANK=${2:-32}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lr_orth' -s 3 -c 3 -f -o $OUT/${TAG}_orth \
  python tools/kernel_times.py lowrank --rank $RANK --reps 1 > $OUT/${TAG}_log_orth.log 2>&1; tail -1 $OUT/${TAG}_log_orth.log
python tools/ncu_summary.py full $OUT/${TAG}_orth.ncu-rep $OUT/${TAG}_ncu_orth.md 2>&1 | tail -2
python tools/ncu_hot_lines.py $OUT/${TAG}_orth.ncu-rep $OUT/${TAG}_hot_orth.md 60 2>&1 | tail -2
rm -f $OUT/${TAG}_orth.ncu-rep
head -80 $OUT/${TAG}_hot_orth.md
