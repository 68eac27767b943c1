#!/bin/bash
# Round-end validation on one GPU: all GPU tests, smoke, the default bench line (both arms), configs[0], and the ncu
# launch list of the bench command (shares only: per-launch times under ncu are cold-cache and serialised).
TAG=${1:-r2_final}
OUT=gpurun_out; mkdir -p $OUT
(time timeout 1200 python -m pytest tests -m gpu -q -rfs) > $OUT/${TAG}_tests.log 2>&1; tail -4 $OUT/${TAG}_tests.log
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; head -c 600 $OUT/${TAG}_bench.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err; head -c 400 $OUT/${TAG}_bench_reference_arm.json; echo
timeout 300 python bench.py --workload config1_int4_roundtrip > $OUT/${TAG}_bench_config1.json 2>> $OUT/${TAG}_bench.err; head -c 300 $OUT/${TAG}_bench_config1.json; echo
for r in 32 8 64; do echo "-- lowrank --rank $r" >> $OUT/${TAG}_kernel_times_lowrank.md; timeout 120 python tools/kernel_times.py lowrank --rank $r 2>&1 | grep -v -i warn >> $OUT/${TAG}_kernel_times_lowrank.md; done
grep "sum of" $OUT/${TAG}_kernel_times_lowrank.md
timeout 200 python bench.py --codec lowrankq32 --no-e2e --no-cpu-baseline --no-gpu-reference --steps 5 > $OUT/${TAG}_bench_lrq.json 2>> $OUT/${TAG}_bench.err; head -c 250 $OUT/${TAG}_bench_lrq.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 800 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --no-parity > $OUT/${TAG}_ncu_list.log 2>&1
python tools/ncu_summary.py list $OUT/${TAG}_launches.csv $OUT/${TAG}_launches.md 2>&1 | tail -2; head -14 $OUT/${TAG}_launches.md
rm -f $OUT/${TAG}_launches.csv
