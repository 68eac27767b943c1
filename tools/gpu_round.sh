#!/bin/bash
# One GPU-box session: parity tests, smoke, the headline bench (both arms), the ncu launch list of the bench
# command and one `ncu --set full` capture of the step's kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r1h'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests" ; (time timeout 600 python -m pytest tests -m gpu -x -q) > $OUT/${TAG}_tests.log 2>&1 ; tail -3 $OUT/${TAG}_tests.log
echo "== smoke" ; timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -1 $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; cat $OUT/${TAG}_bench.json
echo "== bench int2" ; timeout 300 python bench.py --codec int2 --steps 10 --no-cpu-baseline > $OUT/${TAG}_bench_int2.json 2>> $OUT/${TAG}_bench.err
echo "== reference arm" ; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err ; cat $OUT/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_list.log 2>&1
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_delta_stats|k_finalize|k_apply|k_int2' -s 12 -c 9 \
  -f -o $OUT/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 8 --no-graph > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
