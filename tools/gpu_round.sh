#!/bin/bash
# One GPU-box session: parity tests, smoke, the headline bench (both arms), the ncu launch list of the bench
# command, one `ncu --set full` capture of the step's kernels, A/B runs of the launch options and the codec
# sweep.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r1i [quick]'
TAG=${1:-run}
MODE=${2:-full}
OUT=gpurun_out
export CF_EXPERIMENTAL=1  # arm the tests of opt-in features (two-chain step)
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests" ; (time timeout 900 python -m pytest tests -m gpu -q -rf) > $OUT/${TAG}_tests.log 2>&1 ; tail -12 $OUT/${TAG}_tests.log
echo "== smoke" ; timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -1 $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; cat $OUT/${TAG}_bench.json
echo "== bench int2" ; timeout 300 python bench.py --codec int2 --steps 10 --no-cpu-baseline > $OUT/${TAG}_bench_int2.json 2>> $OUT/${TAG}_bench.err
echo "== A/B: launch options (binary, no e2e / cpu)"
for v in "CF_L2_HINTS=0" "CF_PDL=0" "CF_L2_HINTS=0 CF_PDL=0"; do
  echo "-- $v" >> $OUT/${TAG}_ab.log
  env $v timeout 200 python bench.py --steps 10 --no-e2e --no-cpu-baseline >> $OUT/${TAG}_ab.log 2>&1
done
[ "$MODE" = quick ] && exit 0
echo "== reference arm" ; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err ; cat $OUT/${TAG}_bench_ref.json
echo "== ncu launch list (our kernels are all named k_*)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_list.log 2>&1
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_delta_stats|k_finalize|k_apply|k_int2' -s 12 -c 9 \
  -f -o $OUT/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 8 --no-graph > $OUT/${TAG}_ncu_full.log 2>&1
echo "== kernel times (CUPTI)"
for args in "lowrank --rank 32" "lowrank --rank 8" "codec --codec int4" "step --codec binary" "step --codec binary --overlap"; do
  echo "-- $args" >> $OUT/${TAG}_kernel_times.md; timeout 120 python tools/kernel_times.py $args >> $OUT/${TAG}_kernel_times.md 2>&1
done
echo "== A/B: two-chain step (opt-in)"
timeout 200 python bench.py --steps 10 --no-e2e --no-cpu-baseline --overlap > $OUT/${TAG}_bench_overlap.json 2>> $OUT/${TAG}_bench.err ; tail -c 600 $OUT/${TAG}_bench_overlap.json
echo "== the reference's own tests against this library (needs tools/stage_reference.sh run in the build container)"
[ -d baseline/_ref/tests/compact ] && (timeout 600 python tools/run_reference_tests.py -x > $OUT/${TAG}_reference_tests.log 2>&1; tail -3 $OUT/${TAG}_reference_tests.log)
echo "== sweep"
timeout 420 python sweep.py --sizes-mb 1,8,27,256,1024 --shapes 4608x3072,576x3072,4388x3072,8192x1152 --reps 5 \
  --out $OUT/${TAG}_sweep.jsonl --md $OUT/${TAG}_sweep.md > $OUT/${TAG}_sweep.log 2>&1
tail -5 $OUT/${TAG}_sweep.log
ls -la $OUT | tail -20
