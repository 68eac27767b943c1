#!/bin/bash
# Round-2 first 1-GPU session: all GPU tests (virtual-rank oracle tests included), smoke, the headline bench with the
# parity leg, the two-chain A/B, per-kernel times + ncu of the low-rank / INT4 / top-k kernels (the round's perf
# targets), the reference's own unit tests against this library, the CPU arm.  Output: gpurun_out/<tag>_*.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests" ; (time timeout 1200 python -m pytest tests -m gpu -q -rfs --durations=15) > $OUT/${TAG}_tests.log 2>&1 ; tail -25 $OUT/${TAG}_tests.log
echo "== smoke" ; timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -1 $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "fidelity", "ranks_identical", "parity_ok")}, d["oracle_parity"], d["e2e"]["value"], d["roofline"]["frac"])
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 $OUT/${TAG}_bench.err
echo "== bench --overlap" ; timeout 300 python bench.py --steps 10 --no-e2e --no-cpu-baseline --overlap > $OUT/${TAG}_bench_overlap.json 2>> $OUT/${TAG}_bench.err ; head -c 300 $OUT/${TAG}_bench_overlap.json; echo
echo "== bench int2" ; timeout 300 python bench.py --codec int2 --steps 10 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_int2.json 2>> $OUT/${TAG}_bench.err ; head -c 300 $OUT/${TAG}_bench_int2.json; echo
echo "== kernel times (CUPTI)"
for args in "lowrank --rank 32" "lowrank --rank 8" "lowrank --rank 32 --shape 576x3072" "lowrank --rank 64" "codec --codec int4" "codec --codec sparse" ; do
  echo "-- $args" >> $OUT/${TAG}_kernel_times.md; timeout 120 python tools/kernel_times.py $args >> $OUT/${TAG}_kernel_times.md 2>&1
done
cat $OUT/${TAG}_kernel_times.md | head -120
echo "== ncu full: low-rank kernels"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lr_' -c 40 -f -o $OUT/${TAG}_lowrank_full \
  python tools/kernel_times.py lowrank --rank 32 --reps 1 > $OUT/${TAG}_ncu_lowrank.log 2>&1
echo "== ncu full: INT4 / top-k kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_minmax|k_int4|k_topk' -c 16 -f -o $OUT/${TAG}_int4_full \
  python tools/kernel_times.py codec --codec int4 --reps 1 > $OUT/${TAG}_ncu_int4.log 2>&1
echo "== the reference's own tests against this library"
[ -d baseline/_ref/tests/compact ] && (timeout 900 python tools/run_reference_tests.py > $OUT/${TAG}_reference_tests.log 2>&1; tail -15 $OUT/${TAG}_reference_tests.log)
echo "== reference arm" ; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err ; head -c 400 $OUT/${TAG}_bench_ref.json; echo
echo "== sweep (this round's starting point)"
timeout 300 python sweep.py --sizes-mb 27 --shapes 4608x3072,576x3072 --ops int4,int8,topk,lowrank --reps 5 \
  --out $OUT/${TAG}_sweep.jsonl --md $OUT/${TAG}_sweep.md > $OUT/${TAG}_sweep.log 2>&1
tail -5 $OUT/${TAG}_sweep.log
ls -la $OUT | tail -30
