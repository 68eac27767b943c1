#!/bin/bash
# 1-GPU session: all GPU tests, smoke, bench (engine + drop-in), low-rank kernel times after the cluster
# orthonormalisation, sweep of the codecs being worked on.  Output: gpurun_out/<tag>_*.
TAG=${1:-r2c}
WHAT=${2:-"tests smoke bench dropin lowrank sweep"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
line() { python - "$1" <<'PY'
import json, sys
try:
    txt = open(sys.argv[1]).read()
    d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("  %.3f ms/step  %.0f GB/s  parity_ok=%s  rel_l2=%s  mode=%s" % (d["ms_per_step"], d["value"], d.get("parity_ok"),
          (d.get("fidelity") or {}).get("rel_l2"), d["config"].get("launch_mode")))
    for k in (d.get("roofline") or {}).get("kernels", []):
        print("    %-22s %7.2f us  frac %.3f" % (k["kernel"], k["avg_launch_us"], k["frac"]))
    for key in ("e2e", "gpu_reference", "cpu_baseline"):
        if d.get(key): print("   ", key, {k: v for k, v in d[key].items() if k in ("value", "ms_per_step", "kind", "error")})
except Exception as e:
    print("  no line:", e)
PY
}
has tests && { echo "== tests"; (time timeout 1500 python -m pytest tests -m gpu -q -rfs --durations=8) > $OUT/${TAG}_tests.log 2>&1 ; tail -14 $OUT/${TAG}_tests.log; }
has smoke && { echo "== smoke"; timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -1 $OUT/${TAG}_smoke.log; }
has bench && { echo "== bench"; timeout 500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; line $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err; }
has dropin && { echo "== bench --api dropin"; timeout 300 python bench.py --api dropin --no-e2e --no-cpu-baseline --no-gpu-reference > $OUT/${TAG}_bench_dropin.json 2> $OUT/${TAG}_bench_dropin.err ; line $OUT/${TAG}_bench_dropin.json; tail -2 $OUT/${TAG}_bench_dropin.err; }
has config1 && { echo "== BASELINE configs[0]: INT4 round trip"; timeout 300 python bench.py --workload config1_int4_roundtrip > $OUT/${TAG}_bench_config1.json 2> $OUT/${TAG}_bench_config1.err ; head -c 1800 $OUT/${TAG}_bench_config1.json; echo; tail -2 $OUT/${TAG}_bench_config1.err; }
has lrq && { for w in flux1024_patch_parallel cogvideox5b_ring; do echo "== lowrankq32 $w (N=1)"; timeout 300 python bench.py --codec lowrankq32 --workload $w --no-e2e --no-cpu-baseline --no-gpu-reference --steps 5 > $OUT/${TAG}_bench_lrq_$w.json 2> $OUT/${TAG}_bench_lrq_$w.err ; line $OUT/${TAG}_bench_lrq_$w.json; tail -2 $OUT/${TAG}_bench_lrq_$w.err; done; }
has refarm && { echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err ; head -c 500 $OUT/${TAG}_bench_ref.json; echo; }
if has lowrank; then
  echo "== low-rank kernel times"
  for args in "lowrank --rank 32" "lowrank --rank 8" "lowrank --rank 32 --shape 576x3072" "lowrank --rank 64"; do
    echo "-- $args" >> $OUT/${TAG}_kernel_times.md; timeout 120 python tools/kernel_times.py $args 2>&1 | grep -v Warn >> $OUT/${TAG}_kernel_times.md
  done
  echo "-- CF_LR_ORTH=legacy lowrank --rank 32" >> $OUT/${TAG}_kernel_times.md; CF_LR_ORTH=legacy timeout 120 python tools/kernel_times.py lowrank --rank 32 2>&1 | grep -v Warn >> $OUT/${TAG}_kernel_times.md
  cat $OUT/${TAG}_kernel_times.md | grep -v "^$" | head -80
fi
has gemmsmall && { echo "-- CF_LR_GEMM=small lowrank --rank 32" >> $OUT/${TAG}_kernel_times.md; CF_LR_GEMM=small timeout 120 python tools/kernel_times.py lowrank --rank 32 2>&1 | grep -v Warn | tee -a $OUT/${TAG}_kernel_times.md | head -8; }
has codecs && { for args in "codec --codec int4" "codec --codec sparse"; do echo "-- $args" >> $OUT/${TAG}_kernel_times.md; timeout 120 python tools/kernel_times.py $args 2>&1 | grep -v Warn | tee -a $OUT/${TAG}_kernel_times.md; done; }
if has sweep; then
  echo "== sweep"
  timeout 300 python sweep.py --sizes-mb 27 --shapes 4608x3072,576x3072 --ops ${SWEEP_OPS:-lowrank} --ranks 8,32,64 --reps 5 \
    --out $OUT/${TAG}_sweep.jsonl --md $OUT/${TAG}_sweep.md > $OUT/${TAG}_sweep.log 2>&1
  cat $OUT/${TAG}_sweep.md | head -40
fi
ls $OUT | grep ${TAG}_ | tail -20
