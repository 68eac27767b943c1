#!/bin/bash
# Where does a hook call's time go at N ranks?  Host enqueue time of a step and host time inside the C calls.
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
run() { local name=$1; shift
  CF_BENCH_VERBOSE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --api dropin "$@" > $OUT/hookdiag_n${N}_$name.json 2> $OUT/hookdiag_n${N}_$name.err
  grep -E "timed region done|CF_PLAN_TIMING" $OUT/hookdiag_n${N}_$name.err | head -4
  python - <<PY
import json
d = json.loads([l for l in open("$OUT/hookdiag_n${N}_$name.json").read().splitlines() if l.startswith("{")][-1])
print("$name: %.3f ms/step host_enqueue=%s launches=%s" % (d["ms_per_step"], d["config"].get("host_enqueue_ms_per_step"), d["gpu_launches"]))
PY
}
WHAT=${2:-"plain stable engine_eager"}
has() { [[ " $WHAT " == *" $1 "* ]]; }
has plain && run plain
has plan_timing && CF_PLAN_TIMING=1 run plan_timing
has stable && CF_DROPIN_INPUTS_STABLE=1 run stable
has engine_eager && run engine_eager --api engine --no-graph
has engine_eager_unstable && CF_ENGINE_INPUTS_STABLE=0 run engine_eager_unstable --api engine --no-graph
