#!/bin/bash
# One multi-GPU box session (N = 2, 4 or 8): the multi-rank parity tests, then bench.py at N ranks in the variants
# named in $3 (default: all).  Everything lands in gpurun_out/<tag>_*.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_multi2.sh r2b 2'
#   gpurun --gpus 8 --timeout 900  -- 'bash tools/gpu_multi2.sh r2d 8 "tests serial overlap dropin raw ring"'
TAG=${1:-multi}
N=${2:-2}
WHAT=${3:-"tests serial overlap publish1 nobulk dropin dropin_nccl nccl int2 int2_nobulk raw raw_ring raw_async ring ring_raw pixart sd3"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
  echo "== multi-rank tests"
  (time timeout 900 python -m pytest tests -m gpu -q -rfs -k "two_gpu") > $OUT/${TAG}_tests_n${N}.log 2>&1 ; tail -8 $OUT/${TAG}_tests_n${N}.log
fi
PORT=29500
run() {  # run <name> <timeout> -- bench args   (env vars may be set by the caller)
  local name=$1 tmo=$2; shift 2
  PORT=$((PORT + 1))
  CF_BENCH_VERBOSE=1 timeout $tmo python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $PORT bench.py --gpus $N --hang-dump $((tmo - 10)) "$@" > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_${name}.json").read().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("$name: %.3f ms/step  %.0f GB/s  parity_ok=%s  rel_l2=%s  identical=%s  transport=%s  schedule=%s" % (
        d["ms_per_step"], d["value"], d.get("parity_ok"), (d.get("fidelity") or {}).get("rel_l2"),
        (d.get("ranks_identical") or {}).get("ok"), d["config"].get("transport"), d["config"].get("schedule")))
    for k in r.get("kernels", []):
        print("    %-22s %7.2f us  frac %.3f" % (k["kernel"], k["avg_launch_us"], k["frac"]))
    if d.get("e2e"):
        print("    e2e %.1f GB/s  %.2f ms/step" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$name: no line (%s)" % e)
PY
  tail -2 $OUT/${TAG}_bench_n${N}_${name}.err | cut -c1-300
}
has serial      && { echo "== serial (default: fused bulk put)"; run serial 300 --steps 20 --warmup 3 ; }
has overlap     && { echo "== two-chain step"; run overlap 200 --steps 20 --warmup 3 --no-e2e --overlap ; }
has publish0    && { echo "== CF_PUBLISH_MODE=0 (round 1: every CTA fences and adds to every flag)"; CF_PUBLISH_MODE=0 run publish0 200 --steps 20 --warmup 3 --no-e2e ; }
has publish1    && { echo "== CF_PUBLISH_MODE=1"; CF_PUBLISH_MODE=1 run publish1 200 --steps 20 --warmup 3 --no-e2e ; }
has fence_sc    && { echo "== CF_PUBLISH_FENCE=sc"; CF_PUBLISH_FENCE=sc run fence_sc 200 --steps 20 --warmup 3 --no-e2e ; }
has ring_lrq    && { echo "== CogVideoX ring, LOW_RANK_Q r=32 (the example's preset)"; run ring_lrq 400 --steps 3 --warmup 3 --no-e2e --workload cogvideox5b_ring --codec lowrankq32 ; }
has lrq         && { echo "== FLUX, LOW_RANK_Q r=32"; run lrq 300 --steps 3 --warmup 3 --no-e2e --codec lowrankq32 ; }
has nobulk      && { echo "== CF_PUT_BULK=0 (round-1 sub-word remote stores)"; CF_PUT_BULK=0 run nobulk 200 --steps 20 --warmup 3 --no-e2e ; }
has dropin      && { echo "== drop-in hooks (compact_fwd per layer)"; run dropin 200 --steps 20 --warmup 3 --no-e2e --api dropin ; }
has dropin_graphs && { echo "== drop-in hooks, CF_LAYER_GRAPHS=1 (pointer-keyed per-layer graphs)"; CF_LAYER_GRAPHS=1 run dropin_graphs 200 --steps 20 --warmup 3 --no-e2e --api dropin ; }
has dropin_nccl && { echo "== drop-in hooks, NCCL transport"; run dropin_nccl 200 --steps 10 --warmup 3 --no-e2e --api dropin --transport nccl ; }
has nccl        && { echo "== engine, NCCL all-gather"; run nccl 200 --steps 10 --warmup 3 --no-e2e --transport nccl ; }
has int2        && { echo "== INT2"; run int2 200 --steps 20 --warmup 3 --no-e2e --codec int2 ; }
has int2_nobulk && { echo "== INT2, CF_PUT_BULK=0"; CF_PUT_BULK=0 run int2_nobulk 200 --steps 20 --warmup 3 --no-e2e --codec int2 ; }
has raw         && { echo "== uncompressed all-gather"; run raw 200 --steps 10 --warmup 3 --no-e2e --codec raw ; }
has raw_ring    && { echo "== uncompressed NCCL P2P ring"; run raw_ring 200 --steps 10 --warmup 3 --no-e2e --codec raw --raw-exchange ring ; }
has raw_async   && { echo "== uncompressed stale-async all-gather (DistriFusion)"; run raw_async 200 --steps 10 --warmup 3 --no-e2e --codec raw --raw-exchange async ; }
has ring        && { echo "== CogVideoX ring"; run ring 300 --steps 5 --warmup 3 --no-e2e --workload cogvideox5b_ring ; }
has ring_raw    && { echo "== CogVideoX ring, uncompressed"; run ring_raw 300 --steps 5 --warmup 3 --no-e2e --workload cogvideox5b_ring --codec raw --raw-exchange ring ; }
has pixart      && { echo "== PixArt"; run pixart 200 --steps 10 --warmup 3 --no-e2e --workload pixart_patch_parallel ; }
has pixart_raw  && { echo "== PixArt, uncompressed"; run pixart_raw 200 --steps 10 --warmup 3 --no-e2e --workload pixart_patch_parallel --codec raw ; }
has sd3         && { echo "== SD3"; run sd3 200 --steps 10 --warmup 3 --no-e2e --workload sd3_patch_parallel ; }
has sd3_raw     && { echo "== SD3, uncompressed"; run sd3_raw 200 --steps 10 --warmup 3 --no-e2e --workload sd3_patch_parallel --codec raw ; }
ls $OUT | grep ${TAG}_ | tail -40
