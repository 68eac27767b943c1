#!/bin/bash
# One multi-GPU box session (N = 2, 4 or 8): the multi-rank parity tests, then bench.py at N ranks for the
# default (one-sided exchange fused into the codec kernels), the separate put kernel, NCCL, INT2, the
# uncompressed baseline and the ring workload.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_multi.sh r2a 2'
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out
export CF_EXPERIMENTAL=1  # arm the tests of opt-in features (two-chain step)
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_smi.txt 2>&1
echo "== multi-rank tests"
(time timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_zz_ring_engine.py -m gpu -q -k "two_gpu or fused or ring") \
  > $OUT/${TAG}_tests.log 2>&1 ; tail -6 $OUT/${TAG}_tests.log
PORT=29500
run() {  # run <name> <timeout> [env VAR=..] -- bench args
  local name=$1 tmo=$2; shift 2
  PORT=$((PORT + 1))
  CF_BENCH_VERBOSE=1 timeout $tmo python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $PORT bench.py --gpus $N --hang-dump $((tmo - 10)) "$@" > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  tail -c 400 $OUT/${TAG}_bench_n${N}_${name}.json; echo; tail -2 $OUT/${TAG}_bench_n${N}_${name}.err
}
echo "== bench: fused put (default)" ; run fused 200 --steps 20 --warmup 3
echo "== bench: separate put kernel" ; CF_FUSED_PUT=0 run putkernel 120 --steps 10 --warmup 3 --no-e2e
echo "== bench: NCCL all-gather"     ; run nccl 120 --steps 10 --warmup 3 --no-e2e --transport nccl
echo "== bench: two-chain step"      ; run overlap 120 --steps 10 --warmup 3 --no-e2e --overlap
echo "== bench: INT2"                ; run int2 120 --steps 10 --warmup 3 --no-e2e --codec int2
echo "== bench: uncompressed"        ; run raw 120 --steps 10 --warmup 3 --no-e2e --codec raw
echo "== bench: CogVideoX ring"      ; run ring 200 --steps 5 --warmup 3 --no-e2e --workload cogvideox5b_ring
echo "== bench: PixArt"              ; run pixart 120 --steps 10 --warmup 3 --no-e2e --workload pixart_patch_parallel
ls -la $OUT | tail -20
