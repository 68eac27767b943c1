OUT=gpurun_out; TAG=r1k; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_smi.txt 2>&1
(time timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "two_gpu") > $OUT/${TAG}_tests2.log 2>&1 ; tail -4 $OUT/${TAG}_tests2.log
for N in 2 4; do
  CF_BENCH_VERBOSE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 10 --warmup 3 --hang-dump 200 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  tail -c 3000 $OUT/${TAG}_bench_n$N.json; tail -5 $OUT/${TAG}_bench_n$N.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 10 --warmup 3 --transport nccl --no-e2e --hang-dump 150 > $OUT/${TAG}_bench_n4_nccl.json 2> $OUT/${TAG}_bench_n4_nccl.err
tail -c 1500 $OUT/${TAG}_bench_n4_nccl.json
