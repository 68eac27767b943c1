OUT=gpurun_out; TAG=r1m; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_smi.txt 2>&1
date +%s > $OUT/${TAG}_t0.txt
CF_BENCH_VERBOSE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 10 --warmup 3 --hang-dump 170 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
tail -c 2500 $OUT/${TAG}_bench_n2.json; tail -4 $OUT/${TAG}_bench_n2.err
CF_BENCH_VERBOSE=1 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 \
    bench.py --gpus 2 --steps 10 --warmup 3 --transport nccl --no-e2e --hang-dump 80 > $OUT/${TAG}_bench_n2_nccl.json 2> $OUT/${TAG}_bench_n2_nccl.err
tail -c 600 $OUT/${TAG}_bench_n2_nccl.json; tail -3 $OUT/${TAG}_bench_n2_nccl.err
date +%s >> $OUT/${TAG}_t0.txt
