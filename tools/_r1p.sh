OUT=gpurun_out; TAG=r1p; mkdir -p $OUT
(time timeout 200 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fused or two_gpu") > $OUT/${TAG}_tests.log 2>&1 ; tail -15 $OUT/${TAG}_tests.log
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 20 --warmup 3 --hang-dump 90 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
tail -c 300 $OUT/${TAG}_bench_n2.err
