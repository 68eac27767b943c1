OUT=gpurun_out; TAG=r1j; mkdir -p $OUT
(time timeout 600 python -m pytest tests -m gpu -x -q) > $OUT/${TAG}_tests.log 2>&1 ; tail -4 $OUT/${TAG}_tests.log
timeout 300 python sweep.py --sizes-mb 8,256 --shapes 4608x3072,576x3072,4096x3072 --ops int4,int8 --reps 5 --out $OUT/${TAG}_sweep.jsonl > $OUT/${TAG}_sweep.log 2>&1; cat $OUT/${TAG}_sweep.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_lr' --csv --log-file $OUT/${TAG}_lr_launches.csv python tools/lr_probe.py 32 32 > $OUT/${TAG}_lr.log 2>&1; tail -2 $OUT/${TAG}_lr.log
