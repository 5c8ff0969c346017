OUT=gpurun_out/s3q; mkdir -p $OUT
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.json 2>$OUT/bench_2gpu.err ) 2> $OUT/bench_2gpu.time
tail -c 600 $OUT/bench_2gpu.err; cat $OUT/bench_2gpu.time; tail -c 3000 $OUT/bench_2gpu.json
