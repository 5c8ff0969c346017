OUT=gpurun_out/s3u; mkdir -p $OUT
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_4gpu.json 2>$OUT/bench_4gpu.err ) 2> $OUT/bench_4gpu.time
tail -c 600 $OUT/bench_4gpu.err; cat $OUT/bench_4gpu.time; tail -c 1500 $OUT/bench_4gpu.json
