python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -5 gpurun_out/pytest.log
OUT=gpurun_out/shard6; mkdir -p $OUT
P="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 40 --warmup 5"
python bench.py $P --workload mnist_time_bin_thr_adaptive --settle 4 > $OUT/main_mnist.json 2>$OUT/main_mnist.err
python bench.py $P --workload vga_rate_inhibition --streams 2048 > $OUT/main_vga2048.json 2>$OUT/main_vga2048.err
python bench.py $P --workload davis346_rate > $OUT/main_davis.json 2>$OUT/main_davis.err
python bench.py $P --streams 64 --ring 8 > $OUT/main_s64.json 2>$OUT/main_s64.err
Q="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $OUT/launches_mnist.csv python bench.py $Q --workload mnist_time_bin_thr_adaptive --settle 4 --graph off > $OUT/ncu1.log 2>&1
