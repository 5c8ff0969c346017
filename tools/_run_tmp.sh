python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -5 gpurun_out/pytest.log
OUT=gpurun_out/shard4; mkdir -p $OUT
P="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 40 --warmup 5"
for L in main nopdl; do
if [ "$L" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$L.so; fi
python bench.py $P --streams 32 > $OUT/${L}_s32.json 2>$OUT/${L}_s32.err
python bench.py $P --workload 1080p_time_adaptive --streams 8 --settle 8 > $OUT/${L}_c4s8.json 2>$OUT/${L}_c4s8.err
python bench.py $P --workload vga_rate_inhibition > $OUT/${L}_vga.json 2>$OUT/${L}_vga.err
python bench.py $P --workload mnist_time_bin_thr_adaptive --settle 4 > $OUT/${L}_mnist.json 2>$OUT/${L}_mnist.err
done
