python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
OUT=gpurun_out/s3w; mkdir -p $OUT
( time python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2>$OUT/bench_default.err ) 2> $OUT/bench_default.time
tail -c 300 $OUT/bench_default.err; cat $OUT/bench_default.time
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
