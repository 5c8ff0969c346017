OUT=gpurun_out/s3m; mkdir -p $OUT
python -m pytest tests/test_gpu_compact_state.py tests/test_gpu_fullsize.py -x -q > $OUT/pytest_main.log 2>&1; tail -3 $OUT/pytest_main.log
for LIB in main adpt6 adpt4 adpteager; do
  if [ "$LIB" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$LIB.so; fi
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 4"
  for W in "--workload 1080p_time_adaptive --settle 8" "--workload 1080p_time_adaptive --settle 8 --frame-dtype uint8" "--workload 1080p_time_adaptive --settle 8 --streams 8" "--workload mnist_time_bin_thr_adaptive --settle 4" "--workload 1080p_time_adaptive --settle 8 --pattern dense"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': '$LIB', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'frac': round(r['frac'], 3), 'events': d['details']['events_last_step']}))" >> $OUT/sweep.jsonl
  done
done
cat $OUT/sweep.jsonl
