python -m pytest tests/test_gpu_host_feed.py -x -q > gpurun_out/pytest_hf.log 2>&1; tail -3 gpurun_out/pytest_hf.log
OUT=gpurun_out/s3o; mkdir -p $OUT
for LIB in main ex5 ex6 ex8; do
  if [ "$LIB" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$LIB.so; fi
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8 --graph off"
  for W in "--streams 256" "--streams 64" "--streams 32" "--streams 64 --pattern dense --settle 2" "--workload 1080p_time_adaptive --settle 8" "--workload mnist_time_bin_thr_adaptive --settle 4"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': '$LIB', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'rest_ms': round(r['scan_expand_ms'] or 0, 4), 'events': d['details']['events_last_step']}))" >> $OUT/sweep.jsonl
  done
done
cat $OUT/sweep.jsonl
