OUT=gpurun_out/s3k; mkdir -p $OUT
PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_bd512.so python -m pytest tests/test_gpu_compact_state.py tests/test_gpu_fullsize.py -x -q > $OUT/pytest_bd512.log 2>&1; tail -3 $OUT/pytest_bd512.log
for LIB in main bd512; do
  if [ "$LIB" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$LIB.so; fi
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8"
  for W in "--streams 64" "--streams 64 --frame-dtype uint8" "--streams 64 --pattern dense --settle 2" "--streams 32" "--streams 256" "--streams 256 --frame-dtype uint8"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': '$LIB', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'frac': round(r['frac'], 3), 'events': d['details']['events_last_step']}))" >> $OUT/sweep.jsonl
  done
done
cat $OUT/sweep.jsonl
