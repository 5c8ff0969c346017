OUT=gpurun_out/s3s; mkdir -p $OUT
PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_qskip.so python -m pytest tests/test_gpu_compact_state.py -x -q > $OUT/pytest_qskip.log 2>&1; tail -2 $OUT/pytest_qskip.log
for LIB in main qskip; do
  if [ "$LIB" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$LIB.so; fi
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8"
  for W in "--streams 256" "--streams 64" "--streams 64 --pattern dense --settle 2" "--streams 64 --frame-dtype uint8" "--workload vga_rate_inhibition --streams 2048"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': '$LIB', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'events': d['details']['events_last_step']}))" >> $OUT/sweep.jsonl
  done
done
cat $OUT/sweep.jsonl
