python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
OUT=gpurun_out/s3h; mkdir -p $OUT
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8"
  for W in "--streams 64" "--streams 64 --pattern dense --settle 2" "--streams 64 --frame-dtype uint8" "--streams 64 --state-dtype int16" "--workload vga_rate_inhibition --streams 2048" "--workload 1080p_time_adaptive --settle 8" "--workload 4k_rate_no_inhibition --streams 64" "--streams 32"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': 'main', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'frac': round(r['frac'], 3), 'events': d['details']['events_last_step']}))" >> $OUT/sweep.jsonl
  done
cat $OUT/sweep.jsonl
python bench.py --no-e2e --no-cpu-baseline --no-parity --configs none --steps 5 --warmup 3 > $OUT/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 500 --csv --log-file $OUT/launches_default_s256.csv python bench.py --no-e2e --no-cpu-baseline --no-parity --configs none --steps 5 --warmup 3 > $OUT/ncu_launches.log 2>&1
tail -n 8 $OUT/launches_default_s256.csv | cut -c1-200
