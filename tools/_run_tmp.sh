python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
rm -rf gpurun_out/sweep8; bash tools/k1_sweep_adpt.sh gpurun_out/sweep8/adpt.jsonl main te0
export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_j2.so
B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8"
for W in "--workload 1080p_time_adaptive --settle 8 --tile-rows 2" "--workload 1080p_time_adaptive --settle 8 --tile-rows 2 --pattern dense"; do
python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': 'j2', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'frac': round(r['frac'], 3), 'events': d['details']['events_last_step']}))" >> gpurun_out/sweep8/adpt.jsonl
done
unset PYDVS_B200_LIB
bash tools/k1_sweep.sh gpurun_out/sweep8/all.jsonl main
cat gpurun_out/sweep8/adpt.jsonl gpurun_out/sweep8/all.jsonl
