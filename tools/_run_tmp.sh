python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -12 gpurun_out/pytest.log
OUT=gpurun_out/s3t; mkdir -p $OUT
B="--no-e2e --no-cpu-baseline --configs none --steps 20 --warmup 3 --ring 8"
python bench.py $B --workload davis346_rate_inhibition > $OUT/davis_inh.json 2>$OUT/davis_inh.err; tail -c 300 $OUT/davis_inh.err
python bench.py $B --workload davis346_rate_inhibition --k1-variant generic > $OUT/davis_inh_generic.json 2>$OUT/davis_inh_generic.err
python bench.py $B --workload davis346_rate > $OUT/davis.json 2>$OUT/davis.err
for f in davis_inh davis_inh_generic davis; do python -c "
import json
d=json.loads(open('$OUT/$f.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$f', d['ms_per_step'], r['k1_ms'], r['frac'], d['details']['events_last_step'], d.get('parity_check',{}).get('all_configs_equal'))"; done
