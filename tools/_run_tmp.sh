# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`: GPU parity suite + one default bench run
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
mkdir -p gpurun_out/run
python bench.py --steps 20 --warmup 5 > gpurun_out/run/bench_default.json 2> gpurun_out/run/bench_default.err; tail -c 300 gpurun_out/run/bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_default.json').read().strip().splitlines()[-1]); e=d['e2e']
print(d['value'], d['ms_per_step'], d['roofline']['k1_ms'], e['value'], e['plain_int16_copy']['value'], d['parity_check']['all_configs_equal'], len(d['configs']))
print({k:v.get('error') for k,v in d['configs'].items() if 'error' in v})"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
