# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`: GPU parity suite + the default bench run of both arms
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
mkdir -p gpurun_out/run
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/run/bench_reference.json 2> gpurun_out/run/bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/run/bench_default.json 2> gpurun_out/run/bench_default.err; tail -c 300 gpurun_out/run/bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_default.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/run/bench_reference.json').read().strip().splitlines()[-1])
e=d['e2e']
print(d['value'], d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'], e['value'], e['plain_int16_copy']['value'], e['uint8_host_frames']['value'], d['parity_check']['all_configs_equal'], r['value'], e['value']/r['value'], r['cpu_baseline']['cores'])
print(e['host_narrowing'])"
