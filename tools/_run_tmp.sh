# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`: GPU parity suite + one default bench run
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
mkdir -p gpurun_out/run
python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline > gpurun_out/run/bench.json 2> gpurun_out/run/bench.err; tail -c 300 gpurun_out/run/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['k1_ms'], d['e2e']['value'], d['parity_check']['all_configs_equal'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
