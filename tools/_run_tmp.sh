# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
mkdir -p gpurun_out/run
python -m pytest tests/test_gpu_host_feed.py tests/test_sinks.py tests/test_gpu_next_rows.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --configs cfg2_pipeline_virtualcam_4096_streams --no-cpu-baseline --no-e2e > gpurun_out/run/bench_p.json 2> gpurun_out/run/bench_p.err; tail -c 400 gpurun_out/run/bench_p.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_p.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','stream_frames_per_s','ms_per_step_with_sink','stream_frames_per_s_with_sink','ms_per_step_with_synchronous_sink','parity','error')})"
