# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
mkdir -p gpurun_out/run
python bench.py --steps 20 --warmup 5 --configs cfg2_pipeline_virtualcam_4096_streams,cfg2_mnist_shape_4096_streams_time_bin_thr_adaptive --no-cpu-baseline --no-e2e > gpurun_out/run/bench_p.json 2> gpurun_out/run/bench_p.err; tail -c 300 gpurun_out/run/bench_p.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_p.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','stream_frames_per_s','ms_per_step_with_sink','stream_frames_per_s_with_sink','parity','error')})"
