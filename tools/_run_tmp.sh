# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
mkdir -p gpurun_out/run
python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline > gpurun_out/run/bench_e.json 2> gpurun_out/run/bench_e.err; tail -c 300 gpurun_out/run/bench_e.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_e.json').read().strip().splitlines()[-1]); e=d['e2e']
print(e['value'], e['plain_int16_copy']['value'], e['uint8_host_frames']['value'], e['host_narrowing'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-weak-leg > gpurun_out/run/bench_e2.json 2> gpurun_out/run/bench_e2.err; tail -c 300 gpurun_out/run/bench_e2.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_e2.json').read().strip().splitlines()[-1]); e=d['e2e']
print(d['value'], e['value'], e['plain_int16_copy']['value'], e['uint8_host_frames']['value'], e['host_narrowing'], e['h2d_peak_gbs'])"
