# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
mkdir -p gpurun_out/run
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/run/bench_4.json 2> gpurun_out/run/bench_4.err; tail -c 300 gpurun_out/run/bench_4.err
python -c "
import json
d=json.loads(open('gpurun_out/run/bench_4.json').read().strip().splitlines()[-1]); e=d['e2e']
print(d['value'], d['ms_per_step'], d['weak']['value'], e['value'], e['plain_int16_copy']['value'], e['uint8_host_frames']['value'], e['host_narrowing'], e['h2d_peak_gbs'], d['parity_check']['all_configs_equal'])"
