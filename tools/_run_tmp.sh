python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
OUT=gpurun_out/s3f; mkdir -p $OUT
P="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 5"
for sd in int16 uint8; do
python bench.py $P --state-dtype $sd > $OUT/cfg5_$sd.json 2>$OUT/cfg5_$sd.err
python bench.py $P --state-dtype $sd --frame-dtype uint8 > $OUT/cfg5_u8f_$sd.json 2>$OUT/cfg5_u8f_$sd.err
python bench.py $P --state-dtype $sd --pattern dense --settle 2 --ring 4 > $OUT/dense_$sd.json 2>$OUT/dense_$sd.err
python bench.py $P --state-dtype $sd --workload 1080p_time_adaptive --settle 8 --ring 4 > $OUT/cfg4_$sd.json 2>$OUT/cfg4_$sd.err
python bench.py $P --state-dtype $sd --workload 1080p_time_adaptive --settle 8 --ring 4 --frame-dtype uint8 > $OUT/cfg4_u8f_$sd.json 2>$OUT/cfg4_u8f_$sd.err
python bench.py $P --state-dtype $sd --streams 32 > $OUT/cfg5_s32_$sd.json 2>$OUT/cfg5_s32_$sd.err
done
for f in $OUT/*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['ms_per_step'], r['k1_ms'], r.get('scan_expand_ms'), r['state_dtype'], d['details']['events_last_step'])"; done
