# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
OUT=gpurun_out/s4a; mkdir -p $OUT
D="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 5"
python bench.py $D > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1200 --csv --log-file $OUT/r2_launches_default_s256.csv python bench.py $D > $OUT/ncu_launches.log 2>&1
tail -n 3 $OUT/r2_launches_default_s256.csv | cut -c1-220; wc -l $OUT/r2_launches_default_s256.csv; tail -c 400 $OUT/plain_launches.log
