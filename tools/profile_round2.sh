#!/bin/bash
# ncu evidence for profiles/ (round 2, final build): launch list of the default command + one --set full capture per kernel family.
# Each capture follows a plain run of the same command line (profiling recipe).  Run on the GPU box: tools/profile_round2.sh <outdir>
OUT=${1:-gpurun_out/prof_r2}; mkdir -p $OUT
KEEP=${2:-none}
P="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 3 --warmup 3 --ring 8"
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  python bench.py $P "$@" > $OUT/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $OUT/$name python bench.py $P "$@" > $OUT/ncu_$name.log 2>&1
  # text summaries are what gets committed under profiles/; the reports are large (gpurun_out is capped at 64 MiB)
  { echo "# bench.py $P $@"; echo "# ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1"; python profiles/ncu_summary.py $OUT/$name.ncu-rep; echo; echo "## phases (numbered comments of the kernel header)"; python profiles/ncu_phases.py $OUT/$name.ncu-rep; echo; echo "## top source lines"; python profiles/ncu_lines.py $OUT/$name.ncu-rep 25; } > $OUT/r2_${name}_ncu_full.txt 2>&1
  ncu -i $OUT/$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]; d = dict(zip(h, rows[2]))
for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active'): print(k, d.get(k), rows[1][h.index(k)] if k in h else '')" >> $OUT/r2_${name}_ncu_full.txt 2>&1
  [ "$KEEP" = "$name" ] || rm -f $OUT/$name.ncu-rep
}
# the launch list of the DEFAULT bench command (256 streams, 16-frame ring), timed region included
D="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 5"
python bench.py $D > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file $OUT/r2_launches_default_s256.csv python bench.py $D > $OUT/ncu_launches.log 2>&1
cap k1_default_s64 k_tile_ranked 36 --streams 64
cap k1_int16_state_s64 k_tile_ranked 36 --streams 64 --state-dtype int16
cap k1_dense_s16 k_tile_ranked 14 --streams 16 --pattern dense --settle 2
cap k1_adaptive_1080p_s16 k_tile_fast2 20 --workload 1080p_time_adaptive --streams 16 --settle 8
cap k1_4x4_s16 k_tile_fast4 36 --workload 4k_rate_inhibition_4x4 --streams 16
cap k1_uint8_s16 k_tile_ranked 36 --streams 16 --frame-dtype uint8
cap k3_expand_sparse_s64 k_expand 4 --streams 64
cap k3_expand_dense_s16 k_expand 4 --streams 16 --pattern dense --settle 2
cap k2_scan_s64 k_scan_tiles 36 --streams 64
cap k1_generic_k3_s4 k_tile_pass 20 --workload 4k_rate_inhibition_4x4 --streams 4 --k1-variant generic
ls -la $OUT
