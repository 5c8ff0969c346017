#!/bin/bash
# usage: tools/k1_sweep_pair.sh <out.jsonl> <lib> ...   -- K1 / step times of the 2x2-inhibition int16 workloads per library build
OUT=$1; shift; mkdir -p $(dirname $OUT)
for LIB in "$@"; do
  if [ "$LIB" = "main" ]; then unset PYDVS_B200_LIB; else export PYDVS_B200_LIB=$PWD/pydvs_b200/libpydvs_b200_$LIB.so; fi
  B="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 20 --warmup 3 --ring 8"
  for W in "--streams 64" "--streams 64 --pattern dense --settle 2" "--streams 64 --ring 4" "--workload vga_rate_inhibition --streams 2048"; do
    python bench.py $B $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print(json.dumps({'lib': '$LIB', 'args': '$W', 'ms': round(d['ms_per_step'], 4), 'k1_ms': round(r['k1_ms'], 4), 'frac': round(r['frac'], 3), 'events': d['details']['events_last_step']}))" >> $OUT
  done
done
