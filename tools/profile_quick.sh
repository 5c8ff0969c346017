#!/bin/bash
# usage: tools/profile_quick.sh <outdir> <name> <kernel regex> <skip> <bench args...>   -- one --set full capture with text summaries
# (PYDVS_B200_LIB selects a variant build; see tools/profile_round2.sh for the full set)
OUT=$1; name=$2; rx=$3; skip=$4; shift 4; mkdir -p $OUT
P="--no-e2e --no-cpu-baseline --no-parity --configs none --steps 3 --warmup 3 --ring 8"
python bench.py $P "$@" > $OUT/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $OUT/$name python bench.py $P "$@" > $OUT/ncu_$name.log 2>&1
{ echo "# bench.py $P $@"; echo "# ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1"; python profiles/ncu_summary.py $OUT/$name.ncu-rep; echo; echo "## phases (numbered comments of the kernel header)"; python profiles/ncu_phases.py $OUT/$name.ncu-rep; echo; echo "## top source lines"; python profiles/ncu_lines.py $OUT/$name.ncu-rep 40; } > $OUT/${name}_ncu_full.txt 2>&1
rm -f $OUT/$name.ncu-rep
