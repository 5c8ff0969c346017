python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tpc in 1 2 4 8; do PYDVS_TILES_PER_CTA=$tpc python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tpc $tpc', d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'], d['roofline']['scan_expand_ms'])"; done
