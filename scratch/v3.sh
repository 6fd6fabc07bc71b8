(timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('main', d['ms_per_step']*1e3, d['roofline']['frac'], d['reset_fraction_per_step'], d['episode_stats'])"
B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -3 | cut -c1-330
