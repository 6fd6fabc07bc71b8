run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reset_fraction_per_step'], d['clocks']['sm_mhz'], d['episode_stats'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run
run --launch single
run --steps 8000
B2D_LIBRARY=scratch/libs/lib_tm.so python bench.py --no-cpu-baseline --no-e2e 2>&1 | grep timing
