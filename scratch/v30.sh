(timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['clocks']['reasons'], d['episode_stats']['n'], d['gpu_launches'])"; }
run tape
run tape
run single "--launch single"
run strict "--math strict"
python bench_rollout.py 2>&1 | tail -1 | cut -c1-170
B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -2 | cut -c1-300
