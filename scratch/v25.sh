(timeout 900 python -m pytest tests/test_race_parity_gpu.py tests/test_race_golden_gpu.py tests/test_full_size_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q) 2>&1 | tail -3
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['clocks']['reasons'], d['episode_stats']['n'])"; }
run tape
run single "--launch single"
export B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so
python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -2 | cut -c1-300
python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 --launch single 2>&1 | tail -2 | cut -c1-300
