(timeout 900 python -m pytest tests/test_race_parity_gpu.py tests/test_race_golden_gpu.py tests/test_edge_cases_gpu.py tests/test_wrappers_gpu.py tests/test_rollout_gpu.py -m gpu -x -q) 2>&1 | tail -5
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['gpu_launches'], d['clocks']['reasons'], d['episode_stats']['n'])"; }
run fused250
B2D_TAPE_CHUNK=8 run fused8
B2D_TAPE_FUSED=0 run unfused
run single "--launch single"
B2D_RACE_BALANCE=3 run fused250-balanced
B2D_LIBRARY=/root/repo/scratch/libs/lib_skipmath.so run skipmath-fused250
export B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so
python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 100 2>&1 | tail -2 | cut -c1-330
B2D_TAPE_FUSED=0 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 100 2>&1 | tail -2 | cut -c1-330
