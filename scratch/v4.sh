(timeout 900 python -m pytest tests/test_race_parity_gpu.py tests/test_race_golden_gpu.py tests/test_edge_cases_gpu.py tests/test_wrappers_gpu.py tests/test_rollout_gpu.py -m gpu -x -q) 2>&1 | tail -12
for l in tape single; do
python bench.py --no-e2e --no-cpu-baseline --launch $l 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', d['ms_per_step']*1e3, d['roofline']['frac'], d['gpu_launches'], d['reset_fraction_per_step'], d['episode_stats'])"
done
B2D_TAPE_FUSED=0 python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tape-unfused', d['ms_per_step']*1e3, d['roofline']['frac'], d['gpu_launches'])"
B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -2 | cut -c1-330
B2D_TAPE_FUSED=0 B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -2 | cut -c1-330
