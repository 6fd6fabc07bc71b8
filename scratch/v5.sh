(timeout 900 python -m pytest tests/test_race_parity_gpu.py -m gpu -x -q -k fused) 2>&1 | tail -5
for c in 2 4 8 16 64 250; do
B2D_TAPE_CHUNK=$c python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c', d['ms_per_step']*1e3, d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
done
B2D_TAPE_CHUNK=8 B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -2 | cut -c1-330
