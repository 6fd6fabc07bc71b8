export B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so
python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 100 2>&1 | tail -2 | cut -c1-330
