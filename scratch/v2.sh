for n in timing skipmath; do
B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -3 | cut -c1-400
done
B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 --launch single 2>&1 | tail -3 | cut -c1-400
