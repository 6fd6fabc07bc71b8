run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['clocks']['reasons'], d['episode_stats']['n'])"; }
for n in inline inline3 inline5; do
B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so run $n-tape
B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so run $n-single "--launch single"
done
run base-tape
B2D_LIBRARY=/root/repo/scratch/libs/lib_inline.so timeout 600 python -m pytest tests/test_race_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
