(timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -3
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['gpu_launches'], d['clocks']['reasons'], d['episode_stats']['n'])"; }
run tiled-tape
run tiled-tape
run tiled-single "--launch single"
export B2D_LIBRARY=/root/repo/scratch/libs/lib_planar.so
run planar-tape
run planar-tape
run planar-single "--launch single"
