run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), d['clocks']['reasons'], d['reset_fraction_per_step'])"; }
B2D_LIBRARY=/root/repo/scratch/libs/lib_noreset.so run noreset-tape
B2D_LIBRARY=/root/repo/scratch/libs/lib_noreset.so run noreset-single "--launch single"
