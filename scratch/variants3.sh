run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reset_fraction_per_step'])"; }
export B2D_LIBRARY=scratch/libs/lib_cur.so
echo full; run
echo full-20000steps; run --steps 20000
echo noadopt; B2D_EXPERIMENT_NO_ADOPT=1 run
echo mathskip-noadopt; B2D_LIBRARY=scratch/libs/lib_128_4_s0_m1.so B2D_EXPERIMENT_NO_ADOPT=1 run
