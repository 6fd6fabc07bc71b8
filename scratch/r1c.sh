mkdir -p gpurun_out
run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reset_fraction_per_step'], d['clocks']['sm_mhz'])"; }
for f in scratch/libs/lib_v5.so; do echo "== $f"; B2D_LIBRARY=$f run; done
B2D_LIBRARY=scratch/libs/lib_v3.so ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 1 -o gpurun_out/prof_v3 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log
