run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reset_fraction_per_step'], d['clocks']['sm_mhz'])"; }
for v in st c1 c2 c4 c8; do f=scratch/libs/lib_$v.so; echo "== $f"; B2D_LIBRARY=$f run; done
echo "== parity c4"; B2D_LIBRARY=scratch/libs/lib_c4.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== parity st"; B2D_LIBRARY=scratch/libs/lib_st.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
