run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -2 | cut -c1-400; }
for v in t0 tc2 tsk; do f=scratch/libs/lib_$v.so; echo "== $f"; B2D_LIBRARY=$f run; done
