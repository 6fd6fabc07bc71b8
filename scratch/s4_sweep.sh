for n in b128c3 b128c4 b128c5 b64c8 b64c6 b256c2 b256c1 b32c16; do
  B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', d['ms_per_step']*1e3, d['roofline']['frac'])"
done
