python -m pytest tests/test_race_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for f in scratch/libs/*.so; do echo "== $f"; B2D_LIBRARY=$f python bench.py --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
