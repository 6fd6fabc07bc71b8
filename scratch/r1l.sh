timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], d['ms_per_step'], d['roofline']['frac']); print('e2e', d['e2e'])"
for A in 16 64; do python bench_swarm.py --drones $A --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('swarm', d['config']['rows'], d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
