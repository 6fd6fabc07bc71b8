(timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -4
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['clocks']['reasons'], d['episode_stats']['n'])"; }
run tape
run tape
run single "--launch single"
python bench_swarm.py --drones 64 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('swarm64', d['ms_per_step'], d['roofline']['frac'])"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:race_step_kernel -s 20 -c 1 python bench.py --steps 30 --warmup 5 --launch single --no-e2e --no-cpu-baseline 2>&1 | grep -E "smsp__inst|gpu__time"
