timeout 900 python -m pytest tests/test_wrappers_gpu.py -m gpu -x -q 2>&1 | tail -8
for A in 16 64; do python bench_swarm.py --drones $A 2>&1 | tail -1 | cut -c1-1800; done
python bench_swarm.py --drones 64 --math strict --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('strict', d['value'], d['ms_per_step'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s 10 -c 1 -o gpurun_out/prof_swarm64 python bench_swarm.py --drones 64 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_sw.log 2>&1; tail -1 gpurun_out/ncu_sw.log
