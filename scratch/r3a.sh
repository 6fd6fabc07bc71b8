# round-1 session-5 baseline: full GPU tests, smoke, bench lines, launch list
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
(time timeout 1500 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/r3a_bench.json; cut -c1-1500 gpurun_out/r3a_bench.json
python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r3a_bench_ref.json; cut -c1-600 gpurun_out/r3a_bench_ref.json
for A in 16 64; do python bench_swarm.py --drones $A --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r3a_swarm$A.json | cut -c1-700; done
python bench_rollout.py 2>&1 | tail -3 | tee gpurun_out/r3a_rollout.txt | cut -c1-700
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3a_launches.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/r3a_ncu_b.log 2>&1; tail -1 gpurun_out/r3a_ncu_b.log | cut -c1-200
