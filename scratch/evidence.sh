mkdir -p gpurun_out
T=r01h
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/${T}_bench_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/${T}_bench.json; cut -c1-200 gpurun_out/${T}_bench.json
python bench.py --launch single --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${T}_bench_single.json
for A in 16 64; do python bench_swarm.py --drones $A 2>&1 | tail -1 > gpurun_out/${T}_bench_swarm$A.json; done
python bench_rollout.py --precision tf32 2>&1 | tail -1 > gpurun_out/${T}_bench_rollout_fused_tf32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_step_fast.csv python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:race_step_kernel -s 20 -c 2 -o gpurun_out/${T}_race python bench.py --steps 30 --warmup 5 --launch single --no-e2e --no-cpu-baseline > gpurun_out/ncu_r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s 12 -c 1 -o gpurun_out/${T}_swarm64 python bench_swarm.py --drones 64 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_s.log 2>&1
ls gpurun_out/${T}* | wc -l
