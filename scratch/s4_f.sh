for A in 16 64; do
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s 12 -c 1 -o gpurun_out/prof_swarm$A python bench_swarm.py --drones $A --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_sw$A.log 2>&1; tail -1 gpurun_out/ncu_sw$A.log | cut -c1-200
done
