mkdir -p gpurun_out
T=r01g
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/${T}_clocks.csv &
SMI=$!
sleep 0.5
python bench.py --no-e2e --no-cpu-baseline --steps 30000 --warmup 100 2>&1 | tail -1 > gpurun_out/${T}_bench_long.json
kill $SMI
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/${T}_bench_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/${T}_bench.json; cut -c1-200 gpurun_out/${T}_bench.json
python bench.py --launch single --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${T}_bench_single.json
python bench.py --math strict --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${T}_bench_strict.json
for pr in tf32 fp32; do python bench_rollout.py --precision $pr 2>&1 | tail -1 > gpurun_out/${T}_bench_rollout_fused_$pr.json; done
for A in 16 64; do python bench_swarm.py --drones $A 2>&1 | tail -1 > gpurun_out/${T}_bench_swarm$A.json; done
for n in noreset dmath noobs; do B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so python bench.py --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${T}_variant_$n.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_step_fast.csv python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:race_step_kernel -s 20 -c 2 -o gpurun_out/${T}_race python bench.py --steps 30 --warmup 5 --launch single --no-e2e --no-cpu-baseline > gpurun_out/ncu_r.log 2>&1
ls gpurun_out/${T}* | wc -l
