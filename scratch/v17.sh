mkdir -p gpurun_out
# clocks/power during the timed region of a long run
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/r01e_clocks.csv &
SMI=$!
sleep 0.5
python bench.py --no-e2e --no-cpu-baseline --steps 30000 --warmup 100 2>&1 | tail -1 > gpurun_out/r01e_bench_long.json; cut -c1-120 gpurun_out/r01e_bench_long.json
kill $SMI
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/r01e_bench_reference.json; cut -c1-160 gpurun_out/r01e_bench_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/r01e_bench.json; cut -c1-200 gpurun_out/r01e_bench.json
python bench.py --math strict --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r01e_bench_strict.json
for A in 16 64; do python bench_swarm.py --drones $A 2>&1 | tail -1 > gpurun_out/r01e_bench_swarm$A.json; cut -c1-160 gpurun_out/r01e_bench_swarm$A.json; done
for pr in tf32 fp32; do python bench_rollout.py --precision $pr 2>&1 | tail -1 > gpurun_out/r01e_bench_rollout_fused_$pr.json; cut -c1-160 gpurun_out/r01e_bench_rollout_fused_$pr.json; done
# launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches_step_fast.csv python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:race_step_kernel -s 20 -c 2 -o gpurun_out/r01e_race python bench.py --steps 30 --warmup 5 --launch single --no-e2e --no-cpu-baseline > gpurun_out/ncu_r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s 12 -c 1 -o gpurun_out/r01e_swarm64 python bench_swarm.py --drones 64 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_act -s 3 -c 1 -o gpurun_out/r01e_policy python bench_rollout.py --horizon 8 --replays 2 --no-graph > gpurun_out/ncu_p.log 2>&1
ncu --set full --clock-control none -k regex:puff_advantage -c 1 -o gpurun_out/r01e_adv python -m pytest tests/test_advantage_gpu.py -m gpu -q -x > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
