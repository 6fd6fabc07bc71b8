mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:race_ -s 40 -c 8 --csv --log-file gpurun_out/launch_b.csv python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 1 -o gpurun_out/prof_b python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
