B2D_LIBRARY=scratch/libs/lib_tm.so python bench.py --no-cpu-baseline --no-e2e 2>&1 | grep timing
ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 1 -o gpurun_out/prof_lazy python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_lazy.log 2>&1
tail -1 gpurun_out/ncu_lazy.log
