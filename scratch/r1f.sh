B2D_LIBRARY=scratch/libs/lib_c2.so ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 1 -o gpurun_out/prof_c2 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
