mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 2500 gpurun_out/bench_r01b.json; tail -3 gpurun_out/bench_r01b.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 2 -o gpurun_out/prof_r01b_fast python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log
python bench.py --math strict --no-cpu-baseline --no-e2e > gpurun_out/bench_r01b_strict.json 2>&1; tail -c 600 gpurun_out/bench_r01b_strict.json
