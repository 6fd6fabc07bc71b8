set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Core|Socket"
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01a.json 2> gpurun_out/bench_r01a.err; tail -c 3000 gpurun_out/bench_r01a.json; tail -5 gpurun_out/bench_r01a.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r01a.json 2>&1; tail -c 1500 gpurun_out/bench_ref_r01a.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:race_step -s 30 -c 2 -o gpurun_out/prof_r01_fast python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
