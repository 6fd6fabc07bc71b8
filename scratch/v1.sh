mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6 | tee gpurun_out/v1_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/v1_bench_reference.json; cut -c1-200 gpurun_out/v1_bench_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/v1_bench.json; cut -c1-300 gpurun_out/v1_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v1_launches.csv python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/v1_ncu.log 2>&1; tail -1 gpurun_out/v1_ncu.log | cut -c1-200
