(timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>&1 | tail -1 > gpurun_out/r01f_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/r01f_bench.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
python bench.py --math strict --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r01f_bench_strict.json
for pr in tf32; do python bench_rollout.py --precision $pr 2>&1 | tail -1 > gpurun_out/r01f_bench_rollout_fused_$pr.json; cut -c1-200 gpurun_out/r01f_bench_rollout_fused_$pr.json; done
