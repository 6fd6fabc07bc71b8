#!/bin/bash
# usage (on the GPU box, from the repo root): profiles/tools/evidence.sh <tag>
# The round's evidence set -> gpurun_out/<tag>_*: GPU tests, the default bench line (all workloads) and the reference
# arm, the ncu launch list of a short bench run, one `ncu --set full` capture per hot kernel.
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_gpu.log
( time python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2> gpurun_out/${tag}_bench.time
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${tag}_launches_bench.log 2>&1
profiles/tools/ncu_capture.sh $tag race race_step_kernel 30 2
profiles/tools/ncu_capture.sh $tag swarm16 swarm_kernel 30 1
profiles/tools/ncu_capture.sh $tag swarm64 swarm_kernel 30 1
profiles/tools/ncu_capture.sh $tag rollout race_rollout_kernel 1 1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt
ls -la gpurun_out | grep $tag
