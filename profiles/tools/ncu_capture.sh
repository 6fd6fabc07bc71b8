#!/bin/bash
# usage: profiles/tools/ncu_capture.sh tag workload kernel_regex [launch_skip] [launch_count]
# one `ncu --set full` capture of a step kernel out of a short bench.py run -> gpurun_out/<tag>_<workload>.ncu-rep
tag=$1; w=$2; k=$3; skip=${4:-30}; cnt=${5:-1}; extra=${6:-}
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -f -o gpurun_out/${tag}_$w \
    python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-extra $extra > gpurun_out/${tag}_ncu_$w.log 2>&1
ls -la gpurun_out/${tag}_$w.ncu-rep
