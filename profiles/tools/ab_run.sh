#!/bin/bash
# usage: profiles/tools/ab_run.sh "<bench.py args>" name [name ...]   (names of build/variants/lib_<name>.so; "base" = the in-tree library)
args="$1"; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset B2D_LIBRARY; else export B2D_LIBRARY=/root/repo/build/variants/lib_$v.so; fi
  python bench.py $args --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$args', round(d['ms_per_step']*1000,2), 'us', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
