run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -8 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['gpu_launches'])
    else: print(l.rstrip()[:200])"; }
export B2D_BALANCE_DEBUG=1
B2D_BALANCE_DUMP=gpurun_out/balance.csv B2D_RACE_BALANCE=6 run balanced6
