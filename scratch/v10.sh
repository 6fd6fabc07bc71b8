nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 50 > gpurun_out/smi_math.csv &
SMI=$!
sleep 1
B2D_RACE_BALANCE=0 python bench.py --no-e2e --no-cpu-baseline --steps 40000 --warmup 100 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('math', round(d['ms_per_step']*1e3,2), d['clocks'])"
sleep 1
kill $SMI
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 50 > gpurun_out/smi_skip.csv &
SMI=$!
sleep 1
B2D_LIBRARY=/root/repo/scratch/libs/lib_skipmath.so B2D_RACE_BALANCE=0 python bench.py --no-e2e --no-cpu-baseline --steps 40000 --warmup 100 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skipmath', round(d['ms_per_step']*1e3,2), d['clocks'])"
sleep 1
kill $SMI
python - <<'PY'
import csv
for n in ("math","skip"):
    rows=list(csv.reader(open(f"gpurun_out/smi_{n}.csv")))[1:]
    busy=[r for r in rows if float(r[3].split()[0])>300]
    print(n, len(rows), "samples;", len(busy), "busy")
    if busy:
        import statistics
        print("  sm MHz", statistics.median(int(r[1].split()[0]) for r in busy), "min", min(int(r[1].split()[0]) for r in busy), "power W median", statistics.median(float(r[3].split()[0]) for r in busy), "max", max(float(r[3].split()[0]) for r in busy), "limit", busy[0][4], "temp", busy[-1][5], "reasons", set(r[6].strip() for r in busy))
PY
