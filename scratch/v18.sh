(timeout 900 python -m pytest tests/test_race_parity_gpu.py tests/test_race_golden_gpu.py -m gpu -x -q) 2>&1 | tail -2
run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), d['gpu_launches'], d['clocks']['reasons'], d['episode_stats']['n'])"; }
for k in 0 0xe0 0x3e0 0x300 0x3ff; do B2D_RACE_L2_KEEP=$k run keep-$k; done
B2D_RACE_L2_KEEP=0x3e0 run keep-0x3e0-single "--launch single"
B2D_RACE_L2_KEEP=0 run keep-0-single "--launch single"
for k in 0 0x3e0; do
B2D_RACE_L2_KEEP=$k ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:race_step_kernel -s 30 -c 2 --csv --log-file gpurun_out/l2keep_$k.csv python bench.py --steps 40 --warmup 5 --launch single --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l2keep_$k.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print("$k", r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
done
