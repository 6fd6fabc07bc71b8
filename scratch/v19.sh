run() { python bench.py --no-e2e --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4))"; }
for k in 0 0xe0 0x3e0 0x300; do B2D_RACE_L2_KEEP=$k run keep-$k; done
for k in 0 0x3e0 0xe0; do
B2D_RACE_L2_KEEP=$k ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:race_step_kernel -s 30 -c 1 --csv --log-file gpurun_out/l2keep_$k.csv python bench.py --steps 40 --warmup 5 --launch single --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l2keep_$k.csv")) if len(r)>10]
h=rows[0]
print("$k", [(r[h.index("Metric Name")], r[h.index("Metric Value")]) for r in rows[1:]])
PY
done
