export B2D_LIBRARY=/root/repo/scratch/libs/lib_timing.so
B2D_TRACE_FILE=gpurun_out/trace_fused.csv python bench.py --no-e2e --no-cpu-baseline --steps 1000 --warmup 250 2>&1 | tail -2 | cut -c1-200
B2D_TAPE_FUSED=0 B2D_TRACE_FILE=gpurun_out/trace_unfused.csv python bench.py --no-e2e --no-cpu-baseline --steps 1000 --warmup 250 2>&1 | tail -2 | cut -c1-200
python - <<'PY'
import csv, collections
for name in ("fused", "unfused"):
    rows = list(csv.DictReader(open(f"gpurun_out/trace_{name}.csv")))
    dur = [(int(r["done_ns"]) - int(r["go_ns"])) / 1e3 for r in rows]
    wait = [(int(r["go_ns"]) - int(r["entry_ns"])) / 1e3 for r in rows]
    t0 = min(int(r["entry_ns"]) for r in rows); t1 = max(int(r["done_ns"]) for r in rows)
    print(name, "launch span us", (t1 - t0) / 1e3, "cta dur us: min %.1f mean %.1f max %.1f" % (min(dur), sum(dur) / len(dur), max(dur)), "chain wait mean %.1f max %.1f" % (sum(wait) / len(wait), max(wait)))
    by = collections.defaultdict(list)
    for r, d in zip(rows, dur): by[int(r["smid"])].append(d)
    sm = sorted((sum(v) / len(v), k, len(v)) for k, v in by.items())
    print("  per-SM mean dur: slowest", [(k, round(m, 1), n) for m, k, n in sm[-8:]], "fastest", [(k, round(m, 1), n) for m, k, n in sm[:8]])
    import statistics
    within = statistics.mean(max(v) - min(v) for v in by.values() if len(v) > 1)
    print("  mean within-SM spread %.1f us; across-SM sd %.1f us" % (within, statistics.pstdev([m for m, _, _ in sm])))
    s = sorted(dur); print("  dur percentiles", [round(s[int(len(s) * q)], 1) for q in (0.05, 0.25, 0.5, 0.75, 0.95, 0.99)])
PY
