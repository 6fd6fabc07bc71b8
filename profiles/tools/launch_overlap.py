#!/usr/bin/env python
"""Summarise a B2D_TRACE_FILE (library built with -DB2D_EXPERIMENT_TIMING=1): per-CTA (entry, go, done) globaltimer
stamps of the LAST TWO race_step_kernel launches of a handle.  Shows what CUDA-event averages only imply: CTAs of
launch t+1 are resident and running while launch t drains (programmatic dependent launch + per-CTA chain flags)."""
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
seqs = sorted({int(r["launch_seq"]) for r in rows})
L = {s: [r for r in rows if int(r["launch_seq"]) == s and int(r["done_ns"]) > 0] for s in seqs}
a, b = L[seqs[0]], L[seqs[1]]
f = lambda rs, k: [int(r[k]) for r in rs]  # noqa: E731
t0 = min(f(a, "entry_ns"))
a_done_last, a_done_first = max(f(a, "done_ns")), min(f(a, "done_ns"))
b_go = f(b, "go_ns")
print(f"launch {seqs[0]}: {len(a)} CTAs, first entry 0.0 us, first done {(a_done_first - t0) / 1e3:.1f} us, last done {(a_done_last - t0) / 1e3:.1f} us")
print(f"launch {seqs[1]}: {len(b)} CTAs, first entry {(min(f(b, 'entry_ns')) - t0) / 1e3:.1f} us, first tile started {(min(b_go) - t0) / 1e3:.1f} us, "
      f"last done {(max(f(b, 'done_ns')) - t0) / 1e3:.1f} us")
early = sum(1 for g in b_go if g < a_done_last)
print(f"CTAs of launch {seqs[1]} that began their first tile before the last CTA of launch {seqs[0]} finished: {early} of {len(b)} "
      f"({100.0 * early / len(b):.1f} %)")
print(f"overlap window (last done of {seqs[0]} - first tile of {seqs[1]}): {(a_done_last - min(b_go)) / 1e3:.1f} us")
print(f"launch period (last done to last done): {(max(f(b, 'done_ns')) - a_done_last) / 1e3:.1f} us")
