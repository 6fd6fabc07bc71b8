#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep source page (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sc, src, ie = ci['# Samples'], ci['Source'], ci['Instructions Executed']
data = [r for r in rows[2:] if len(r) > sc and r[sc].isdigit()]
tot = sum(int(r[sc]) for r in data)
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]]) for r in data if r[ci[h]].isdigit()) for h in stall}
print('total samples', tot, 'total warp-instr', sum(int(r[ie]) for r in data if r[ie].isdigit()))
print(sorted(agg.items(), key=lambda x: -x[1])[:10])
order = sorted(range(len(data)), key=lambda i: -int(data[i][sc]))
for i in order[:topn]:
    r = data[i]
    st = {h[6:]: int(r[ci[h]]) for h in stall if r[ci[h]].isdigit() and int(r[ci[h]]) > 0.2 * int(r[sc])}
    prev = data[i - 1][src][:50].strip() if i > 0 else ''
    print(r[sc], r[ie], '|', prev, '|', r[src][:70].strip(), st)
