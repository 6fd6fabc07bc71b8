#!/usr/bin/env python
"""Stall reasons per kernel source line from an .ncu-rep (source page) joined with nvdisasm line info.
usage: ncu_stalls.py report.ncu-rep mangled_kernel_name [top] [launch_index]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, fun = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lib = os.environ.get('B2D_LIBRARY') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'drone_b200', 'lib', 'libb200drone.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
dis = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.' + fun + ':'))
ins, stack, pend = [], [], []
for l in dis[start + 1:]:
    if l.startswith('//-----') or l.startswith('.text.') and not l.startswith('.text.' + fun):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pend.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if pend: stack, pend = pend, []
        ins.append((stack[-1] if stack else ('?', 0), m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci['# Samples'] and r[ci['# Samples']].isdigit()]
if len(data) != len(ins) and len(data) % len(ins) == 0:
    data = data[:len(ins)]  # several launches (or views) in the report: the first one
assert len(data) == len(ins), (len(data), len(ins))
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
per = collections.defaultdict(collections.Counter); tot = collections.Counter(); instr = collections.Counter()
for (line, txt), r in zip(ins, data):
    instr[line] += int(r[ci['Instructions Executed']] or 0)
    for h in reasons:
        v = int(r[ci[h]] or 0)
        per[line][h] += v; tot[h] += v
all_s = sum(tot.values()); all_i = sum(instr.values())
print('all samples', all_s, 'warp-instr', all_i)
print('by reason: ' + ', '.join(f'{h[6:]} {100.0 * v / all_s:.1f}%' for h, v in tot.most_common()))
print('--- lines by samples: samples%, instr%, top reasons')
for line, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(c.values())
    print(f'{line[0]}:{line[1]:<5d} {100.0 * s / all_s:6.2f}% {100.0 * instr[line] / all_i:6.2f}%  ' + ', '.join(f'{h[6:]} {100.0 * v / all_s:.1f}' for h, v in c.most_common(4) if v))
