#!/usr/bin/env python
"""Join an .ncu-rep SASS source page with nvdisasm -gi line info of the in-tree library:
dynamic warp-instructions and stall samples per kernel source line (outermost inline frame).
usage: ncu_lines.py report.ncu-rep mangled_kernel_name [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, fun = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.environ.get('B2D_LIBRARY') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'drone_b200', 'lib', 'libb200drone.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
dis = subprocess.run(['nvdisasm', '-gi', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.' + fun + ':'))
ins = []  # (offset, outer_line, inner (file,line))
stack = []
pend = []
for l in dis[start + 1:]:
    if l.startswith('//-----') or l.startswith('.text.') and not l.startswith('.text.' + fun):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pend.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if pend:
            stack = pend
            pend = []
        ins.append((int(m.group(1), 16), stack[-1] if stack else ('?', 0), stack[0] if stack else ('?', 0), m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci['# Samples'] and r[ci['# Samples']].isdigit()]
if len(data) != len(ins) and len(data) % len(ins) == 0:
    data = data[:len(ins)]  # several launches (or views) in the report: the first one
assert len(data) == len(ins), (len(data), len(ins))
by_outer = collections.Counter(); by_inner = collections.Counter(); s_outer = collections.Counter(); s_inner = collections.Counter()
tot = 0
for (off, outer, inner, txt), r in zip(ins, data):
    n = int(r[ci['Instructions Executed']] or 0); s = int(r[ci['# Samples']])
    by_outer[outer] += n; by_inner[inner] += n; s_outer[outer] += s; s_inner[inner] += s; tot += n
print('total warp-instr', tot, 'samples', sum(s_outer.values()))
print('--- by kernel line (outermost frame): instr%, samples%')
for k, v in by_outer.most_common(top):
    print(f'{k[0]}:{k[1]:<5d} {100.0 * v / tot:6.2f}%  {100.0 * s_outer[k] / sum(s_outer.values()):6.2f}%')
print('--- by innermost frame')
for k, v in by_inner.most_common(top):
    print(f'{k[0]}:{k[1]:<5d} {100.0 * v / tot:6.2f}%  {100.0 * s_inner[k] / sum(s_inner.values()):6.2f}%')
print('--- by kernel line, sorted by samples: samples%, instr%')
ts = sum(s_outer.values())
for k, v in s_outer.most_common(top):
    print(f'{k[0]}:{k[1]:<5d} {100.0 * v / ts:6.2f}%  {100.0 * by_outer[k] / tot:6.2f}%')
