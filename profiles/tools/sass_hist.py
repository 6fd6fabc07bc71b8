#!/usr/bin/env python
"""SASS op-histogram per kernel of the in-tree library (cuobjdump -sass): what the hot kernels are made of.
usage: sass_hist.py [library] > profiles/<tag>_sass_histograms.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'drone_b200', 'lib', 'libb200drone.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fun, hist = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        fun = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[fun] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', l)
    if m and fun:
        hist[fun][m.group(1)] += 1
KEY = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCATOMSWS', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'HMMA', 'FFMA2', 'FMUL2', 'FADD2',
       'FFMA', 'MUFU', 'DMUL', 'DADD', 'DFMA', 'ATOMS', 'ATOMG', 'RED', 'REDUX', 'BAR', 'FCHK', 'CALL', 'STL', 'LDL']
for f, h in hist.items():
    tot = sum(h.values())
    print(f'== {f[:110]}: {tot} instructions')
    print('   marker ops: ' + ', '.join(f'{k} {h[k]}' for k in KEY if h[k]))
    print('   top: ' + ', '.join(f'{k} {v}' for k, v in h.most_common(14)))
