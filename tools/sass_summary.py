"""per-kernel SASS instruction counts of the built library (cuobjdump -sass): which kernels are tcgen05 / TMEM / bulk-copy code.
usage: python tools/sass_summary.py > profiles/<round>_sass_summary.txt"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'poco_b200', 'libpoco_b200.so')
OPS = ['UTCHMMA', 'LDTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'UTCBAR', 'SYNCS', 'STG.E.128', 'LDG.E.128']
sass = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
chunks = re.split(r'\n\s*Function : \S+\n', sass)[1:]
sha = hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]
print(f'# cuobjdump -sass poco_b200/libpoco_b200.so (sha256 {sha}): instruction counts per kernel (sm_100a)')
print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D bulk copy, TMA engine), UTMALDG/UTMASTG = tensor-map TMA '
      '(none: the planar layout needs no tensor maps), LDGSTS = cp.async, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops')
print('kernel,' + ','.join(OPS))
rows = []
for name, body in zip(names, chunks):
    short = name.replace('poco::(anonymous namespace)::', '').replace('<unnamed>::', '').replace('poco::', '')
    depth, cut = 0, len(short)          # cut the argument list: the first '(' outside template brackets
    for i, ch in enumerate(short):
        depth += ch == '<'
        depth -= ch == '>'
        if ch == '(' and depth == 0:
            cut = i
            break
    short = short[:cut].strip()
    rows.append((short, [len(re.findall(r'\b' + re.escape(op) + r'\b', body)) for op in OPS]))
for short, counts in sorted(rows, key=lambda r: (-r[1][0], r[0])):
    print(f'"{short}",' + ','.join(str(c) for c in counts))
