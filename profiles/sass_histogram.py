'''
Static SASS opcode histogram of the built library (no GPU needed):
    python profiles/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt
'''
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'aircraft_trajectory_optimization_b200', 'libraceline_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
op_re = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)')
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = op_re.match(line)
    if m and cur is not None:
        cur[m.group(1)] += 1
total = collections.Counter()
for c in per.values():
    total.update(c)
fmt = lambda c, n=18: ', '.join(f'{k} {v}' for k, v in c.most_common(n))
print('# SASS opcode histogram (cuobjdump -sass libraceline_b200.so, sm_100a), static instruction counts per kernel')
print('# whole library:', fmt(total, 17) + ',')
keys = ['UBLKCP', 'SYNCS', 'LDGSTS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'BAR', 'UTCHMMA', 'UTCQMMA', 'DMMA', 'HMMA']
print('# async / tensor-related mnemonics in the library:', {k: total[k] for k in keys if total[k]})
for name, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    print(f'\n== {name}  ({sum(c.values())} instructions)')
    print('   ' + fmt(c))
