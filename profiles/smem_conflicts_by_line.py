'''
Per-source-line shared-memory wavefronts (excessive = bank conflicts) and stall samples of one kernel; same join as
stalls_by_line.py.  Usage: python profiles/smem_conflicts_by_line.py X.ncu-rep KERNEL_REGEX MANGLED_SUBSTRING [lib.so]
'''
import csv, io, os, re, subprocess, sys, tempfile
rep, kregex, mangled = sys.argv[1:4]
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        'aircraft_trajectory_optimization_b200', 'libraceline_b200.so')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kregex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == 'Address':
        break
    data.append(r)
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
start = next(i for i, l in enumerate(dis) if l.startswith('_Z') and mangled in l and l.rstrip().endswith(':'))
lines, cur = [], None
for l in dis[start + 1:]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
    elif l.startswith('//-----') or l.startswith('\t.section'):
        break
assert len(lines) == len(data), (len(lines), len(data))
agg = {}
f = lambda r, k: int(float(r[ci[k]] or 0))
for ln, r in zip(lines, data):
    a = agg.setdefault(ln, dict(w=0, x=0, inst=0, n=0))
    a['w'] += f(r, 'L1 Wavefronts Shared')
    a['x'] += f(r, 'L1 Wavefronts Shared Excessive')
    a['inst'] += f(r, 'Instructions Executed')
    a['n'] += f(r, '# Samples')
tw, tx, tn = (sum(a[k] for a in agg.values()) for k in ('w', 'x', 'n'))
print(f'# {kregex}: shared-memory wavefronts {tw}, excessive {tx} ({100 * tx / max(tw, 1):.1f} %), {tn} stall samples')
for ln, a in sorted(agg.items(), key=lambda t: -t[1]['x'])[:25]:
    print(f'{ln[0]}:{ln[1]:<5d} wavefronts {a["w"]:>11d} excessive {a["x"]:>11d} ({100 * a["x"] / max(tx, 1):5.1f} % of all excessive)  '
          f'inst {a["inst"]:>10d}  samples {100 * a["n"] / max(tn, 1):4.1f} %')
