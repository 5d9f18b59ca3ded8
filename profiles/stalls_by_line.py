'''
Per-source-line warp-stall samples of one kernel: joins the SASS page of an ncu report
(`ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME`) with `nvdisasm -g` line markers of the SAME build
(instruction order is identical).  Usage:
    python profiles/stalls_by_line.py X.ncu-rep KERNEL_REGEX MANGLED_SUBSTRING [libraceline_b200.so]
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
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg, tot = {}, 0
for ln, r in zip(lines, data):
    n = int(r[ci['# Samples']] or 0)
    tot += n
    a = agg.setdefault(ln, dict(n=0, inst=0))
    a['n'] += n
    a['inst'] += int(r[ci['Instructions Executed']] or 0)
    for h in stall_cols:
        v = int(r[ci[h]] or 0)
        if v:
            a[h[6:]] = a.get(h[6:], 0) + v
print(f'# {kregex}: {tot} samples over {len(data)} SASS instructions')
for ln, a in sorted(agg.items(), key=lambda t: -t[1]['n'])[:40]:
    st = sorted(((k, v) for k, v in a.items() if k not in ('n', 'inst')), key=lambda t: -t[1])[:4]
    print(f'{ln[0]}:{ln[1]:<5d} {100 * a["n"] / tot:5.1f}%  inst {a["inst"]:>10d}  ' + ' '.join(f'{k}={v}' for k, v in st))
