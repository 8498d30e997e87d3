#!/usr/bin/env python
"""Attribute ncu SASS-level samples to CUDA source lines: ncu_lines.py <rep> <cubin> <mangled-kernel-substring> [N]"""
import re, csv, subprocess, io, collections, sys, os
rep, cubin, kname = sys.argv[1:4]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(txt) if l.startswith('.text.') and kname in l][0]
cur = None; locs = []
for l in txt[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-------------'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): locs.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(locs) == len(data), (len(locs), len(data))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for loc, r in zip(locs, data):
    a = agg[loc]; a[0] += int(r[ix['# Samples']]); a[1] += int(r[ix['Instructions Executed']])
    for s in stall_cols: a[2][s] += int(r[ix[s]] or 0)
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
files = {}
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ramses_b200', 'csrc')
def srcline(loc):
    f, l = loc
    if f not in files:
        try: files[f] = open(os.path.join(root, f)).read().split('\n')
        except Exception: files[f] = None
    return files[f][l - 1].strip()[:100] if files[f] and l - 1 < len(files[f]) else ''
print(f"{len(data)} SASS instructions; columns: %stall-samples %warp-instr file:line [top stalls] source")
for loc, a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    top = ','.join(f"{k[6:]}:{v}" for k, v in a[2].most_common(3))
    print(f"{a[0]/ts*100:5.2f}%s {a[1]/ti*100:5.2f}%i {loc[0]}:{loc[1]:4d} [{top}] {srcline(loc)}")
