#!/usr/bin/env python
"""Per-source-line warp instructions of one kernel from an ncu report captured with --import-source on (-lineinfo build).
usage: tools/src_hot.py report.ncu-rep kernel-regex [top]"""
import csv, collections, io, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for i, r in enumerate(rows):
    if 'Instructions Executed' in r:
        hdr = r; start = i + 1; break
iln = hdr.index('Line No'); isrc = hdr.index('Source'); ia = hdr.index('Instructions Executed'); ist = hdr.index('Warp Stall Sampling (All Samples)')
per = collections.Counter(); stl = collections.Counter(); text = {}; tot = 0
for r in rows[start:]:
    try: n = int(r[ia]); s = int(r[ist])
    except Exception: continue
    key = r[iln]; per[key] += n; stl[key] += s; text.setdefault(key, r[isrc].strip()); tot += n
print("total", tot)
for k, n in per.most_common(top):
    print(f"{n:9d} {n/tot:6.3f} stalls {stl[k]:5d}  L{k:>5}: {text[k][:130]}")
# per-function totals when a 4th argument gives "start:name,start:name,..." (line ranges)
if len(sys.argv) > 4:
    marks = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[4].split(','))
    ft = collections.Counter()
    for k, n in per.items():
        if not k.strip().isdigit(): ft['(no line)'] += n; continue
        ln = int(k); name = '(before)'
        for s0, nm in marks:
            if ln >= s0: name = nm
        ft[name] += n
    real = tot - ft['(no line)']
    for nm, n in ft.most_common():
        print(f"{nm:24s} {n:10d} {n/real:6.3f}")
