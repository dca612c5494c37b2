#!/usr/bin/env python
"""Opcode histogram and hot SASS segments of one kernel from an ncu report (run here, no GPU needed).
usage: tools/sass_hot.py gpurun_out/prof_TAG.ncu-rep kernel-regex"""
import csv, collections, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); ist = hdr.index('Warp Stall Sampling (All Samples)')
ops = collections.Counter(); stalls = collections.Counter(); tot = 0; data = []
for r in rows[2:]:
    try: n = int(r[ia]); s = int(r[ist])
    except Exception: continue
    txt = r[isrc].strip(); parts = txt.split()
    op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    ops[op] += n; stalls[op] += s; tot += n; data.append((n, s, txt))
print('total warp instructions', tot, 'sass lines', len(data), 'stall samples', sum(stalls.values()))
print(' '.join(f"{op}={n/tot:.3f}" for op, n in ops.most_common(24)))
seg = []; cur = None
for i, (n, s, t) in enumerate(data):
    if cur is None or abs(n - cur[2]) > 0.05 * max(n, cur[2], 1):
        if cur: seg.append(cur)
        cur = [i, i, n, n, s]
    else:
        cur[1] = i; cur[3] += n; cur[4] += s
seg.append(cur)
for c in sorted(sorted(seg, key=lambda c: -c[3])[:22]):
    print(f"sass {c[0]:5d}-{c[1]:5d} exec/inst {c[2]:9d} total {c[3]:10d} ({c[3]/tot:.3f}) stalls {c[4]:5d}  first: {data[c[0]][2][:56]}")
