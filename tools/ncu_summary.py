#!/usr/bin/env python
"""Summarise an ncu report here (no GPU needed): per-kernel headline metrics, and the hottest source lines of one kernel.
usage: tools/ncu_summary.py gpurun_out/prof_TAG.ncu-rep [kernel-regex-for-source-view]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size']
idx = [hdr.index(w) for w in want if w in hdr]
print(",".join(hdr[i] for i in idx)); print(",".join(rows[1][i] for i in idx))
seen = set()
for r in rows[2:]:
    if r[idx[0]] in seen: continue
    seen.add(r[idx[0]]); print(",".join(r[i].split('(')[0] if i == idx[0] else r[i] for i in idx))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    lines = []
    for r in rows[3:]:
        if r and r[0] != '':
            try: lines.append((int(r[0]), r[1].strip()[:100], int(r[4]), int(r[7])))
            except Exception: pass
    ts = sum(l[2] for l in lines) or 1; ti = sum(l[3] for l in lines) or 1
    print(f"# source view of {sys.argv[2]}: total stall samples {ts}, warp instructions {ti}")
    for key, name in ((3, "instructions"), (2, "stall samples")):
        print(f"# top lines by {name}")
        for l in sorted(lines, key=lambda l: -l[key])[:22]:
            print(f"{l[0]:5d} inst={l[3]/ti:6.3f} stall={l[2]/ts:6.3f}  {l[1]}")
