#!/usr/bin/env python
"""SASS instruction count per source function (line ranges of the .cu file) of one kernel, from nvdisasm -g -c output.
usage: tools/code_size.py disasm.txt kernel-substring start:name,start:name,..."""
import collections, re, sys
txt, kname = sys.argv[1], sys.argv[2]
marks = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[3].split(','))
cnt = collections.Counter(); inside = False; cur = '(none)'
for line in open(txt):
    if line.startswith('//-----'):
        inside = ('.text.' in line and kname in line and (kname + '_') not in line.replace('_u8', '_X') if not kname.endswith('_u8') else ('.text.' in line and kname in line))
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        f, ln = m.group(1), int(m.group(2))
        if f.endswith('yk_analyze.cu'):
            cur = '(before)'
            for s0, nm in marks:
                if ln >= s0: cur = nm
        else:
            cur = 'hdr:' + f.split('/')[-1]
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): cnt[cur] += 1
tot = sum(cnt.values())
for nm, n in cnt.most_common():
    print(f"{nm:34s} {n:6d} instr {n*16/1024:6.1f} KB")
print("total", tot, tot * 16 / 1024, "KB")
