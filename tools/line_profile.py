#!/usr/bin/env python3
"""Executed warp-instructions and stall samples per CUDA source line of one kernel.
Joins `nvdisasm -g -c <cubin>` (line info per SASS instruction, in program order) with an
`ncu --page source --csv --print-source sass` export (executed counts per SASS instruction).

  python tools/line_profile.py <cubin> <kernel substring> <sass.csv> [section index]
"""
import csv, re, subprocess, sys, collections
cubin, kname, sass_csv = sys.argv[1:4]
sec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(txt) if l.startswith('.text.') and kname in l and l.rstrip().endswith(':'))
lines = []; cur = None
for l in txt[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-----'):
        if lines: break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and kname.replace('ILi', '<').split('<')[0] in r[1]]
s0 = secs[sec]; nxt = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and i > s0]
s1 = nxt[0] if nxt else len(rows)
hdr = rows[s0 + 1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[s0 + 2:s1] if len(r) == len(hdr)]
print('kernel:', rows[s0][1][:80]); print('sass instrs: nvdisasm', len(lines), 'ncu', len(body))
ex = collections.Counter(); sm = collections.Counter()
for ln, r in zip(lines, body):
    ex[ln] += int(r[idx['Instructions Executed']] or 0); sm[ln] += int(r[idx['# Samples']] or 0)
te = sum(ex.values()); ts = sum(sm.values()) or 1
src = {}
for ln, v in sorted(ex.items(), key=lambda kv: -kv[1])[:45]:
    f, n = ln if ln else ('?', 0)
    if f not in src:
        try: src[f] = open('/root/repo/hipace_b200/csrc/' + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][n - 1].strip() if 0 < n <= len(src[f]) else ''
    print(f"{f}:{n:4d} exec {100*v/te:5.1f}%  samples {100*sm[ln]/ts:5.1f}%  {text[:90]}")
