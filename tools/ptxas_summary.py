#!/usr/bin/env python3
"""registers / spills / shared memory per kernel from build/<file>.log (-Xptxas -v):
   python tools/ptxas_summary.py particles [substring]"""
import re, subprocess, sys
name = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ''
txt = open(f'build/{name}.log').read().splitlines()
cur = None
for i, l in enumerate(txt):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(anonymous namespace\)::', '', cur).split('(')[0].replace('void ', '')
        continue
    if cur and 'Used' in l and pat in cur:
        spill = re.search(r'(\d+) bytes spill stores', txt[i - 1])
        regs = re.search(r'Used (\d+) registers', l).group(1)
        smem = re.search(r'(\d+) bytes smem', l)
        print(f'{cur:70s} regs {regs:>3s}  spill {spill.group(1) if spill else "?":>4s}  smem {smem.group(1) if smem else "0":>6s}')
        cur = None
