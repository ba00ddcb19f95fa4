#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source sass` export: hottest instructions by
warp-stall samples and the share of samples per 100-instruction segment."""
import csv, sys
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != 'Address']
def S(r, k):
    try: return int(r[idx[k]] or 0)
    except ValueError: return 0
tot = sum(S(r, '# Samples') for r in body)
print('total samples', tot, 'instructions', len(body))
for k, r in enumerate(body):
    s = S(r, '# Samples')
    if s > tot * thr / 100:
        print(f"{k:5d} {100*s/tot:5.1f}% {r[idx['Source']][:100]:100s} lsb={r[idx['stall_long_sb']]} ex={r[idx['Instructions Executed']]}")
for a in range(0, len(body), 100):
    s = sum(S(r, '# Samples') for r in body[a:a + 100])
    print(f"seg {a:5d}: {100*s/tot:5.1f}%")
