#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export: per kernel duration, DRAM bytes, throughput
percentages -> markdown table on stdout, and (with --traffic out.json) the per-launch
dram__bytes_read.sum + dram__bytes_write.sum of the particle kernels for bench.py's `roofline.traffic`.

  python tools/ncu_summary.py profiles/r01_raw.csv --traffic profiles/ncu_traffic.json --nxy 1024 --ppc 4
"""
import argparse, csv, json, re

ap = argparse.ArgumentParser()
ap.add_argument('raw'); ap.add_argument('--traffic'); ap.add_argument('--nxy', type=int, default=1024)
ap.add_argument('--ppc', type=int, default=4); ap.add_argument('--source', default=None)
a = ap.parse_args()
rd = csv.reader(open(a.raw)); hdr = next(rd); units = next(rd); rows = list(rd)
ix = {h: i for i, h in enumerate(hdr)}
U = {h: u for h, u in zip(hdr, units)}
def val(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return float('nan')
def to_bytes(r, k):
    v = val(r, k); u = U[k].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
def to_us(r, k):
    v = val(r, k); u = U[k].lower()
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1)   # ncu 2025: 'us'
agg = {}
for r in rows:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('<unnamed>::', '').replace('void ', '')
    key = (name, r[ix['Grid Size']])
    d = agg.setdefault(key, dict(n=0, us=0., rd=0., wr=0., lts=0., sm=0., fp64=0., occ=0., regs=0))
    d['n'] += 1; d['us'] += to_us(r, 'gpu__time_duration.sum')
    d['rd'] += to_bytes(r, 'dram__bytes_read.sum'); d['wr'] += to_bytes(r, 'dram__bytes_write.sum')
    d['lts'] += val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
    d['sm'] += val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')
    d['fp64'] += val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')
    d['occ'] += val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')
    d['regs'] = int(val(r, 'launch__registers_per_thread'))
print('| kernel | grid | launches | avg us | DRAM read MB | DRAM write MB | DRAM GB/s | L2 % | SM % | fp64 pipe % | warps active % | regs |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|')
tot = sum(d['us'] for d in agg.values())
for (name, grid), d in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    n = d['n']; us = d['us'] / n
    gbs = (d['rd'] + d['wr']) / n / (us * 1e-6) / 1e9 if us > 0 else 0
    print(f"| `{name}` | {grid} | {n} | {us:.1f} | {d['rd']/n/1e6:.1f} | {d['wr']/n/1e6:.1f} | {gbs:.0f} | "
          f"{d['lts']/n:.0f} | {d['sm']/n:.0f} | {d['fp64']/n:.0f} | {d['occ']/n:.0f} | {d['regs']} |")
print(f'\nsum of kernel durations: {tot:.1f} us over {sum(d["n"] for d in agg.values())} launches')
if a.traffic:
    k = {}
    for (name, grid), d in agg.items():
        base = re.sub(r'<.*', '', name)
        for stem in ('k_advance_plasma', 'k_explicit_deposition', 'k_deposit_current'):
            if base.startswith(stem):          # k_advance_plasma_row / _cta: the push of the current default
                k[stem] = (d['rd'] + d['wr']) / d['n']
    json.dump({'nxy': a.nxy, 'ppc': a.ppc, 'source': a.source or a.raw, 'unit': 'bytes per launch',
               'kernels': k}, open(a.traffic, 'w'), indent=1)
