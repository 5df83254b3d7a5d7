#!/usr/bin/env python3
"""Warp instructions per source region of a kernel:  python profiles/regions.py REP KERNEL FILE 'name:lo-hi,name:lo-hi,...' [divisor]"""
import csv, subprocess, sys
rep, kern, fname, spec = sys.argv[1:5]
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; acc = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0]); inst = int(r[ii])
    except ValueError: continue
    acc[(cur, ln)] = acc.get((cur, ln), 0) + inst
tot = sum(acc.values())
for item in spec.split(','):
    name, rng = item.split(':'); lo, hi = map(int, rng.split('-'))
    s = sum(v for (f, l), v in acc.items() if f == fname and lo <= l <= hi)
    print(f"{name:28s} {100*s/tot:5.1f}%  {s/div:9.1f}")
oth = {}
for (f, l), v in acc.items():
    if f != fname: oth[f] = oth.get(f, 0) + v
for k, v in oth.items(): print(f"{k:28s} {100*v/tot:5.1f}%  {v/div:9.1f}")
print("total", tot, tot / div)
