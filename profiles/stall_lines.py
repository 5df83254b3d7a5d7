#!/usr/bin/env python3
"""Source lines of a kernel ranked by one stall reason:  python profiles/stall_lines.py REP KERNEL stall_long_sb [top]"""
import csv, subprocess, sys
rep, kern, col = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; acc = {}; src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ic = hdr.index(col); continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0]); v = int(r[ic])
    except ValueError: continue
    acc[(cur, ln)] = acc.get((cur, ln), 0) + v; src[(cur, ln)] = r[1]
tot = sum(acc.values()) or 1
for (k, v) in sorted(acc.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100*v/tot:5.1f}%  {k[0]}:{k[1]}  {src[k][:110]}")
