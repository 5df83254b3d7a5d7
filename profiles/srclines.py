#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on at capture time).

  python profiles/srclines.py gpurun_out/prof.ncu-rep k_encode_stream [min_share_pct]

Reads `ncu -i REP --page source --csv --print-source cuda,sass` and adds up, per (file, line), the
"Instructions Executed" and "# Samples" columns of the line's summary row.
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, hdr, first_fn = None, None, None
    acc = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if first_fn is None:
                first_fn = r[1]
            elif r[1] != first_fn:      # a second launch of the kernel: one is enough
                pass
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            i_thr = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            ln = int(r[0])
            inst = int(r[i_inst])
            samp = int(r[i_samp])
            thr = int(r[i_thr])
        except ValueError:
            continue
        a = acc.setdefault((cur_file, ln), [0, 0, 0, r[1]])
        a[0] += inst
        a[1] += samp
        a[2] += thr
    tot_i = sum(a[0] for a in acc.values()) or 1
    tot_s = sum(a[1] for a in acc.values()) or 1
    print(f"# {kern}: {tot_i} warp instructions, {tot_s} stall samples (all launches of the kernel in the report)")
    print("# file:line  inst%  samples%  lanes  source")
    for (f, ln), a in sorted(acc.items()):
        if 100.0 * a[0] / tot_i >= min_share or 100.0 * a[1] / tot_s >= min_share:
            lanes = a[2] / a[0] if a[0] else 0
            print(f"{f}:{ln:<5d} {100.0 * a[0] / tot_i:5.1f} {100.0 * a[1] / tot_s:5.1f} {lanes:5.1f}  {a[3].strip()[:120]}")


if __name__ == "__main__":
    main()
