#!/usr/bin/env python3
"""Turns gpurun_out/ captures into the small, tracked summaries under profiles/.

    python profiles/summarize.py <round-tag> <launches.csv> <bench.json> <full.ncu-rep> [<full.ncu-rep> ...]

* launches: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ ...` over bench.py
  -> per-kernel launch count, total time and SHARE of the step (cold-cache, serialised: shares, not absolutes)
* full: `ncu --set full ... -k regex:"k_encode_stream|k_dec_expand_grid"` -> DRAM bytes, throughput %, stall picture
"""
import collections
import csv
import json
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].replace("void ", "").split("(")[0].split("<")[0].replace("vcfc::", "")     # (k_encode_stream<(bool)0> -> k_encode_stream)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
    return "\n".join(out), len(rows)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")].split("(")[0]
        out.append(f"\n### `{name}`\n\n| metric | unit | value |\n|---|---|---:|")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                out.append(f"| {h} | {u} | {v} |")
    return "\n".join(out)


if __name__ == "__main__":
    tag, lpath, bpath = sys.argv[1:4]
    md = [f"# ncu summaries, {tag}\n"]
    t, n = launches(lpath)
    md.append(f"## Launch list ({n} launches; `--metrics gpu__time_duration.sum --clock-control none`)\n\n{t}\n")
    md.append("## `--set full` captures of the dominant kernels (400k-line run of the bench workload: 4.07 GB of text, 0.238 GB of .vcfc)")
    for fpath in sys.argv[4:]:
        md.append(full(fpath))
    md.append("\n## bench.py line of the same build\n\n```json\n" + json.dumps(json.load(open(bpath)), indent=1) + "\n```\n")
    open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(md))
    print(f"profiles/{tag}_ncu_summary.md")
