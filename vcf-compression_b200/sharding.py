"""Line-block sharding of a data-line region across ranks (host logic, no GPU needed).

The path partitions by blocks of data lines: rank g encodes a newline-aligned byte
range, outputs are concatenated in rank order and every shard's line offsets are rebased by the bytes
that precede it.  No collective is on the data path; `gather_concat` is the host-side assembly.
"""
from __future__ import annotations


def split_ranges(data, n_shards: int):
    """Newline-aligned, contiguous, near-equal byte ranges [(start, end)] covering `data` (len == n_shards)."""
    n = len(data)
    cuts = [0]
    for g in range(1, n_shards):
        target = max(cuts[-1], (n * g) // n_shards)
        if target >= n:
            cuts.append(n)
            continue
        if target == 0 or data[target - 1:target] == b"\n":
            cuts.append(target)
            continue
        nl = data.find(b"\n", target)
        cuts.append(n if nl < 0 else nl + 1)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


def gather_concat(parts):
    """parts: per rank (encoded bytes, n_lines, [line offsets]) in rank order -> (bytes, n_lines, offsets)."""
    out = bytearray()
    offs, lines = [], 0
    for enc, nl, lo in parts:
        offs.extend(o + len(out) for o in lo)
        out += enc
        lines += nl
    return bytes(out), lines, offs
