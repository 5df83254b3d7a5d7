"""Synthetic VCF generators for tests and bench (test infrastructure, not product).

Two generators, both deterministic in (shape, seed):

* ``random_vcf_like``  -- the line grammar and allele distribution of the reference's
  ``other/random_vcf.py`` (/root/reference/other/random_vcf.py:36-72): CHROM ``1``, POS
  10000+2i, ID ``var<i>``, REF/ALT bases, ``100 PASS INFO GT`` and phased diploid genotypes
  with iid alleles P(0,1,2) = (.90,.08,.02).  It uses numpy's RNG, not Python's
  ``random`` stream, so bytes differ from the reference script's output for the same
  seed; the byte-exact reference stream is only needed for tests/golden (see
  oracle/make_golden.py, which executes the reference script itself).
* ``kg_like`` -- "1000 Genomes chr20-shaped" lines: per-line
  alt-allele frequency from a 1/x site-frequency spectrum clipped to [1/5008, 0.5],
  ~1% multi-allelic lines, and a 150-250 byte INFO column.

Both return ``(header_bytes, data_bytes)``; ``data_bytes`` is the '\\n'-terminated
data-line region that the block C-ABI takes.
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ATGC", dtype=np.uint8)


def header(n_samples: int, prefix: str = "HG") -> bytes:
    digits = max(1, len(str(max(n_samples - 1, 0))))
    names = "\t".join(f"{prefix}{j:0{digits}d}" for j in range(n_samples))
    cols = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT"
    if n_samples:
        cols += "\t" + names
    return (
        "##fileformat=VCFv4.1\n"
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
        "##fileDate=20150218\n" + cols + "\n"
    ).encode()


def _assemble(prefixes: list[bytes], gt: np.ndarray) -> bytes:
    """prefixes[i] = required columns incl. trailing tab; gt = uint8 [L, S, 4] (a,'|',b,sep)."""
    n_lines, n_samples, _ = gt.shape
    body = gt.reshape(n_lines, n_samples * 4)
    body[:, -1] = ord("\n")
    lens = np.fromiter((len(p) for p in prefixes), dtype=np.int64, count=n_lines)
    row = n_samples * 4
    starts = np.zeros(n_lines + 1, dtype=np.int64)
    np.cumsum(lens + row, out=starts[1:])
    out = np.empty(int(starts[-1]), dtype=np.uint8)
    for i, p in enumerate(prefixes):
        s = int(starts[i])
        out[s:s + len(p)] = np.frombuffer(p, dtype=np.uint8)
    # scatter genotype rows
    idx = (starts[:-1] + lens)[:, None] + np.arange(row, dtype=np.int64)[None, :]
    out[idx.reshape(-1)] = body.reshape(-1)
    return out.tobytes()


def _gt_block(a1: np.ndarray, a2: np.ndarray, phased: bool = True) -> np.ndarray:
    n_lines, n_samples = a1.shape
    gt = np.empty((n_lines, n_samples, 4), dtype=np.uint8)
    gt[:, :, 0] = a1 + ord("0")
    gt[:, :, 1] = ord("|") if phased else ord("/")
    gt[:, :, 2] = a2 + ord("0")
    gt[:, :, 3] = ord("\t")
    return gt


def random_vcf_like(n_lines: int, n_samples: int, seed: int = 5,
                    probs=(0.90, 0.08, 0.02), first_line: int = 0):
    rng = np.random.default_rng([seed, n_samples, first_line])
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
    cdf /= cdf[-1]
    a = np.searchsorted(cdf, rng.random((n_lines, n_samples, 2)), side="right").astype(np.uint8)
    a = np.minimum(a, len(probs) - 1)
    gt = _gt_block(a[:, :, 0], a[:, :, 1])
    prefixes = []
    for i in range(first_line, first_line + n_lines):
        perm = rng.permutation(4)
        ref = chr(_BASES[perm[0]])
        alts = ",".join(chr(_BASES[k]) for k in perm[1:3])
        prefixes.append(f"1\t{10000 + 2 * i}\tvar{i}\t{ref}\t{alts}\t100\tPASS\tINFO\tGT\t".encode())
    return header(n_samples), _assemble(prefixes, gt)


def kg_like(n_lines: int, n_samples: int = 2504, seed: int = 20, first_line: int = 0,
            chrom: str = "20"):
    rng = np.random.default_rng([seed, n_samples, first_line, 1000])
    # 1/x site-frequency spectrum on [1/(2S), 0.5]: inverse-CDF sampling of p(x) ~ 1/x
    lo, hi = 1.0 / (2 * n_samples), 0.5
    af = lo * (hi / lo) ** rng.random(n_lines)
    multi = rng.random(n_lines) < 0.01
    u = rng.random((n_lines, n_samples, 2))
    a = (u < af[:, None, None]).astype(np.uint8)
    # multi-allelic lines: a share of the alt alleles become allele 2
    two = (rng.random((n_lines, n_samples, 2)) < 0.3) & multi[:, None, None]
    a = np.where((a == 1) & two, 2, a).astype(np.uint8)
    gt = _gt_block(a[:, :, 0], a[:, :, 1])
    ac = a.reshape(n_lines, -1)
    ac1 = (ac == 1).sum(axis=1)
    an = 2 * n_samples
    pops = ["EAS", "AMR", "AFR", "EUR", "SAS"]
    prefixes = []
    pos = 60000 + 35 * first_line
    for k in range(n_lines):
        i = first_line + k
        pos += 1 + int(rng.integers(1, 70))
        perm = rng.permutation(4)
        ref = chr(_BASES[perm[0]])
        alt = chr(_BASES[perm[1]]) + ("," + chr(_BASES[perm[2]]) if multi[k] else "")
        f = ac1[k] / an
        popaf = ";".join(f"{p}_AF={max(0.0, f * (0.5 + rng.random())):.4f}" for p in pops)
        info = (f"AC={ac1[k]};AF={f:.6f};AN={an};NS={n_samples};DP={int(rng.integers(8000, 30000))};"
                f"{popaf};AA={ref}|||;VT=SNP" + (";MULTI_ALLELIC" if multi[k] else ""))
        if rng.random() < 0.3:
            info += ";EX_TARGET"
        prefixes.append(f"{chrom}\t{pos}\trs{100000 + i}\t{ref}\t{alt}\t100\tPASS\t{info}\tGT\t".encode())
    return header(n_samples), _assemble(prefixes, gt)


def edge_case_lines(n_samples: int = 8) -> list[bytes]:
    """Data lines covering SURVEY.md 8(a)'s irregular-input table (all decodable)."""
    req = b"1\t%d\trs%d\tA\tT,G\t50\tPASS\tAC=1;AF=0.5\tGT"
    pos = [100]

    def line(samples, fmt=None):
        r = req % (pos[0], pos[0])
        pos[0] += 2
        if fmt is not None:
            r = r[:-2] + fmt
        return r + b"\t" + b"\t".join(samples) + b"\n"

    s = n_samples
    out = [
        line([b"0|0"] * 3 + [b"0|1", b"0|1", b"1|0", b"1|1", b"0|0"][: max(0, s - 3)] + [b"0|0"] * max(0, s - 8)),
        line(([b"0|2", b"0|0", b"./.", b"0/0", b"0/1", b"2|2", b"1|1", b"2|1"] * ((s + 7) // 8))[:s]),
        line([b"0|0:3", b"0|0:3", b"0|1:9", b"0|0:3"] * (s // 4) + [b"1|1:7"] * (s % 4), fmt=b"GT:DP"),
        line([b"0|0"] * s),
        line([b"1|1"] * s),
        line(([b"0|1", b"1|0"] * s)[:s]),
        line([b"."] * s),
        line(([b"0", b"1", b"0|0|0", b"10|0", b"0|10"] * s)[:s]),
        line([b"0|0"] * (s - 1) + [b"0|2"]),
        line([b"2|0"] + [b"0|0"] * (s - 1)),
        line([b"0|0:12:99:0,36,400"] * s, fmt=b"GT:DP:GQ:PL"),
    ]
    return out


def run_length_lines(lengths=(1, 30, 31, 32, 62, 63, 126, 127, 128, 254, 255, 300), n_samples: int = 700):
    """Lines whose runs hit the 31 / 127 chunk limits exactly."""
    gts = [b"0|0", b"0|1", b"1|0", b"1|1"]
    lines = []
    pos = 5000
    for g_i, g in enumerate(gts):
        for n in lengths:
            other = gts[(g_i + 1) % 4]
            samples = []
            while len(samples) < n_samples:
                samples += [g] * n + [other]
            samples = samples[:n_samples]
            lines.append(b"2\t%d\t.\tC\tG\t.\t.\t.\tGT\t" % pos + b"\t".join(samples) + b"\n")
            pos += 1
    return lines


def odd_mix_lines(rng, n_samples: int, n_lines: int):
    """Data lines whose sample columns are a random mix, drawn per line, of coded genotypes (in runs around the 31 / 127 chunk
    limits), 3-byte literals, haploid calls, GT:DP[:GQ] columns, multi-digit alleles and (rarely) literals of kilobytes."""
    coded = (b"0|0", b"0|0", b"0|0", b"0|1", b"1|0", b"1|1")
    lit3 = (b"./.", b"0/0", b"0/1", b"2|0", b"1|2", b".|.")

    def term(kind):
        if kind == 0: return rng.choice(coded)
        if kind == 1: return rng.choice(lit3)
        if kind == 2: return bytes(rng.choice(b"01.2") for _ in range(rng.choice((1, 1, 1, 2, 2))))       # haploid calls, "12"
        if kind == 3: return rng.choice(coded + lit3) + b":" + b":".join(str(rng.randrange(10 ** rng.randrange(1, 4))).encode() for _ in range(rng.randrange(1, 4)))
        if kind == 4: return rng.choice((b"10|0", b"0|10", b"0|0|0", b"11|12"))
        return b"0|1:" + bytes(rng.choice(b"0123456789,") for _ in range(rng.choice((300, 2040, 2100, 5000))))

    lines = []
    for i in range(n_lines):
        weights = [rng.choice((0, 1, 8, 40)), rng.choice((0, 0, 1, 5)), rng.choice((0, 0, 3, 30)), rng.choice((0, 0, 3, 30)),
                   rng.choice((0, 0, 1)), rng.choice((0, 0, 0, 0.02))]
        if not any(weights): weights[0] = 1
        terms = []
        while len(terms) < n_samples:
            kind = rng.choices(range(6), weights)[0]
            if kind == 0 and rng.random() < 0.5:
                terms += [rng.choice(coded)] * rng.choice((1, 2, 30, 31, 32, 126, 127, 128, 300))      # runs at the chunk limits
            else:
                terms += [term(kind) for _ in range(rng.choice((1, 1, 3, 9, 40)))]
        lines.append(b"%d\t%d\t.\tA\tC\t.\tPASS\tDP=%d\tGT\t" % (rng.randrange(1, 23), 100 + i, i) + b"\t".join(terms[:n_samples]) + b"\n")
    return lines
