"""Device-side synthetic VCF generator for bench.py and the full-size GPU tests (test/bench
infrastructure, not product).  Pure torch ops, so the same code runs on cuda (bench) and on cpu
(the CPU suite checks it against the oracle's tokeniser on small shapes).

Shapes follow BASELINE.json / SURVEY.md 8(d):

* ``kg``      config 2, "1000 Genomes chr20-shaped": per-line alt-allele frequency from a 1/x
              site-frequency spectrum clipped to [1/(2S), 0.5], phased diploid, ~1 % multi-allelic
              lines (allele 2), a 150-200 byte INFO column of varying width.
* ``random``  the distribution of the reference's other/random_vcf.py:36-72: iid alleles with
              P(0,1,2) = (.90,.08,.02), short INFO.

Lines are generated in chunks; every chunk is a deterministic function of (kind, seed, first line,
sample count), so any slice can be regenerated on the host for oracle comparison.
Line grammar: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT(GT) then S genotypes ``a|b``.
"""
from __future__ import annotations

import torch

_PAD = 0          # filler byte removed when rows are compacted


def _digits(x: torch.Tensor, width: int, pad_leading: bool) -> torch.Tensor:
    """int64 [L] -> uint8 [L, width] decimal digits; leading zeros become _PAD when pad_leading."""
    pw = 10 ** torch.arange(width - 1, -1, -1, device=x.device, dtype=torch.int64)
    d = (x[:, None] // pw[None, :]) % 10
    out = (d + 48).to(torch.uint8)
    if pad_leading:
        lead = (x[:, None] < pw[None, :]) & (torch.arange(width, device=x.device)[None, :] < width - 1)
        out = torch.where(lead, torch.full_like(out, _PAD), out)
    return out


def _lit(s: str, n: int, device) -> torch.Tensor:
    t = torch.tensor(list(s.encode()), dtype=torch.uint8, device=device)
    return t[None, :].expand(n, -1)


def _chunk(kind: str, seed: int, first_line: int, n_lines: int, n_samples: int, device, chrom: str = "20"):
    """Returns (flat uint8 bytes of n_lines data lines, int64 line lengths)."""
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1_000_003 + first_line) & 0x7FFFFFFFFFFF)
    L, S = n_lines, n_samples
    idx = torch.arange(first_line, first_line + L, device=device, dtype=torch.int64)
    u = torch.rand((L, 8), generator=g, device=device)
    bases = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ref_i = (u[:, 0] * 4).long().clamp(max=3)
    alt_i = (ref_i + 1 + (u[:, 1] * 3).long().clamp(max=2)) % 4
    alt2_i = (alt_i + 1 + (ref_i == (alt_i + 1) % 4).long()) % 4
    if kind == "kg":
        lo, hi = 1.0 / (2 * S), 0.5
        af = lo * (hi / lo) ** u[:, 2].double()
        multi = u[:, 3] < 0.01
        r = torch.rand((L, S, 2), generator=g, device=device)
        a = (r < af[:, None, None].float()).to(torch.uint8)
        two = (torch.rand((L, S, 2), generator=g, device=device) < 0.3) & multi[:, None, None]
        a = torch.where((a == 1) & two, torch.full_like(a, 2), a)
        del r, two
    else:
        multi = torch.ones(L, dtype=torch.bool, device=device)
        r = torch.rand((L, S, 2), generator=g, device=device)
        a = (r >= 0.90).to(torch.uint8) + (r >= 0.98).to(torch.uint8)
        del r
    gt = torch.empty((L, S, 4), dtype=torch.uint8, device=device)
    gt[:, :, 0] = a[:, :, 0] + 48
    gt[:, :, 1] = ord("|")
    gt[:, :, 2] = a[:, :, 1] + 48
    gt[:, :, 3] = 9
    gt[:, S - 1, 3] = 10
    ac = (a == 1).sum(dim=(1, 2)).to(torch.int64)
    del a
    tab = _lit("\t", L, device)
    if kind == "kg":
        pos = 60000 + 35 * idx + (u[:, 4] * 30).long()
        af6 = (ac.double() / (2 * S) * 1e6).round().long().clamp(max=999999)
        dp = 8000 + (u[:, 5] * 22000).long()
        pop = [(ac.double() / (2 * S) * (0.5 + u[:, 6].double() * (k + 1) / 5) * 1e4).round().long().clamp(max=9999)
               for k in range(5)]
        info = [_lit("AC=", L, device), _digits(ac, 5, True), _lit(";AF=0.", L, device), _digits(af6, 6, False),
                _lit(f";AN={2 * S};NS={S};DP=", L, device), _digits(dp, 5, True)]
        for name, p in zip(("EAS", "AMR", "AFR", "EUR", "SAS"), pop):
            info += [_lit(f";{name}_AF=0.", L, device), _digits(p, 4, False)]
        info += [_lit(";AA=", L, device), bases[ref_i][:, None], _lit("|||;VT=SNP", L, device)]
        ex = _lit(";EX_TARGET", L, device)
        info.append(torch.where((u[:, 7] < 0.3)[:, None], ex, torch.full_like(ex, _PAD)))
        ma = _lit(";MULTI_ALLELIC", L, device)
        info.append(torch.where(multi[:, None], ma, torch.full_like(ma, _PAD)))
        idcol = [_lit("rs", L, device), _digits(100000 + idx, 9, True)]
        chromcol = _lit(chrom, L, device)
    else:
        pos = 10000 + 2 * idx
        info = [_lit("INFO", L, device)]
        idcol = [_lit("var", L, device), _digits(idx, 9, True)]
        chromcol = _lit("1", L, device)
    alt2 = torch.stack([torch.full((L,), ord(","), dtype=torch.uint8, device=device), bases[alt2_i]], dim=1)
    alt2 = torch.where(multi[:, None], alt2, torch.full_like(alt2, _PAD))
    cols = [chromcol, tab, _digits(pos, 9, True), tab, *idcol, tab, bases[ref_i][:, None], tab, bases[alt_i][:, None], alt2,
            _lit("\t100\tPASS\t", L, device), *info, _lit("\tGT\t", L, device), gt.reshape(L, S * 4)]
    rows = torch.cat(cols, dim=1)
    keep = rows != _PAD
    lens = keep.sum(dim=1).to(torch.int64)
    return rows[keep], lens


def generate(kind: str, n_lines: int, n_samples: int, seed: int = 20, first_line: int = 0, device="cuda",
             out: torch.Tensor | None = None, lines_per_chunk: int | None = None):
    """Generates n_lines data lines.  Returns (bytes tensor [total] (a view of `out` if given), lens int64 [n_lines])."""
    device = torch.device(device)
    if lines_per_chunk is None:
        lines_per_chunk = max(1, min(n_lines, (64 << 20) // max(1, 4 * n_samples)))
    pieces, lens, total = [], [], 0
    for lo in range(0, n_lines, lines_per_chunk):
        n = min(lines_per_chunk, n_lines - lo)
        flat, ln = _chunk(kind, seed, first_line + lo, n, n_samples, device)
        if out is not None:
            if total + flat.numel() > out.numel():
                raise ValueError("output tensor too small")
            out[total:total + flat.numel()] = flat
        else:
            pieces.append(flat)
        total += flat.numel()
        lens.append(ln)
    lens = torch.cat(lens) if lens else torch.zeros(0, dtype=torch.int64, device=device)
    if out is not None:
        return out[:total], lens
    return (torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.uint8, device=device)), lens


def max_line_bytes(kind: str, n_samples: int) -> int:
    return 4 * n_samples + (320 if kind == "kg" else 64)


def header(n_samples: int) -> bytes:
    import vcfgen
    return vcfgen.header(n_samples)
