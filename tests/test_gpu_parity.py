"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bit-exact is the bar (byte/integer work).  Sources of truth, in order:
  tests/golden/   fixtures written by the UNMODIFIED reference binary (oracle/make_golden.py)
  oraclelib       the C restatement, itself pinned to those fixtures by tests/test_oracle.py
Nothing here reads /root/reference.
"""
import hashlib
import importlib
import os
import subprocess

import numpy as np
import pytest

import goldenlib
import oraclelib as O
import vcfgen

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("vcf-compression_b200")


@pytest.fixture(scope="module")
def codec():
    c = pkg.Codec(0)          # raises if the CUDA library / device is missing: no silent fallback
    yield c
    c.close()


@pytest.fixture(params=["auto", "walk", "generic"])
def any_path(request, codec):
    """Run a test through the default dispatch, with the span-walking decode kernel, and through the generic kernels."""
    codec.force_generic({"auto": 0, "generic": 1, "walk": 2}[request.param])
    yield codec
    codec.force_generic(False)


def check_block(codec, data: bytes, sample_count=None, expect_path=None):
    """encode == oracle (bytes, line count, offsets); decode(encode) == oracle decode."""
    orc, oout, onl, oel, ooffs = O.compress_block(data, want_offsets=True)
    rc, out, nl, el, offs = codec.compress_block(data, want_offsets=True)
    assert rc == (-orc if orc < 0 else 0), (rc, orc)
    assert out == oout
    assert nl == onl
    if orc != 0:
        assert el == oel
        return out
    assert offs == ooffs
    if expect_path is not None:
        assert codec.last_path == expect_path
    if sample_count is not None and out:
        drc, txt, dnl, _ = codec.decompress_block(out, sample_count)
        orc2, otxt, onl2, _ = O.decompress_block(oout, sample_count)
        assert drc == (-orc2 if orc2 < 0 else 0)
        assert txt == otxt and dnl == onl2
    return out


# ---- known-answer vectors, SURVEY.md 8(c) ------------------------------------------------------
def test_kat_lines(any_path):
    req = b"1\t100\trs1\tA\tT\t100\tPASS\tAC=1\tGT\t"
    line = req + b"0|0\t0|0\t0|0\t0|1\t0|1\t1|0\t1|1\t0|0\n"
    rc, out, nl, _ = any_path.compress_block(line)
    assert rc == 0 and nl == 1
    assert out == bytes.fromhex("c0000029c000001f") + req + bytes.fromhex("03a2c181010a")
    req2 = b"1\t102\trs2\tA\tT,G\t100\tPASS\tAC=1\tGT\t"
    line2 = req2 + b"0|2\t0|0\t./.\t0/0\t0/1\t2|2\t1|1\t2|1\n"
    tail = bytes.fromhex("e1307c320901e12e2f2e09e1302f3009e1302f3109e1327c320981e1327c310a")
    rc, out, _, _ = any_path.compress_block(line2)
    assert rc == 0 and out == bytes.fromhex("c0000045c0000021") + req2 + tail
    rc, txt, nl, _ = any_path.decompress_block(out, 8)
    assert rc == 0 and txt == line2 and nl == 1


@pytest.mark.parametrize("gt,n,expect", [
    (b"0|0", 300, "7f7f2e"), (b"0|1", 70, "bfbfa8"), (b"0|0", 127, "7f"), (b"1|1", 31, "9f"),
    (b"1|0", 32, "dfc1"), (b"0|0", 128, "7f01"), (b"0|1", 62, "bfbf"), (b"1|1", 63, "9f9f81"),
])
def test_run_chunking(any_path, gt, n, expect):
    req = b"1\t1\t.\tA\tT\t.\t.\t.\tGT\t"
    line = req + b"\t".join([gt] * n) + b"\n"
    rc, out, _, _ = any_path.compress_block(line)
    assert rc == 0 and out[8 + len(req):-1].hex() == expect
    rc, txt, _, _ = any_path.decompress_block(out, n)
    assert rc == 0 and txt == line


# ---- fixtures from the reference binary ---------------------------------------------------------
def test_golden_fixtures(any_path, golden):
    for name, g in golden.items():
        rc, out = any_path.compress_vcf(g["vcf"])
        if g["entry"]["compress_rc"] != 0:
            assert rc != 0, name
            continue
        assert rc == 0, name
        assert out == g["vcfc"], name
        assert hashlib.sha256(out).hexdigest() == g["entry"]["vcfc_sha256"]
        rc, txt = any_path.decompress_vcfc(g["vcfc"])
        if g["entry"].get("decompress_rc", 0) != 0:
            assert rc != 0, name
            continue
        assert rc == 0, name
        assert txt == g["rt"], name


# ---- seeded inputs against the oracle -------------------------------------------------------------
@pytest.mark.parametrize("maker,args", [
    (vcfgen.random_vcf_like, dict(n_lines=200, n_samples=333, seed=11)),
    (vcfgen.random_vcf_like, dict(n_lines=64, n_samples=2504, seed=12)),
    (vcfgen.random_vcf_like, dict(n_lines=40, n_samples=2504, seed=13, probs=(0.5, 0.4, 0.1))),
    (vcfgen.random_vcf_like, dict(n_lines=3000, n_samples=7, seed=14)),
    (vcfgen.random_vcf_like, dict(n_lines=500, n_samples=1, seed=15)),
    (vcfgen.random_vcf_like, dict(n_lines=3, n_samples=100000, seed=16, probs=(0.98, 0.02, 0.0))),
    (vcfgen.kg_like, dict(n_lines=300, n_samples=2504, seed=17)),
    (vcfgen.kg_like, dict(n_lines=100, n_samples=1000, seed=18)),
])
def test_seeded_vs_oracle(any_path, maker, args):
    _, data = maker(**args)
    check_block(any_path, data, sample_count=args["n_samples"])


def test_regular_blocks_take_the_tile_kernels(codec):
    """GT-only data must be served by the single-pass tile kernels, not by the generic fallback."""
    for maker, args in ((vcfgen.kg_like, dict(n_lines=120, n_samples=2504, seed=41)),
                        (vcfgen.random_vcf_like, dict(n_lines=50, n_samples=2504, seed=42)),
                        (vcfgen.random_vcf_like, dict(n_lines=2, n_samples=60000, seed=43, probs=(0.999, 0.001, 0.0)))):
        _, data = maker(**args)
        check_block(codec, data, sample_count=args["n_samples"], expect_path=pkg.PATH_FAST)


def test_long_required_sections(codec):
    """INFO columns of 0.5 - 30 KB: the 9th tab lies in a later 512-byte round of the line scan, tile boundaries fall inside
    required sections, and a required section longer than the staging area goes to the log as its own segment; beyond
    ~32 KB the block goes to the generic kernels.  All byte-identical to the oracle."""
    rng = __import__("random").Random(17)
    for info_len, n_samples, expect in ((480, 700, pkg.PATH_FAST), (700, 700, pkg.PATH_FAST), (850, 300, pkg.PATH_FAST),
                                        (1300, 300, pkg.PATH_FAST), (3400, 300, pkg.PATH_FAST), (9000, 40, pkg.PATH_FAST),
                                        (30000, 100, pkg.PATH_FAST), (40000, 100, pkg.PATH_GENERIC)):
        lines = []
        for i in range(160 if info_len < 5000 else 24):
            info = "AC=1;X=" + "".join(rng.choice("ACGT0123456789;=") for _ in range(info_len + rng.randrange(40)))
            gts = "\t".join(rng.choice(("0|0",) * 12 + ("0|1", "1|0", "1|1", "0|2", "./.")) for _ in range(n_samples))
            lines.append(f"7\t{1000 + i}\trs{i}\tA\tC,G\t50\tPASS\t{info}\tGT\t{gts}\n".encode())
        data = b"".join(lines)
        check_block(codec, data, sample_count=n_samples)
        codec.compress_block(data)
        assert codec.last_path == expect, (info_len, codec.last_reject_reason)


def test_short_lines_many_per_tile(codec):
    """Short lines (hundreds of line starts per tile) stay on the tile kernels: a log segment is closed after 62 line
    starts, so a tile's line count is not bounded."""
    _, data = vcfgen.random_vcf_like(600, 140, seed=11)          # ~590-byte lines
    check_block(codec, data, sample_count=140, expect_path=pkg.PATH_FAST)
    _, data = vcfgen.random_vcf_like(600, 30, seed=12)           # ~150-byte lines: more than 62 per tile
    check_block(codec, data, sample_count=30, expect_path=pkg.PATH_FAST)
    _, data = vcfgen.random_vcf_like(20000, 30, seed=13)         # several tiles of ~200 lines each
    check_block(codec, data, sample_count=30, expect_path=pkg.PATH_FAST)
    _, data = vcfgen.random_vcf_like(30000, 1, seed=14)          # one sample per line: ~35-byte lines
    check_block(codec, data, sample_count=1)                     # (lines under 64 bytes exceed the line table: generic kernels)
    _, data = vcfgen.random_vcf_like(9000, 12, seed=15, probs=(0.2, 0.3, 0.5))
    check_block(codec, data, sample_count=12, expect_path=pkg.PATH_FAST)


@pytest.mark.parametrize("n_samples", [1, 2, 3, 5, 31, 32, 33, 127, 128, 129, 255, 1000, 3583, 3584, 3585, 4096, 7000, 8180, 8191, 8192, 8193, 20000])
def test_tile_boundaries_sweep(codec, n_samples):
    """Line widths around the 32 KB tile / 64-byte block / 127- and 31-sample chunk sizes; allele mixes from almost
    all-default to literal-heavy (allele 2: "0|2" ... stay on the 4-byte grid, so the decoder's fill-and-patch kernel
    needs several staging batches per tile)."""
    for seed, probs in ((1, (0.90, 0.08, 0.02)), (2, (0.9995, 0.0005, 0.0)), (3, (0.0, 1.0, 0.0)), (4, (0.3, 0.3, 0.4))):
        n_lines = max(3, min(400, 300000 // (4 * n_samples + 40)))
        _, data = vcfgen.random_vcf_like(n_lines, n_samples, seed=seed, probs=probs)
        check_block(codec, data, sample_count=n_samples)


def test_long_uniform_runs_across_tiles(codec):
    """Runs of one genotype that span many tiles: the 127/31 chunking must stay anchored at the run head."""
    req = b"7\t123\t.\tA\tC\t.\t.\tDP=1\tGT\t"
    for gt in (b"0|0", b"0|1", b"1|0", b"1|1", b"0/0"):
        for n in (3584 * 3 + 17, 50000):
            lines = [req + b"\t".join([gt] * n) + b"\n",
                     req + b"\t".join([b"1|1"] * 5 + [gt] * n + [b"0|1"]) + b"\n",
                     req + b"\t".join([gt] * (n // 2) + [b"2|2"] + [gt] * (n // 2)) + b"\n"]
            check_block(codec, b"".join(lines), sample_count=None)


def test_edge_case_lines(any_path):
    for s in (1, 3, 8, 40, 129):
        lines = vcfgen.edge_case_lines(s)
        check_block(any_path, b"".join(lines), sample_count=s)
        for ln in lines:
            check_block(any_path, ln, sample_count=s)


def test_run_length_lines(any_path):
    check_block(any_path, b"".join(vcfgen.run_length_lines()), sample_count=700)
    check_block(any_path, b"".join(vcfgen.run_length_lines(lengths=(126, 127, 128, 381), n_samples=2504)), sample_count=2504)


def test_fuzz_blocks(any_path):
    """Seeded random blocks over regular and odd genotype columns (the alphabet of tests/test_oracle.py's fuzz against the
    reference binary), run lengths around the 31 / 127 chunk sizes, long and short lines mixed in one block."""
    import random
    rng = random.Random(77)
    alphabet = ["0|0"] * 6 + ["0|1", "1|0", "1|1", "0|2", "2|1", "./.", ".", "0/0", "0/1", "1", "10|0", "0|0|0", "0|0:3", "1|1:12"]
    regular = ["0|0"] * 8 + ["0|1", "1|0", "1|1", "0|2", "./.", "0/1"]
    runs = [1, 2, 3, 30, 31, 32, 33, 62, 63, 126, 127, 128, 129, 254, 255, 300, 1000]
    for case in range(24):
        n_samples = rng.choice([7, 200, 700, 2504, 9000])
        alpha = regular if case % 2 else alphabet            # odd cases stay inside the tile kernels' grammar
        lines = []
        for i in range(rng.randrange(1, 40)):
            gts = []
            while len(gts) < n_samples:
                gts += [rng.choice(alpha)] * rng.choice(runs if rng.random() < 0.5 else [1, 1, 1, 2, 5])
            gts = gts[:n_samples]
            info = "AC=%d;" % rng.randrange(100) + "X" * rng.randrange(0, 300)
            lines.append("%s\t%d\trs%d\tA\tC,G\t%d\tPASS\t%s\tGT\t%s\n" % (
                rng.choice(["1", "20", "X"]), 100 + 7 * i, i, rng.randrange(1000), info, "\t".join(gts)))
        check_block(any_path, "".join(lines).encode(), sample_count=n_samples)


def test_last_line_without_newline(codec):
    """A last line that ends at EOF (compress.cpp:218 getline) stays on the tile kernels when its last sample is whole;
    a trailing tab or a cut sample goes to the generic kernels.  Short blocks and blocks of several tiles."""
    for n_lines, n_samples, seed in ((5, 9, 3), (40, 2504, 4), (3, 30000, 5)):
        _, data = vcfgen.random_vcf_like(n_lines, n_samples, seed=seed)
        body = data[:-1]
        check_block(codec, body, sample_count=n_samples, expect_path=pkg.PATH_FAST)          # ... 0|1<EOF>
        lit = body[:-3] + b"./."
        check_block(codec, lit, sample_count=n_samples, expect_path=pkg.PATH_FAST)           # ... ./.<EOF>
        check_block(codec, body + b"\t")                                                     # trailing tab, no newline
        check_block(codec, body[:-1])                                                        # cut sample "0|"
        check_block(codec, body[:-2])
        check_block(codec, body[:-3])                                                        # ends with a tab


# ---- samples that are not 3 bytes wide: walked term by term inside the tile kernel (round 2) --------------------------
def _mutate_samples(data: bytes, rng, rate: float, pool) -> bytes:
    """Replaces a fraction of the sample columns of every data line by odd-width terms."""
    out = []
    for line in data.split(b"\n")[:-1]:
        cols = line.split(b"\t")
        for i in range(9, len(cols)):
            if rng.random() < rate:
                cols[i] = rng.choice(pool)
        out.append(b"\t".join(cols))
    return b"\n".join(out) + b"\n"


ODD_TERMS = [b"10|0", b"0|10", b".", b"0", b"1", b"0|0|0", b"0|1:35:99", b"./.:.", b"1|1\r"[:3], b"12", b"0|0:3"]


@pytest.mark.parametrize("n_lines,n_samples,rate,seed", [
    (60, 2504, 0.0005, 1),      # a few odd terms per block of several tiles
    (60, 2504, 0.02, 2),        # every line has dozens
    (400, 333, 0.01, 3),
    (8, 40000, 0.0002, 4),      # long lines: an odd term many tiles away from the line start
    (3000, 7, 0.05, 5),
    (12, 2504, 1.0, 6),         # nothing but odd terms (a small block: within the term walker's budget)
])
def test_odd_width_samples_stay_on_the_tile_kernel(codec, n_lines, n_samples, rate, seed):
    """10|0, haploid calls, GT:DP columns ...: the tile kernel walks the rest of such a line term by term (lane 0) and goes on
    on the 4-byte grid with the next line; bytes equal the oracle's, the block is not rerun on the generic kernels."""
    rng = __import__("random").Random(seed)
    _, data = vcfgen.random_vcf_like(n_lines, n_samples, seed=seed)
    data = _mutate_samples(data, rng, rate, ODD_TERMS)
    check_block(codec, data, sample_count=n_samples, expect_path=pkg.PATH_FAST)
    # 1 % irregular LINES in an otherwise regular block
    _, reg = vcfgen.kg_like(n_lines, n_samples, seed=seed + 50)
    lines = reg.split(b"\n")[:-1]
    for i in range(0, len(lines), 100):
        lines[i] = _mutate_samples(lines[i] + b"\n", rng, 0.3, ODD_TERMS)[:-1]
    check_block(codec, b"\n".join(lines) + b"\n", sample_count=n_samples, expect_path=pkg.PATH_FAST)


def test_odd_width_runs_across_odd_terms(codec):
    """Runs before / after an odd term, at the 31 / 127 chunk limits, and odd terms at the line's first / last sample."""
    req = b"7\t123\t.\tA\tC\t.\t.\tDP=1\tGT\t"
    lines = []
    for gt, m in ((b"0|0", 127), (b"0|1", 31), (b"1|1", 31)):
        for n in (1, m - 1, m, m + 1, 2 * m, 2 * m + 1, 700):
            for odd in (b"10|0", b"0", b"0|0:7"):
                lines.append(req + b"\t".join([gt] * n + [odd] + [gt] * n) + b"\n")
                lines.append(req + b"\t".join([odd] + [gt] * n) + b"\n")
                lines.append(req + b"\t".join([gt] * n + [odd]) + b"\n")
                lines.append(req + b"\t".join([gt] * n + [odd, odd] + [gt] * 5 + [odd]) + b"\n")
    data = b"".join(lines)
    check_block(codec, data, expect_path=pkg.PATH_FAST)
    # a run that is open when the tile ends inside an odd line, and an odd term right at tile boundaries
    for n in (8180, 8190, 8191, 8192, 8193, 16383, 16384, 16385, 30000):
        for gt in (b"0|0", b"1|0"):
            line = req + b"\t".join([gt] * n + [b"10|0"] + [gt] * n + [b"0"] + [gt] * 3000) + b"\n"
            check_block(codec, line * 3, expect_path=pkg.PATH_FAST)


def test_gt_dp_gq_block(codec):
    """An all-GT:DP:GQ block (every sample a literal of varying width), and haploid / diploid mixes: after eight odd-width
    terms in a row all 32 lanes walk the rest of the line portion (parallel_portion)."""
    rng = __import__("random").Random(9)
    lines = []
    for i in range(300):
        gts = "\t".join("%s:%d:%d" % (rng.choice(("0|0", "0|1", "1|1", "./.")), rng.randrange(60), rng.randrange(100)) for _ in range(500))
        lines.append(f"2\t{500 + i}\t.\tG\tA\t.\tPASS\tDP=9\tGT:DP:GQ\t{gts}\n".encode())
    check_block(codec, b"".join(lines[:40]), sample_count=500, expect_path=pkg.PATH_FAST)
    check_block(codec, b"".join(lines) * 3, sample_count=500, expect_path=pkg.PATH_FAST)
    # haploid calls mixed with diploid ones (chrX): runs of coded terms between odd-width ones, across the lanes' ranges
    for n_samples, p_hap in ((2504, 0.5), (2504, 0.95), (700, 0.2), (40000, 0.5)):
        mix = []
        for i in range(24 if n_samples < 10000 else 3):
            gts = []
            while len(gts) < n_samples:
                g = rng.choice((b"0", b"1", b".")) if rng.random() < p_hap else rng.choice((b"0|0", b"0|0", b"0|0", b"0|1", b"1|1"))
                gts += [g] * rng.choice((1, 1, 2, 5, 40, 200))
            mix.append(b"X\t%d\t.\tG\tA\t.\tPASS\tDP=9\tGT\t" % (900 + i) + b"\t".join(gts[:n_samples]) + b"\n")
        check_block(codec, b"".join(mix), sample_count=n_samples, expect_path=pkg.PATH_FAST)
    # no newline at the end of an odd-width line, and an odd-width line that ends the input right behind a tile boundary
    check_block(codec, (b"".join(lines[:5]))[:-1], sample_count=500, expect_path=pkg.PATH_FAST)


def test_odd_stretch_transitions(codec):
    """parallel_portion's ways out and in: a stretch of odd-width terms followed by thousands of 3-byte terms (the grid takes over
    with the run that is open, at every chunk phase), a literal longer than the 2 KB window in the middle of odd terms (the term
    walker takes it), windows that end in 3-byte runs at the 31 / 127 limits, the next line going to the walkers at once, and an
    empty term inside an odd stretch (rejected like everywhere else)."""
    rng = __import__("random").Random(77)
    req = b"X\t55\t.\tA\tC\t.\t.\tDP=1\tGT:DP\t"
    odd = lambda n: [b"%s:%d" % (rng.choice((b"0|0", b"0|1", b"1|1", b"./.")), rng.randrange(200)) for _ in range(n)]
    lines = []
    for gt, m in ((b"0|0", 127), (b"0|1", 31), (b"1|0", 31)):
        for n_odd in (9, 40, 300, 777):
            for n_reg in (1, 30, 31, 32, 126, 127, 128, 400, 513, 5000):
                lines.append(req + b"\t".join(odd(n_odd) + [gt] * n_reg) + b"\n")
                lines.append(req + b"\t".join(odd(n_odd) + [gt] * n_reg + odd(3) + [gt] * 40) + b"\n")
                lines.append(req + b"\t".join([gt] * n_reg + odd(n_odd) + [gt] * (n_reg + 1) + odd(n_odd)) + b"\n")
    check_block(codec, b"".join(lines), expect_path=pkg.PATH_FAST)
    rng.shuffle(lines)
    check_block(codec, b"".join(lines), expect_path=pkg.PATH_FAST)
    # literals of 2 ... 9 KB among odd terms; right behind the line start; as the line's last term
    for big in (2040, 2048, 2049, 4100, 9000):
        blob = b"0|1:" + b"7" * big
        ls = [req + b"\t".join(odd(50) + [blob] + odd(50) + [b"0|0"] * 200) + b"\n",
              req + b"\t".join([blob] + odd(20)) + b"\n",
              req + b"\t".join(odd(20) + [blob]) + b"\n",
              req + b"\t".join(odd(300) + [blob, blob] + [b"1|1"] * 70 + odd(9)) + b"\n"]
        check_block(codec, b"".join(ls) * 2, expect_path=pkg.PATH_FAST)
        check_block(codec, (b"".join(ls) * 2)[:-1], expect_path=pkg.PATH_FAST)
    # haploid calls only (2-byte terms: 32 per block), and 1-byte / 3-byte mixes that end exactly at window ends
    for n in (1000, 1023, 1024, 1025, 4096, 20000):
        hap = [rng.choice((b"0", b"1", b".")) for _ in range(n)]
        mix = [rng.choice((b"0", b"0|0", b"0|0", b"1|1", b"10")) for _ in range(n)]
        check_block(codec, req + b"\t".join(hap) + b"\n" + req + b"\t".join(mix) + b"\n" + req + b"\t".join(hap), expect_path=pkg.PATH_FAST)
    # an empty term inside an odd stretch: the reference drops it (utils.cpp:82-116), the tile path hands the block on
    bad = req + b"\t".join(odd(30) + [b""] + odd(30)) + b"\n"
    check_block(codec, lines[0] + bad + lines[1])


@pytest.mark.parametrize("seed", range(12))
def test_odd_width_fuzz(codec, seed):
    """Random blocks of random term mixes against the oracle, both directions (vcfgen.odd_mix_lines: every line draws its own mix
    of coded genotypes, 3-byte literals, odd-width terms of 1 ... 14 bytes and, rarely, literals of kilobytes, in stretches of
    random length), so that the encoder's grid / all-lane windows / literal-only windows / term walker and the decoder's two tile
    kernels meet in every order; sample counts from 1 to a few thousand, lines that end at EOF.  The oracle is pinned on the same
    generator against the unmodified reference binary (tests/test_oracle.py::test_odd_mix_vs_reference_binary)."""
    rng = __import__("random").Random(1000 + seed)
    for _ in range(5):
        n_samples = rng.choice((1, 2, 7, 31, 64, 500, 1000, 2504, 6000))
        data = b"".join(vcfgen.odd_mix_lines(rng, n_samples, rng.choice((1, 3, 20, 60))))
        if rng.random() < 0.3:
            data = data[:-1]                                     # the last line ends with the input
        check_block(codec, data, sample_count=n_samples)


def test_ragged_and_empty_inputs(any_path):
    rc, out, nl, _ = any_path.compress_block(b"")
    assert rc == 0 and out == b"" and nl == 0
    _, data = vcfgen.random_vcf_like(5, 9, seed=3)
    lines = data.split(b"\n")[:-1]
    check_block(any_path, b"\n\n" + b"\n\n\n".join(lines) + b"\n\n", sample_count=9)   # blank lines are dropped
    check_block(any_path, data[:-1], sample_count=9)                                    # no final newline
    check_block(any_path, data.replace(b"\n", b"\r\n"), sample_count=9)                 # CR-LF
    check_block(any_path, b"1\t5\t.\tA\tC\t.\t.\t.\tGT\n", sample_count=0)              # 9 columns, no samples
    check_block(any_path, lines[0] + b"\t\n" + lines[1] + b"\n")                        # trailing tab
    check_block(any_path, lines[0].replace(b"\t0|0\t", b"\t\t", 1) + b"\n")             # empty field
    wide = b"1\t7\t.\tA\tC\t.\t.\t" + b"X" * 70000 + b"\tGT\t" + b"\t".join([b"0|1"] * 50) + b"\n"
    check_block(any_path, wide + lines[2] + b"\n" + wide, sample_count=None)            # a 70 kB INFO column


def test_error_lines(any_path):
    ok = b"1\t5\t.\tA\tC\t.\t.\t.\tGT\t0|0\n"
    for bad, code in ((b"1\t5\t.\tA\tC\t.\t.\t.\n", pkg.E_EIGHTCOLS), (b"1\t5\t.\tA\tC\t.\t.\n", pkg.E_TOOFEW),
                      (b"\t\t\n", pkg.E_TOOFEW), (b"bad\n", pkg.E_TOOFEW)):
        rc, out, nl, el = any_path.compress_block(bad)
        assert rc == code and nl == 0 and el == 0 and out == b""
        rc, out, nl, el = any_path.compress_block(ok * 3 + bad + ok)
        orc, oout, onl, oel = O.compress_block(ok * 3 + bad + ok)
        assert rc == code == -orc and nl == onl == 3 and el == oel == 3 and out == oout
    rc, out, nl, _ = any_path.compress_block(ok, out_cap=10)
    assert rc == pkg.E_CAP


def test_exact_capacity(any_path):
    """A multi-tile block into a buffer of exactly the output size succeeds; one byte less is VCFC_E_CAP (both ways)."""
    _, data = vcfgen.kg_like(60, 2504, seed=3)
    ref = O.compress_block(data)[1]
    rc, out, nl, _ = any_path.compress_block(data, out_cap=len(ref))
    assert rc == 0 and out == ref and nl == 60
    assert any_path.compress_block(data, out_cap=len(ref) - 1)[0] == pkg.E_CAP
    rc, txt, nl, _ = any_path.decompress_block(ref, 2504, out_cap=len(data))
    assert rc == 0 and txt == data and nl == 60
    assert any_path.decompress_block(ref, 2504, out_cap=len(data) - 1)[0] == pkg.E_CAP


def test_decode_rejects_bad_input(any_path):
    _, data = vcfgen.random_vcf_like(3, 8, seed=1)
    rc, out, _, _ = any_path.compress_block(data)
    assert any_path.decompress_block(out, 8)[0] == 0
    assert any_path.decompress_block(out[:-1], 8)[0] == pkg.E_TRUNC
    bad = bytearray(out)
    bad[0] = 0x40
    assert any_path.decompress_block(bytes(bad), 8)[0] == pkg.E_FORMAT
    assert any_path.decompress_block(out, 9)[0] != 0
    for sc in (7, 5, 1):       # too small a sample count: whatever the reference's loop does, token by token
        orc, otxt, _, _ = O.decompress_block(out, sc)
        rc, txt, _, _ = any_path.decompress_block(out, sc)
        assert (rc == 0) == (orc == 0) and txt == otxt
    rc, txt, nl, _ = any_path.decompress_block(out + b"\x01\x02\x03", 8)     # <8 trailing bytes = EOF
    assert rc == 0 and txt == data and nl == 3
    assert any_path.decompress_block(out, 8, out_cap=20)[0] == pkg.E_CAP


def test_chunked_host_path_matches_single_call(codec, monkeypatch):
    """The host-pointer API splits at newlines into chunks; bytes must not depend on the chunk size."""
    _, data = vcfgen.random_vcf_like(400, 2504, seed=21)       # ~4 MB
    ref = O.compress_block(data, want_offsets=True)
    monkeypatch.setenv("VCFC_CHUNK_MB", "1")
    monkeypatch.setenv("VCFC_DCHUNK_MB", "1")
    rc, out, nl, _, offs = codec.compress_block(data, want_offsets=True)
    assert rc == 0 and out == ref[1] and nl == ref[2] and offs == ref[4]
    rc, txt, nl2, _ = codec.decompress_block(out, 2504)
    assert rc == 0 and txt == data and nl2 == nl
    cut = data.rfind(b"\n", 0, 2_000_000) + 1
    bad = data[:cut] + b"oops\n" + data[cut:]
    orc, oout, onl, oel = O.compress_block(bad)
    rc, out, nl, el = codec.compress_block(bad)
    assert rc == -orc and out == oout and nl == onl and el == oel


def test_device_pointer_api(codec):
    import torch
    _, data = vcfgen.kg_like(200, 2504, seed=5)
    ref = O.compress_block(data)[1]
    dev = torch.device("cuda:0")
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_out = torch.empty(len(data) * 2 + 4096, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    d_offs = torch.zeros(256, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    codec.encode_dev(d_in.data_ptr(), len(data), d_out.data_ptr(), d_out.numel(), d_res.data_ptr(), st,
                     d_offs.data_ptr(), 256)
    r = codec.fetch_result(d_res.data_ptr(), st)
    assert r.status == 0 and r.n_lines == 200 and r.out_len == len(ref)
    assert bytes(d_out[:r.out_len].cpu().numpy()) == ref
    offs = d_offs[:200].cpu().numpy()
    assert offs[0] == 0 and all(ref[o] >> 6 == 3 for o in offs)
    sz = codec.decode_size_dev(d_out.data_ptr(), r.out_len, 2504, st)
    assert sz.status == 0 and sz.out_len == len(data) and sz.n_lines == 200
    d_txt = torch.empty(len(data) + 64, dtype=torch.uint8, device=dev)
    codec.decode_dev(d_out.data_ptr(), r.out_len, 2504, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), st)
    r2 = codec.fetch_result(d_res.data_ptr(), st)
    assert r2.status == 0 and r2.out_len == len(data)
    assert torch.equal(d_txt[:len(data)], d_in)
    # device buffers at odd addresses: compressed input and text output one / three bytes off any alignment
    d_c2 = torch.empty(r.out_len + 64, dtype=torch.uint8, device=dev)
    d_c2[1:1 + r.out_len] = d_out[:r.out_len]
    d_t2 = torch.zeros(len(data) + 64, dtype=torch.uint8, device=dev)
    codec.decode_dev(d_c2.data_ptr() + 1, r.out_len, 2504, d_t2.data_ptr() + 3, len(data) + 32, d_res.data_ptr(), st)
    r3 = codec.fetch_result(d_res.data_ptr(), st)
    assert r3.status == 0 and r3.out_len == len(data)
    assert torch.equal(d_t2[3:3 + len(data)], d_in)


# ---- file drivers and the CLI (reference verbs) ----------------------------------------------------
def test_file_drivers_and_cli(codec, golden, tmp_path):
    g = golden["refgen_300x40"]
    ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
    open(ip, "wb").write(g["vcf"])
    assert codec.compress(ip, op) == 0
    assert open(op, "rb").read() == g["vcfc"]
    assert codec.decompress2_fd(op, rp) == 0
    assert open(rp, "rb").read() == g["rt"]
    if os.path.exists(pkg.CLI_PATH):
        op2, rp2 = str(tmp_path / "b.vcfc"), str(tmp_path / "b.rt")
        assert subprocess.run([pkg.CLI_PATH, "compress", ip, op2]).returncode == 0
        assert open(op2, "rb").read() == g["vcfc"]
        assert subprocess.run([pkg.CLI_PATH, "decompress", op2, rp2]).returncode == 0
        assert open(rp2, "rb").read() == g["rt"]
        bad = str(tmp_path / "bad.vcf")
        open(bad, "wb").write(golden["abort_8cols"]["vcf"])
        assert subprocess.run([pkg.CLI_PATH, "compress", bad, op2], capture_output=True).returncode != 0


def test_query_matches_reference_outputs(codec, golden, tmp_path):
    qs = goldenlib.manifest()["_queries"]
    for name, cases in qs.items():
        fp = str(tmp_path / (name + ".vcfc"))
        open(fp, "wb").write(golden[name]["vcfc"])
        for c in cases:
            outp = str(tmp_path / "q.out")
            fd = os.open(outp, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
            rc = codec.query(fp, c["q"], fd)
            os.close(fd)
            got = open(outp, "rb").read()
            assert rc == 0 and len(got) == c["len"], (name, c["q"])
            assert hashlib.sha256(got).hexdigest() == c["sha256"], (name, c["q"])
    assert codec.query(fp, "1:5", 1) == pkg.E_QUERY


# ---- binned index (.vcfci): the next row of the scope table -------------------------------
def test_binned_index_matches_oracle_and_reference(codec, tmp_path):
    """vcfc_create_binned_index_file (per-line END / chromosome index on the GPU) writes the bytes of the reference's
    create-binned-index: against the committed fixtures of the reference binary, the oracle on fresh files, the CLI verb."""
    import json
    idir = os.path.join(goldenlib.GOLDEN, "index")
    man = json.load(open(os.path.join(idir, "MANIFEST.json")))
    for fn, e in man.items():
        name = fn.split(".bin")[0]
        vcfc = open(os.path.join(idir, name + ".vcfc"), "rb").read() if name.startswith("sv_") else goldenlib.read(name + ".vcfc")
        ip, xp = str(tmp_path / "g.vcfc"), str(tmp_path / "g.vcfci")
        open(ip, "wb").write(vcfc)
        rc, n = codec.create_binned_index(ip, xp, e["entries_per_bin"])
        assert rc == 0 and n == e["entries"], fn
        assert open(xp, "rb").read() == open(os.path.join(idir, fn), "rb").read(), fn
    # a larger file with long INFO columns, through the compressor of this library, several bin sizes
    h, d = vcfgen.kg_like(3000, 100, seed=31)
    vp, cp = str(tmp_path / "k.vcf"), str(tmp_path / "k.vcfc")
    open(vp, "wb").write(h + d)
    assert codec.compress(vp, cp) == 0
    vcfc = open(cp, "rb").read()
    for b in (1, 10, 1000, 100000):
        rc, n = codec.create_binned_index(cp, cp + ".x", b)
        orc, oidx = O.build_binned_index(vcfc, b)
        assert rc == 0 and n == orc and open(cp + ".x", "rb").read() == oidx, b
    # the CLI verb writes <file>.vcfci (main.cpp:4097-4115)
    assert subprocess.run([pkg.CLI_PATH, "create-binned-index", "10", cp], capture_output=True).returncode == 0
    assert open(cp + ".vcfci", "rb").read() == O.build_binned_index(vcfc, 10)[1]
    # malformed input: the reference throws, the library returns a code
    bad = bytearray(vcfc)
    data0 = vcfc.index(b"\n", vcfc.index(b"\n#CHROM") + 1) + 1
    tab = vcfc.index(b"\t", data0 + 8)
    bad[tab + 1] = ord("x")                                  # POS of the first line is no longer a number
    open(cp, "wb").write(bytes(bad))
    assert codec.create_binned_index(cp, cp + ".x", 10)[0] == pkg.E_FORMAT
    assert O.build_binned_index(bytes(bad), 10)[0] < 0
    assert codec.create_binned_index(str(tmp_path / "missing.vcfc"), cp + ".x", 10)[0] == pkg.E_IO


def test_indexed_query_matches_reference_outputs(codec, tmp_path):
    """vcfc_query_binned_index_file / `vcfc query-binned-index` against the outputs of the reference binary
    (tests/golden/index/QUERIES.json): the index is built by this library first, then queried."""
    import json
    idir = os.path.join(goldenlib.GOLDEN, "index")
    cases = json.load(open(os.path.join(idir, "QUERIES.json")))
    built = {}
    for key, e in sorted(cases.items()):
        name, b, region = key.split("|")
        cp = str(tmp_path / ("%s.%s.vcfc" % (name, b)))
        if (name, b) not in built:
            vcfc = open(os.path.join(idir, name + ".vcfc"), "rb").read() if name.startswith("sv_") else goldenlib.read(name + ".vcfc")
            open(cp, "wb").write(vcfc)
            assert codec.create_binned_index(cp, cp + ".vcfci", int(b))[0] == 0
            built[(name, b)] = True
        outp = str(tmp_path / "q.out")
        fd = os.open(outp, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
        rc = codec.query_binned_index(cp, region, fd)
        os.close(fd)
        out = open(outp, "rb").read()
        assert rc == 0 and hashlib.sha256(out).hexdigest() == e["sha256"], key
    # the CLI verb, and the error paths
    cp = str(tmp_path / "sv_sorted.4.vcfc")
    r = subprocess.run([pkg.CLI_PATH, "query-binned-index", cp, "X:5000-9000"], capture_output=True)
    assert r.returncode == 0 and hashlib.sha256(r.stdout).hexdigest() == cases["sv_sorted|4|X:5000-9000"]["sha256"]
    assert codec.query_binned_index(cp, "X:5000", 1) == pkg.E_QUERY
    assert codec.query_binned_index(str(tmp_path / "nope.vcfc"), "X:1-2", 1) == pkg.E_IO


def test_live_reference_binary_if_shipped(codec, tmp_path):
    """oracle/_ref/main_release (the unmodified reference, prebuilt) travels to the GPU box."""
    if not O.have_ref_binary():
        pytest.skip("oracle/_ref/main_release not shipped")
    h, d = vcfgen.random_vcf_like(150, 2504, seed=31)
    ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
    open(ip, "wb").write(h + d)
    if subprocess.run([O.REF_BIN, "compress", ip, op]).returncode != 0:
        pytest.skip("reference binary does not run on this box")
    rc, mine = codec.compress_vcf(h + d)
    assert rc == 0 and mine == open(op, "rb").read()
    assert subprocess.run([O.REF_BIN, "decompress", op, rp]).returncode == 0
    rc, txt = codec.decompress_vcfc(mine)
    assert rc == 0 and txt == open(rp, "rb").read() == h + d
