"""CPU suite: pins the C restatement (oracle/vcfc_oracle.c) against the reference.

(a) SURVEY.md 8(c) known-answer vectors, typed in here by hand from the survey table;
(b) tests/golden/ fixtures written by the unmodified reference binary;
(c) the reference binary run live on fresh seeded inputs, when oracle/_ref is present.
"""
import hashlib
import json
import os
import subprocess
import tempfile

import pytest

import goldenlib
import oraclelib as O
import vcfgen

HDR = (b"##fileformat=VCFv4.1\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
       b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + b"\t".join(b"S%d" % i for i in range(8)) + b"\n")


def enc1(line: bytes) -> bytes:
    rc, out, nl, _ = O.compress_block(line)
    assert rc == 0 and nl == 1
    return out


def test_kat_survey_8c_line1():
    req = b"1\t100\trs1\tA\tT\t100\tPASS\tAC=1\tGT\t"
    line = req + b"0|0\t0|0\t0|0\t0|1\t0|1\t1|0\t1|1\t0|0\n"
    assert len(req) == 31
    assert enc1(line) == bytes.fromhex("c0000029c000001f") + req + bytes.fromhex("03a2c181010a")


def test_kat_survey_8c_line2():
    req = b"1\t102\trs2\tA\tT,G\t100\tPASS\tAC=1\tGT\t"
    line = req + b"0|2\t0|0\t./.\t0/0\t0/1\t2|2\t1|1\t2|1\n"
    tail = bytes.fromhex("e1307c320901e12e2f2e09e1302f3009e1302f3109e1327c320981e1327c310a")
    assert enc1(line) == bytes.fromhex("c0000045c0000021") + req + tail


def test_kat_survey_8c_line3_gtdp():
    req = b"1\t104\trs3\tA\tT\t100\tPASS\tAC=1\tGT:DP\t"
    s = [b"0|0:3", b"0|0:3", b"0|1:9", b"0|0:3"] * 2
    body = b"".join(b"\xe1" + x + (b"\t" if i < 7 else b"") for i, x in enumerate(s))
    assert enc1(req + b"\t".join(s) + b"\n") == bytes.fromhex("c000005ec0000022") + req + body + b"\n"


def test_kat_survey_8c_line4_all_ref():
    req = b"1\t106\trs4\tA\tT\t100\tPASS\tAC=1\tGT\t"
    assert enc1(req + b"\t".join([b"0|0"] * 8) + b"\n") == bytes.fromhex("c0000025c000001f") + req + b"\x08\n"


@pytest.mark.parametrize("gt,n,expect", [
    (b"0|0", 300, "7f7f2e"), (b"0|1", 70, "bfbfa8"), (b"0|0", 127, "7f"), (b"1|1", 31, "9f"),
    (b"1|0", 32, "dfc1"), (b"0|0", 128, "7f01"), (b"0|1", 62, "bfbf"), (b"1|1", 63, "9f9f81"),
])
def test_run_chunking(gt, n, expect):
    req = b"1\t1\t.\tA\tT\t.\t.\t.\tGT\t"
    out = enc1(req + b"\t".join([gt] * n) + b"\n")
    assert out[8 + len(req):-1].hex() == expect


def test_error_codes():
    assert O.compress_block(b"1\t5\t.\tA\tC\t.\t.\t.\n")[0] == O.E_EIGHTCOLS
    assert O.compress_block(b"1\t5\t.\tA\tC\t.\t.\n")[0] == O.E_TOOFEW
    assert O.compress_block(b"\t\t\n")[0] == O.E_TOOFEW
    rc, out, nl, el = O.compress_block(b"1\t5\t.\tA\tC\t.\t.\t.\tGT\t0|0\nbad\n")
    assert rc == O.E_TOOFEW and el == 1 and nl == 1


def test_golden_fixtures(golden):
    assert len(golden) >= 12
    for name, g in golden.items():
        rc, out = O.compress_vcf(g["vcf"])
        if g["entry"]["compress_rc"] != 0:
            assert rc != 0, name            # the reference aborted on this input
            continue
        assert rc == 0, name
        assert out == g["vcfc"], name
        assert hashlib.sha256(out).hexdigest() == g["entry"]["vcfc_sha256"]
        rc, txt = O.decompress_vcfc(g["vcfc"])
        assert rc == 0, name
        assert txt == g["rt"], name
        assert hashlib.sha256(txt).hexdigest() == g["entry"]["rt_sha256"]


def test_offsets_and_counts():
    _, data = vcfgen.random_vcf_like(50, 40, seed=3)
    rc, out, nl, _, offs = O.compress_block(data, want_offsets=True)
    assert rc == 0 and nl == 50 and offs[0] == 0
    for i, o in enumerate(offs):
        ll = int.from_bytes(out[o:o + 4], "big") & 0x3FFFFFFF
        nxt = offs[i + 1] if i + 1 < nl else len(out)
        assert o + 4 + ll == nxt


def test_decode_rejects_bad_input():
    _, data = vcfgen.random_vcf_like(3, 8, seed=1)
    rc, out, _, _ = O.compress_block(data)
    assert O.decompress_block(out, 8)[0] == 0
    assert O.decompress_block(out[:-1], 8)[0] == O.E_TRUNC
    bad = bytearray(out); bad[0] = 0x40
    assert O.decompress_block(bytes(bad), 8)[0] == O.E_FORMAT
    assert O.decompress_block(out, 9)[0] < 0


@pytest.mark.skipif(not O.have_ref_binary(), reason="oracle/_ref/main_release not built")
@pytest.mark.parametrize("maker,args", [
    (vcfgen.random_vcf_like, dict(n_lines=120, n_samples=333, seed=11)),
    (vcfgen.random_vcf_like, dict(n_lines=30, n_samples=2504, seed=12, probs=(0.5, 0.4, 0.1))),
    (vcfgen.kg_like, dict(n_lines=40, n_samples=1000, seed=13)),
])
def test_live_reference_binary(maker, args):
    h, d = maker(**args)
    vcf = h + d
    with tempfile.TemporaryDirectory() as wd:
        ip, op, rp = (os.path.join(wd, x) for x in ("a.vcf", "a.vcfc", "a.rt"))
        open(ip, "wb").write(vcf)
        assert subprocess.run([O.REF_BIN, "compress", ip, op]).returncode == 0
        ref = open(op, "rb").read()
        rc, mine = O.compress_vcf(vcf)
        assert rc == 0 and mine == ref
        assert subprocess.run([O.REF_BIN, "decompress", op, rp]).returncode == 0
        rc, txt = O.decompress_vcfc(ref)
        assert rc == 0 and txt == open(rp, "rb").read() == vcf


@pytest.mark.skipif(not O.have_ref_binary(), reason="oracle/_ref/main_release not built")
def test_fuzz_vs_reference_binary(tmp_path):
    """Seeded random files over an alphabet of regular and odd genotype columns, run lengths around the 31 / 127 chunk
    sizes, CR-LF and GT:DP lines: the oracle and the unmodified reference must write the same bytes both ways."""
    import random
    rng = random.Random(2024)
    alphabet = ["0|0"] * 6 + ["0|1", "1|0", "1|1", "0|2", "2|1", "./.", ".", "0/0", "0/1", "1", "10|0", "0|0|0", "0|0:3", "1|1:12"]
    runs = [1, 2, 3, 30, 31, 32, 33, 62, 63, 126, 127, 128, 129, 254, 255, 300]
    for case in range(40):
        n_samples = rng.choice([1, 2, 7, 31, 32, 200, 700])
        crlf = rng.random() < 0.15
        lines = []
        for i in range(rng.randrange(1, 12)):
            gts = []
            while len(gts) < n_samples:
                gts += [rng.choice(alphabet)] * rng.choice(runs if rng.random() < 0.5 else [1, 1, 1, 2, 5])
            gts = gts[:n_samples]
            info = "AC=%d;AF=0.%d" % (rng.randrange(100), rng.randrange(1000))
            fmt = "GT:DP" if any(":" in g for g in gts) else "GT"
            lines.append("%s\t%d\trs%d\tA\tC,G\t%d\tPASS\t%s\t%s\t%s%s\n" % (
                rng.choice(["1", "20", "X"]), 100 + 7 * i, i, rng.randrange(1000), info, fmt, "\t".join(gts), "\r" if crlf else ""))
        vcf = vcfgen.header(n_samples) + "".join(lines).encode()
        ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
        open(ip, "wb").write(vcf)
        assert subprocess.run([O.REF_BIN, "compress", ip, op], capture_output=True).returncode == 0, case
        ref = open(op, "rb").read()
        rc, mine = O.compress_vcf(vcf)
        assert rc == 0 and mine == ref, case
        assert subprocess.run([O.REF_BIN, "decompress", op, rp], capture_output=True).returncode == 0, case
        rc, txt = O.decompress_vcfc(ref)
        assert rc == 0 and txt == open(rp, "rb").read(), case


@pytest.mark.skipif(not O.have_ref_binary(), reason="oracle/_ref/main_release not built")
def test_odd_mix_vs_reference_binary(tmp_path):
    """The generator of the GPU suite's odd-width fuzz (haploid stretches, GT:DP:GQ stretches, multi-digit alleles, literals of
    kilobytes, runs at the chunk limits) through the unmodified reference binary: the oracle writes the same bytes both ways, so the
    CUDA path is compared with something the reference itself confirms on this class of input."""
    import random
    for seed in range(8):
        rng = random.Random(500 + seed)
        n_samples = rng.choice((1, 2, 7, 31, 64, 500, 1000, 2504))
        lines = vcfgen.odd_mix_lines(rng, n_samples, rng.choice((1, 3, 20)))
        vcf = vcfgen.header(n_samples) + b"".join(lines)
        if rng.random() < 0.3:
            vcf = vcf[:-1]                                       # the last line ends with the file
        ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
        open(ip, "wb").write(vcf)
        assert subprocess.run([O.REF_BIN, "compress", ip, op], capture_output=True).returncode == 0, seed
        ref = open(op, "rb").read()
        rc, mine = O.compress_vcf(vcf)
        assert rc == 0 and mine == ref, seed
        assert subprocess.run([O.REF_BIN, "decompress", op, rp], capture_output=True).returncode == 0, seed
        rc, txt = O.decompress_vcfc(ref)
        assert rc == 0 and txt == open(rp, "rb").read(), seed


# ---- binned index (.vcfci): oracle of the next scope row, not yet served by the CUDA library ----
def test_binned_index_golden():
    """vcfc_oracle_build_binned_index == the .vcfci files written by `main_release create-binned-index`
    (oracle/make_golden_index.py): plain SNP files, edge-case lines, structural variants over several chromosomes."""
    idir = os.path.join(goldenlib.GOLDEN, "index")
    man = json.load(open(os.path.join(idir, "MANIFEST.json")))
    assert len(man) >= 12
    for fn, e in man.items():
        name = fn.split(".bin")[0]
        vcfc = open(os.path.join(idir, name + ".vcfc"), "rb").read() if name.startswith("sv_") else goldenlib.read(name + ".vcfc")
        rc, idx = O.build_binned_index(vcfc, e["entries_per_bin"])
        assert rc == e["entries"], fn
        assert idx == open(os.path.join(idir, fn), "rb").read(), fn


@pytest.mark.skipif(not O.have_ref_binary(), reason="oracle/_ref/main_release not built")
def test_binned_index_live_reference(tmp_path):
    h, d = vcfgen.kg_like(150, 40, seed=21)
    ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
    open(ip, "wb").write(h + d)
    assert subprocess.run([O.REF_BIN, "compress", ip, op], capture_output=True).returncode == 0
    vcfc = open(op, "rb").read()
    for b in (1, 2, 7, 1000):
        assert subprocess.run([O.REF_BIN, "create-binned-index", str(b), op], capture_output=True).returncode == 0
        rc, idx = O.build_binned_index(vcfc, b)
        assert rc > 0 and idx == open(op + ".vcfci", "rb").read(), b


@pytest.mark.skipif(not O.have_ref_binary(), reason="oracle/_ref/main_release not built")
def test_binned_index_fuzz_vs_reference_binary(tmp_path):
    """Seeded random ALT / INFO / POS columns, including ones the reference throws on (bad integers, "a=b=c" pairs): the
    oracle fails exactly when the reference aborts, and writes the same index bytes otherwise."""
    import random
    rng = random.Random(99)
    alts = ["C", "G,T", "ACGTT,A", "<DEL>", "<DUP>", "<INS:ME:ALU>", "<CN0>,<CN2>", "A,<DEL>", ""]
    infos = ["AC=1", "SVTYPE=DEL;END=%d", "END=%d,%d", "SVLEN=-%d", "SVLEN=%d,-%d;END", "END=", "END", "SVLEN=;X=1", ";;AC=2;",
             "END=12x", "SVLEN=abc", "A=B=C", "=", "END= 77", "END=+%d", "END=-5", "DB;H2;END=%d", "END=%d;END=%d"]
    n_fail = n_ok = 0
    for case in range(60):
        risky = case % 3 == 0
        lines, pos = [], 1000
        for i in range(rng.randrange(1, 30)):
            pos += rng.randrange(1, 900)
            info = rng.choice(infos if risky else infos[:9] + infos[13:])
            info = info.replace("%d", "{}").format(*[pos + rng.randrange(5000) for _ in range(info.count("%d"))])
            pos_s = str(pos) if not (risky and rng.random() < 0.05) else rng.choice(["12a", "", " 7", "-3"])
            lines.append("%s\t%s\trs%d\t%s\t%s\t9\tPASS\t%s\tGT\t0|0\t0|1\n" % (
                rng.choice(["1", "2", "X", "chrUn"]), pos_s, i, rng.choice(["A", "ACGT", "ACGTACGTAC"]), rng.choice(alts), info))
        vcf = vcfgen.header(2) + "".join(lines).encode()
        ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
        open(ip, "wb").write(vcf)
        assert subprocess.run([O.REF_BIN, "compress", ip, op], capture_output=True).returncode == 0
        vcfc = open(op, "rb").read()
        b = rng.choice([1, 2, 5, 100])
        if os.path.exists(op + ".vcfci"):
            os.remove(op + ".vcfci")
        r = subprocess.run([O.REF_BIN, "create-binned-index", str(b), op], capture_output=True)
        rc, idx = O.build_binned_index(vcfc, b)
        if r.returncode != 0:
            assert rc < 0, (case, rc)
            n_fail += 1
        else:
            assert rc >= 0 and idx == open(op + ".vcfci", "rb").read(), case
            n_ok += 1
    assert n_ok >= 30 and n_fail >= 3


def test_indexed_query_golden():
    """The restatement of query_binned_index_binarysearch (tests/oraclelib.py) prints what the reference binary printed
    for every committed (file, bin size, region) case: sorted and unsorted files, symbolic alleles, unknown chromosomes."""
    idir = os.path.join(goldenlib.GOLDEN, "index")
    cases = json.load(open(os.path.join(idir, "QUERIES.json")))
    assert len(cases) >= 60
    for key, e in cases.items():
        name, b, region = key.split("|")
        vcfc = open(os.path.join(idir, name + ".vcfc"), "rb").read() if name.startswith("sv_") else goldenlib.read(name + ".vcfc")
        index = open(os.path.join(idir, "%s.bin%s.vcfci" % (name, b)), "rb").read()
        out = O.query_binned_index(vcfc, index, region)
        assert hashlib.sha256(out).hexdigest() == e["sha256"] and out.count(b"\n") == e["lines"], key
        if e["file"]:
            assert out == open(os.path.join(idir, e["file"]), "rb").read(), key


def test_file_drivers(tmp_path):
    h, d = vcfgen.random_vcf_like(20, 50, seed=2)
    ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
    open(ip, "wb").write(h + d)
    assert O.lib().vcfc_oracle_compress_file(ip.encode(), op.encode()) == 0
    assert open(op, "rb").read() == O.compress_vcf(h + d)[1]
    assert O.lib().vcfc_oracle_decompress_file(op.encode(), rp.encode()) == 0
    assert open(rp, "rb").read() == h + d


def test_config1_exact_file_pins_the_oracle():
    """BASELINE.json configs[0]: the reference generator's 10k x 2504 file (restated in tests/config1gen.py; its sha256
    is the one recorded from the reference's own script) compresses to the sha256 the unmodified reference binary wrote
    (BASELINE.md), and round-trips."""
    import config1gen
    vcf = config1gen.generate()
    assert len(vcf) == config1gen.INPUT_LEN and hashlib.sha256(vcf).hexdigest() == config1gen.INPUT_SHA256
    hl = vcf.index(b"\n1\t") + 1
    rc, enc, nl, _ = O.compress_block(vcf[hl:])
    out = vcf[:hl] + enc
    assert rc == 0 and nl == 10000 and len(out) == config1gen.VCFC_LEN
    assert hashlib.sha256(out).hexdigest() == config1gen.VCFC_SHA256
    rc, txt = O.decompress_vcfc(out)
    assert rc == 0 and txt == vcf
