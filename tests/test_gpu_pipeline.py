"""GPU tests of the file pipeline (csrc/vcfc_pipeline.cu): pinned chunk ring, reader / worker / writer threads, several
contexts with host concatenation by chunk offsets.  Expected bytes come from the oracle
(oraclelib.compress_vcf restates compress(), /root/reference/src/compress.cpp:205-257) and the golden files written by
the reference binary."""
import importlib
import os
import subprocess

import pytest

import oraclelib as O
import vcfgen

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("vcf-compression_b200")


@pytest.fixture(scope="module")
def codecs():
    cs = [pkg.Codec(0) for _ in range(3)]          # three contexts on one device: the worker threads are what is tested
    yield cs
    for c in cs:
        c.close()


def _roundtrip(codecs, vcf: bytes, tmp_path, n_ctx, expect_rt=None):
    ip, op, rp = (str(tmp_path / f"{n_ctx}_{x}") for x in ("a.vcf", "a.vcfc", "a.rt"))
    open(ip, "wb").write(vcf)
    orc, want = O.compress_vcf(vcf)
    assert orc == 0
    assert pkg.Codec.compress_multi(codecs[:n_ctx], ip, op) == 0
    got = open(op, "rb").read()
    assert got == want, f"compressed file differs with {n_ctx} context(s)"
    assert pkg.Codec.decompress_multi(codecs[:n_ctx], op, rp) == 0
    assert open(rp, "rb").read() == (vcf if expect_rt is None else expect_rt)
    return got


@pytest.mark.parametrize("n_ctx", [1, 2, 3])
def test_many_chunks_equal_the_oracle_file(codecs, tmp_path, monkeypatch, n_ctx):
    """~25 MB in 1 MB chunks over 1-3 worker threads: output bytes do not depend on the chunking or the worker count."""
    monkeypatch.setenv("VCFC_FILE_CHUNK_MB", "1")
    monkeypatch.setenv("VCFC_FILE_DCHUNK_MB", "1")
    h, d = vcfgen.kg_like(2500, 2504, seed=77)
    _roundtrip(codecs, h + d, tmp_path, n_ctx)


def test_hash_lines_and_blank_lines_anywhere(codecs, tmp_path, monkeypatch):
    """'#' lines between data lines pass through, blank lines vanish, a missing final newline is supplied
    (compress.cpp:219-238) -- also when they fall next to chunk boundaries."""
    monkeypatch.setenv("VCFC_FILE_CHUNK_MB", "1")
    h, d = vcfgen.random_vcf_like(420, 2504, seed=3)
    lines = d.split(b"\n")[:-1]
    body = []
    for i, ln in enumerate(lines):
        body.append(ln)
        if i % 97 == 5:
            body.append(b"")
        if i % 101 == 7:
            body.append(b"##late=meta line %d" % i)
    vcf = h + b"\n".join(body)                                      # no final newline
    orc, want = O.compress_vcf(vcf)
    assert orc == 0
    ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
    open(ip, "wb").write(vcf)
    for n_ctx in (1, 3):
        assert pkg.Codec.compress_multi(codecs[:n_ctx], ip, op) == 0
        assert open(op, "rb").read() == want


def test_golden_files_through_the_pipeline(codecs, golden, tmp_path):
    for name, g in golden.items():
        ip, op, rp = (str(tmp_path / f"{name}.{x}") for x in ("vcf", "vcfc", "rt"))
        open(ip, "wb").write(g["vcf"])
        rc = pkg.Codec.compress_multi(codecs[:2], ip, op)
        if g["vcfc"] is None:
            assert rc != 0, name                                     # the reference aborts on these
            continue
        assert rc == 0 and open(op, "rb").read() == g["vcfc"], name
        rc = pkg.Codec.decompress_multi(codecs[:2], op, rp)
        if g["entry"].get("decompress_rc", 0) != 0:
            assert rc != 0, name
        else:
            assert rc == 0 and open(rp, "rb").read() == g["rt"], name


def test_chrom_line_with_too_few_columns_fails(codecs, tmp_path):
    """compress.cpp:227-233 throws "VCF Header did not have enough columns" for a '#' line with fewer than 8 terms."""
    h, d = vcfgen.random_vcf_like(3, 8, seed=1)
    meta = b"".join(ln + b"\n" for ln in h.split(b"\n") if ln.startswith(b"##"))
    ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
    open(ip, "wb").write(meta + b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\n" + d)
    assert codecs[0].compress(ip, op) == pkg.E_HEADER
    assert open(op, "rb").read() == meta                            # what the reference had written before it threw
    open(ip, "wb").write(meta + b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    assert codecs[0].compress(ip, op) == 0                           # exactly 8 terms: accepted (compress.cpp:235)
    if O.have_ref_binary():
        open(ip, "wb").write(meta + b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\n" + d)
        r = subprocess.run([O.REF_BIN, "compress", ip, op + ".ref"], capture_output=True)
        if r.returncode not in (0, 127):
            assert r.returncode != 0


def test_error_in_the_middle_keeps_the_lines_before_it(codecs, tmp_path, monkeypatch):
    """A data line with fewer than 8 columns aborts the reference at that line; everything before it is in the file."""
    monkeypatch.setenv("VCFC_FILE_CHUNK_MB", "1")
    h, d = vcfgen.random_vcf_like(600, 2504, seed=9)
    cut = d.rfind(b"\n", 0, 3_500_000) + 1
    vcf = h + d[:cut] + b"oops\n" + d[cut:]
    orc, want = O.compress_vcf(vcf)
    assert orc != 0
    ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
    open(ip, "wb").write(vcf)
    for n_ctx in (1, 3):
        rc = pkg.Codec.compress_multi(codecs[:n_ctx], ip, op)
        assert rc == pkg.E_TOOFEW
        assert open(op, "rb").read() == want


def test_cli_over_two_contexts_env(tmp_path, golden):
    if not os.path.exists(pkg.CLI_PATH):
        pytest.skip("CLI not built")
    g = golden["refgen_300x40"]
    ip, op, rp = (str(tmp_path / x) for x in ("a.vcf", "a.vcfc", "a.rt"))
    open(ip, "wb").write(g["vcf"])
    env = dict(os.environ, VCFC_GPUS="all")
    assert subprocess.run([pkg.CLI_PATH, "compress", ip, op], env=env).returncode == 0
    assert open(op, "rb").read() == g["vcfc"]
    assert subprocess.run([pkg.CLI_PATH, "decompress", op, rp], env=env).returncode == 0
    assert open(rp, "rb").read() == g["rt"]


def test_fused_index_equals_the_separate_pass_and_the_oracle(codecs, tmp_path, monkeypatch):
    """compress + create-binned-index in one pass (index fields from the encoder's own line offsets, chunk by chunk, any
    number of contexts) writes the bytes of create_binned_index4 (main.cpp:1284-1637) on the finished file."""
    monkeypatch.setenv("VCFC_FILE_CHUNK_MB", "1")
    h, d = vcfgen.kg_like(3000, 500, seed=41)                       # ~6 MB: several chunks
    vp, cp = str(tmp_path / "k.vcf"), str(tmp_path / "k.vcfc")
    open(vp, "wb").write(h + d)
    orc, want = O.compress_vcf(h + d)
    assert orc == 0
    for n_ctx in (1, 3):
        for b in (1, 7, 1000):
            rc, n = pkg.Codec.compress_index_multi(codecs[:n_ctx], vp, cp, cp + ".vcfci", b)
            assert rc == 0 and open(cp, "rb").read() == want
            on, oidx = O.build_binned_index(want, b)
            assert n == on and open(cp + ".vcfci", "rb").read() == oidx, (n_ctx, b)
            rc2, n2 = codecs[0].create_binned_index(cp, cp + ".sep", b)
            assert rc2 == 0 and n2 == n and open(cp + ".sep", "rb").read() == oidx
    # structural variants (INFO END / SVLEN decide the END position): the committed fixture of the reference binary
    import json
    import goldenlib
    idir = os.path.join(goldenlib.GOLDEN, "index")
    man = json.load(open(os.path.join(idir, "MANIFEST.json")))
    for fn, e in man.items():
        name = fn.split(".bin")[0]
        if name.startswith("sv_") or not goldenlib.exists(name + ".vcf"):
            continue
        open(vp, "wb").write(goldenlib.read(name + ".vcf"))
        rc, n = pkg.Codec.compress_index_multi(codecs[:2], vp, cp, cp + ".vcfci", e["entries_per_bin"])
        assert rc == 0 and n == e["entries"], fn
        assert open(cp + ".vcfci", "rb").read() == open(os.path.join(idir, fn), "rb").read(), fn
    # a '#' line behind a data line: the compressed file stands, the index is refused (the reference's walk fails there)
    open(vp, "wb").write(h + d[:200000].rsplit(b"\n", 1)[0] + b"\n##late\n" + d[:5000].rsplit(b"\n", 1)[0] + b"\n")
    rc, _ = pkg.Codec.compress_index_multi(codecs[:1], vp, cp, cp + ".bad", 10)
    assert rc == pkg.E_FORMAT and open(cp, "rb").read() == O.compress_vcf(open(vp, "rb").read())[1]
    # the CLI: VCFC_INDEX_BIN
    open(vp, "wb").write(h + d)
    env = dict(os.environ, VCFC_INDEX_BIN="7")
    assert subprocess.run([pkg.CLI_PATH, "compress", vp, cp], env=env).returncode == 0
    assert open(cp + ".vcfci", "rb").read() == O.build_binned_index(want, 7)[1]
