"""The reference's OWN binary on the GPU path: oracle/_ref_gpu/main_gpu is the reference's main.cpp / utils.cpp /
sparse.cpp (compiled where they lie, oracle/Makefile target ref_gpu) linked with host/compress_gpu.cpp in place of
src/compress.cpp, against libvcfc_gpu.so -- the functions of src/compress.hpp:17-56 with their exact signatures.
Its verbs must produce the bytes the unmodified reference binary produced (tests/golden/, written by oracle/_ref/main_release)."""
import hashlib
import importlib
import json
import os
import subprocess

import pytest

import goldenlib

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("vcf-compression_b200")
MAIN_GPU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref_gpu", "main_gpu")


@pytest.fixture(scope="module")
def main_gpu():
    if not (os.path.exists(MAIN_GPU) and os.access(MAIN_GPU, os.X_OK)):
        pytest.skip("oracle/_ref_gpu/main_gpu not shipped (needs /root/reference at build time)")
    r = subprocess.run(["ldd", MAIN_GPU], capture_output=True, text=True)
    assert "libvcfc_gpu.so" in r.stdout and "not found" not in r.stdout, r.stdout
    return MAIN_GPU


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, timeout=300)


def test_compress_and_decompress_verbs(main_gpu, golden, tmp_path):
    for name, g in golden.items():
        ip, op, rp = (str(tmp_path / f"{name}.{x}") for x in ("vcf", "vcfc", "rt"))
        open(ip, "wb").write(g["vcf"])
        r = run(main_gpu, "compress", ip, op)
        if g["entry"]["compress_rc"] != 0:
            assert r.returncode != 0, name                    # the reference aborts on these (uncaught exception)
            continue
        assert r.returncode == 0, (name, r.stderr[-300:])
        assert open(op, "rb").read() == g["vcfc"], name
        r = run(main_gpu, "decompress", op, rp)
        if g["entry"].get("decompress_rc", 0) != 0:
            assert r.returncode != 0, name
            continue
        assert r.returncode == 0, (name, r.stderr[-300:])
        assert open(rp, "rb").read() == g["rt"], name


def test_query_verb_decodes_line_by_line_through_the_gpu(main_gpu, golden, tmp_path):
    """query_compressed_file (main.cpp:3777-3929) stays the reference's code; every matching line reaches the GPU through
    decompress2_data_line_FILEwrapper (one-line blocks)."""
    qs = goldenlib.manifest()["_queries"]
    for name, cases in qs.items():
        fp = str(tmp_path / (name + ".vcfc"))
        open(fp, "wb").write(golden[name]["vcfc"])
        for c in cases:
            r = run(main_gpu, "query", fp, c["q"])
            assert r.returncode == c["rc"], (name, c["q"], r.stderr[-300:])
            if c["rc"] == 0:
                assert len(r.stdout) == c["len"] and hashlib.sha256(r.stdout).hexdigest() == c["sha256"], (name, c["q"])


def test_binned_index_verbs(main_gpu, tmp_path):
    """create-binned-index / query-binned-index of the reference's main.cpp (1284-1637, 2974-3350) over the shim's
    read_compressed_line_length_headers / decompress2_metadata_headers / decompress2_data_line."""
    idir = os.path.join(goldenlib.GOLDEN, "index")
    man = json.load(open(os.path.join(idir, "MANIFEST.json")))
    cases = json.load(open(os.path.join(idir, "QUERIES.json")))
    for fn, e in man.items():
        name = fn.split(".bin")[0]
        vcfc = open(os.path.join(idir, name + ".vcfc"), "rb").read() if name.startswith("sv_") else goldenlib.read(name + ".vcfc")
        cp = str(tmp_path / "g.vcfc")
        open(cp, "wb").write(vcfc)
        r = run(main_gpu, "create-binned-index", str(e["entries_per_bin"]), cp)
        assert r.returncode == 0, (fn, r.stderr[-300:])
        assert open(cp + ".vcfci", "rb").read() == open(os.path.join(idir, fn), "rb").read(), fn
        for key, q in cases.items():
            qname, b, region = key.split("|")
            if qname != name or int(b) != e["entries_per_bin"]:
                continue
            r = run(main_gpu, "query-binned-index", cp, region)
            assert r.returncode == 0 and hashlib.sha256(r.stdout).hexdigest() == q["sha256"], key


def test_sparse_verbs(main_gpu, tmp_path):
    """sparsify / sparse-query / create-sparse-index / query-sparse-index (sparse.cpp:290-580, main.cpp:235-582, 854-1281) stay
    the reference's host code (lseek / SEEK_DATA over a ~4.9 TB holey file); every header walk and every decoded line goes
    through the shim into libvcfc_gpu.so.  Pinned by what the unmodified reference binary produced (oracle/make_golden_sparse.py)."""
    man = json.load(open(os.path.join(goldenlib.GOLDEN, "sparse", "MANIFEST.json")))
    for name, e in man.items():
        fp, sp = str(tmp_path / (name + ".vcfc")), str(tmp_path / (name + ".sparse"))
        open(fp, "wb").write(goldenlib.read(name + ".vcfc"))
        r = run(main_gpu, "sparsify", fp, sp)
        assert r.returncode == e["sparsify_rc"], (name, r.stderr[-300:])
        assert goldenlib.sparse_digest(sp) == e["sparse"], name
        for q, c in e["queries"].items():
            r = run(main_gpu, "sparse-query", sp, q)
            assert r.returncode == c["rc"], (name, q, r.stderr[-300:])
            if c["rc"] == 0:
                assert len(r.stdout) == c["len"] and hashlib.sha256(r.stdout).hexdigest() == c["sha256"], (name, q)
        r = run(main_gpu, "create-sparse-index", fp)
        assert r.returncode == e["index_rc"], (name, r.stderr[-300:])
        assert goldenlib.sparse_digest(fp + ".vcfci-sparse") == e["index"], name
        for q, c in e["index_queries"].items():
            r = run(main_gpu, "query-sparse-index", fp, q)
            assert r.returncode == c["rc"], (name, q, r.stderr[-300:])
            if c["rc"] == 0:
                assert len(r.stdout) == c["len"] and hashlib.sha256(r.stdout).hexdigest() == c["sha256"], (name, q)
        for x in (sp, fp + ".vcfci-sparse"):
            os.remove(x)
