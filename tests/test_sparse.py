"""The reference's sparse-file verbs: vcfc_sparsify_file (host only: runs in the CPU suite) and
vcfc_sparse_query_file (decodes on the GPU) against what the UNMODIFIED reference binary produced
(tests/golden/sparse/MANIFEST.json, written by oracle/make_golden_sparse.py)."""
import hashlib
import importlib
import json
import os

import pytest

import goldenlib

pkg = importlib.import_module("vcf-compression_b200")
MAN = json.load(open(os.path.join(goldenlib.GOLDEN, "sparse", "MANIFEST.json")))


@pytest.mark.parametrize("name", sorted(MAN))
def test_sparsify_matches_the_reference(name, tmp_path):
    fp, sp = str(tmp_path / (name + ".vcfc")), str(tmp_path / (name + ".sparse"))
    open(fp, "wb").write(goldenlib.read(name + ".vcfc"))
    assert pkg.sparsify(fp, sp) == 0
    assert goldenlib.sparse_digest(sp) == MAN[name]["sparse"]
    os.remove(sp)


def test_sparsify_rejects_what_the_reference_aborts_on(tmp_path):
    fp, sp = str(tmp_path / "x.vcfc"), str(tmp_path / "x.sparse")
    open(fp, "wb").write(b"not a vcfc file\n")
    assert pkg.sparsify(fp, sp) != 0
    assert pkg.sparsify(str(tmp_path / "missing.vcfc"), sp) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MAN))
def test_sparse_query_matches_the_reference(name, tmp_path):
    fp, sp = str(tmp_path / (name + ".vcfc")), str(tmp_path / (name + ".sparse"))
    open(fp, "wb").write(goldenlib.read(name + ".vcfc"))
    assert pkg.sparsify(fp, sp) == 0
    with pkg.Codec(0) as codec:
        for q, c in MAN[name]["queries"].items():
            op = str(tmp_path / "q.out")
            fd = os.open(op, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
            rc = codec.sparse_query(sp, q, fd)
            os.close(fd)
            out = open(op, "rb").read()
            if c["rc"] != 0:
                assert rc != 0, (name, q)                 # the reference aborts (uncaught exception)
                continue
            assert rc == 0, (name, q, rc)
            assert len(out) == c["len"] and hashlib.sha256(out).hexdigest() == c["sha256"], (name, q)
    os.remove(sp)
