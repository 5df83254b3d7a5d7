#!/usr/bin/env python3
"""Generate tests/golden/sparse/MANIFEST.json from the UNMODIFIED reference (run in the build container only):
for some committed .vcfc fixtures, what `main_release sparsify`, `sparse-query`, `create-sparse-index` and
`query-sparse-index` produce (sparse.cpp:290-580, main.cpp:235-582, 854-1281) -- the last scope row.
The holey output files (~4.9 TB logical) are pinned by a filesystem-independent digest of their non-zero bytes
(tests/goldenlib.py:sparse_digest); query outputs by length + sha256 + exit code.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "main_release")
OUT = os.path.join(ROOT, "tests", "golden", "sparse")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenlib  # noqa: E402

CASES = {
    "refgen_300x40": ["1:10010-10030", "1:10000-10000", "1:10004-10004", "1:10005-10005", "1:10011-10070", "1:0-99999999", "1:10078-10078",
                      "1:20000-30000", "1:10077-20000"],
    "edge_8samples": ["1:104-110", "1:100-100", "1:101-101", "1:100-120", "1:118-118"],
    "kg_2504x60": ["20:60000-60600", "20:60000-60000", "20:61000-62200", "20:100-200"],
}


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, timeout=300)


def main():
    os.makedirs(OUT, exist_ok=True)
    man = {}
    with tempfile.TemporaryDirectory() as td:
        for name, regions in CASES.items():
            fp = os.path.join(td, name + ".vcfc")
            open(fp, "wb").write(goldenlib.read(name + ".vcfc"))
            sp = os.path.join(td, name + ".sparse")
            r = run("sparsify", fp, sp)
            e = {"sparsify_rc": r.returncode, "sparse": goldenlib.sparse_digest(sp), "queries": {}, "index_queries": {}}
            for q in regions:
                r = run("sparse-query", sp, q)
                e["queries"][q] = {"rc": r.returncode, "len": len(r.stdout), "sha256": hashlib.sha256(r.stdout).hexdigest()}
            r = run("create-sparse-index", fp)
            e["index_rc"] = r.returncode
            e["index"] = goldenlib.sparse_digest(fp + ".vcfci-sparse")
            for q in regions:
                r = run("query-sparse-index", fp, q)
                e["index_queries"][q] = {"rc": r.returncode, "len": len(r.stdout), "sha256": hashlib.sha256(r.stdout).hexdigest()}
            man[name] = e
            print(name, e["sparsify_rc"], e["sparse"]["logical_size"], e["sparse"]["nonzero_bytes"],
                  {q: (v["rc"], v["len"]) for q, v in e["queries"].items()}, {q: (v["rc"], v["len"]) for q, v in e["index_queries"].items()})
            for x in (sp, fp + ".vcfci-sparse"):
                if os.path.exists(x):
                    os.remove(x)
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
