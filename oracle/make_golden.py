#!/usr/bin/env python3
"""Generate tests/golden/ from the UNMODIFIED reference (run in the build container only).

Needs /root/reference (for other/random_vcf.py) and oracle/_ref/main_release (built by
oracle/Makefile from the reference sources).  Every fixture is
    <name>.vcf[.gz]      input
    <name>.vcfc[.gz]     output of `main_release compress`
    <name>.rt.sha256     sha256 of `main_release decompress` output (only when it round-trips)
plus MANIFEST.json describing each case (sizes, sha256, reference exit codes, query cases).
The GPU box never runs this script; it only reads the committed fixtures.
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VCFC_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "main_release")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vcfgen  # noqa: E402


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def reference_random_vcf(sample_count: int, variant_count: int, workdir: str) -> bytes:
    """Execute the reference generator with its two shape constants overridden."""
    src = open(os.path.join(REF, "other", "random_vcf.py")).read()
    src = src.replace("sample_count = 1000", f"sample_count = {sample_count}")
    src = src.replace("variant_count = 1000000", f"variant_count = {variant_count}")
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        exec(compile(src, "random_vcf.py", "exec"), {"__name__": "__main__"})
    finally:
        os.chdir(cwd)
    return open(os.path.join(workdir, f"test-{sample_count}-{variant_count}.vcf"), "rb").read()


def run_ref(args, workdir):
    p = subprocess.run([BIN] + args, cwd=workdir, capture_output=True)
    return p.returncode, p.stdout


def write(name: str, data: bytes, gz: bool):
    path = os.path.join(OUT, name + (".gz" if gz else ""))
    if gz:
        with open(path, "wb") as f:
            f.write(gzip.compress(data, 9, mtime=0))
    else:
        with open(path, "wb") as f:
            f.write(data)


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    hdr8 = vcfgen.header(8, prefix="S")
    cases = {}

    with tempfile.TemporaryDirectory() as wd:
        # 1. the reference's own generator (byte-exact Python `random` stream, seed 5)
        cases["refgen_300x40"] = reference_random_vcf(300, 40, wd)
        cases["refgen_2504x24"] = reference_random_vcf(2504, 24, wd)
        # 2. SURVEY.md 8(c) known-answer lines
        kat = [
            b"1\t100\trs1\tA\tT\t100\tPASS\tAC=1\tGT\t0|0\t0|0\t0|0\t0|1\t0|1\t1|0\t1|1\t0|0\n",
            b"1\t102\trs2\tA\tT,G\t100\tPASS\tAC=1\tGT\t0|2\t0|0\t./.\t0/0\t0/1\t2|2\t1|1\t2|1\n",
            b"1\t104\trs3\tA\tT\t100\tPASS\tAC=1\tGT:DP\t" + b"\t".join([b"0|0:3", b"0|0:3", b"0|1:9", b"0|0:3"] * 2) + b"\n",
            b"1\t106\trs4\tA\tT\t100\tPASS\tAC=1\tGT\t" + b"\t".join([b"0|0"] * 8) + b"\n",
        ]
        cases["kat_survey8c"] = hdr8 + b"".join(kat)
        # 3. edge cases (decodable)
        cases["edge_8samples"] = hdr8 + b"".join(vcfgen.edge_case_lines(8))
        cases["edge_runs_700"] = vcfgen.header(700) + b"".join(vcfgen.run_length_lines())
        crlf = b"".join(l[:-1] + b"\r\n" for l in vcfgen.edge_case_lines(8)[:4])
        cases["edge_crlf"] = hdr8.replace(b"\n", b"\r\n") + crlf
        cases["edge_nosamples_9col"] = vcfgen.header(0) + b"1\t5\t.\tA\tC\t.\t.\t.\tGT\n1\t6\t.\tA\tC\t.\t.\tX=1\tGT\n"
        cases["edge_1sample"] = vcfgen.header(1) + b"".join(
            b"1\t%d\t.\tA\tC\t.\t.\t.\tGT\t%s\n" % (i, g) for i, g in enumerate([b"0|0", b"1|1", b"./.", b"0|1", b"0|0"]))
        # 4. inputs whose text the reference normalises (round trip differs by design)
        cases["norm_blank_and_notrailingnl"] = hdr8 + kat[0] + b"\n\n" + kat[3][:-1]
        cases["norm_trailing_tab"] = hdr8 + kat[0][:-1] + b"\t\n" + kat[3]
        # 5. encodes "successfully" but the reference decoder rejects it
        cases["undecodable_empty_field"] = hdr8 + kat[0].replace(b"\t0|1\t0|1", b"\t0|1\t\t0|1")
        # 6. kg-like medium case, gzip'd
        h, d = vcfgen.kg_like(60, 2504, seed=20)
        cases["kg_2504x60"] = h + d

        for name, vcf in cases.items():
            ip = os.path.join(wd, name + ".vcf")
            op = os.path.join(wd, name + ".vcfc")
            rp = os.path.join(wd, name + ".rt")
            open(ip, "wb").write(vcf)
            rc, _ = run_ref(["compress", ip, op], wd)
            entry = {"vcf_sha256": sha(vcf), "vcf_len": len(vcf), "compress_rc": rc}
            gz = len(vcf) > 64 * 1024
            entry["gz"] = gz
            write(name + ".vcf", vcf, gz)
            if rc == 0:
                vcfc = open(op, "rb").read()
                entry.update(vcfc_sha256=sha(vcfc), vcfc_len=len(vcfc))
                write(name + ".vcfc", vcfc, gz)
                rc2, _ = run_ref(["decompress", op, rp], wd)
                entry["decompress_rc"] = rc2
                if rc2 == 0:
                    rt = open(rp, "rb").read()
                    entry["roundtrip_identical"] = rt == vcf
                    entry["rt_sha256"] = sha(rt)
                    if rt != vcf:
                        write(name + ".rt", rt, gz)
            manifest[name] = entry

        # 7. inputs the reference aborts on (no output kept)
        for name, line in {
            "abort_8cols": b"1\t5\t.\tA\tC\t.\t.\t.\n",
            "abort_7cols": b"1\t5\t.\tA\tC\t.\t.\n",
        }.items():
            ip = os.path.join(wd, name + ".vcf")
            open(ip, "wb").write(hdr8 + kat[0] + line)
            rc, _ = run_ref(["compress", ip, ip + "c"], wd)
            write(name + ".vcf", hdr8 + kat[0] + line, False)
            manifest[name] = {"compress_rc": rc, "gz": False}

        # 8. range queries (query prints matching data lines to stdout; main.cpp:3777-3929)
        queries = {}
        for name, q in [("refgen_300x40", "1:10010-10030"), ("refgen_300x40", "1:0-99999999"),
                        ("refgen_300x40", "2:0-99999999"), ("kg_2504x60", "20:60000-60600"),
                        ("edge_8samples", "1:104-110")]:
            op = os.path.join(wd, name + ".vcfc")
            rc, so = run_ref(["query", op, q], wd)
            queries.setdefault(name, []).append({"q": q, "rc": rc, "sha256": sha(so), "len": len(so)})
        manifest["_queries"] = queries

    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    for k, v in manifest.items():
        print(k, {kk: vv for kk, vv in v.items() if "sha" not in kk} if isinstance(v, dict) else v)


if __name__ == "__main__":
    main()
