#!/usr/bin/env python3
"""Generate tests/golden/index/ from the UNMODIFIED reference (run in the build container only):
.vcfci files written by `main_release create-binned-index <bin> <file.vcfc>` (main.cpp:1284-1637) for some of
the committed .vcfc fixtures and for a structural-variant / multi-chromosome file made here.  They pin
oracle/vcfc_oracle.c:vcfc_oracle_build_binned_index, the oracle of the next scope row.
"""
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "main_release")
OUT = os.path.join(ROOT, "tests", "golden", "index")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenlib  # noqa: E402
import vcfgen  # noqa: E402


def sv_mix(n_lines=240, seed=5) -> bytes:
    rng = random.Random(seed)
    lines, pos = [], 1000
    for i in range(n_lines):
        chrom = rng.choice(["1", "1", "2", "X", "22", "GL000207.1", "M"])
        pos += rng.randrange(1, 500)
        kind = rng.randrange(7)
        if kind == 0:
            ref, alt, info = "A", "<DEL>", "SVTYPE=DEL;END=%d" % (pos + rng.randrange(5000))
        elif kind == 1:
            ref, alt, info = "A", "<DUP>", "SVTYPE=DUP;SVLEN=%d" % rng.randrange(-3000, 3000)
        elif kind == 2:
            ref, alt, info = "A", "<INS:ME:ALU>", "SVTYPE=ALU;TSD=null"
        elif kind == 3:
            ref, alt, info = "ACGTACGT", "A,AC", "AC=1;AF=0.5"
        elif kind == 4:
            ref, alt, info = "A", "ACGTTTT,G", "AC=1;;DB"
        elif kind == 5:
            ref, alt, info = "A", "<CN0>,<CN2>", "END=%d,%d;CS=x" % (pos + 10, pos + 999)
        else:
            ref, alt, info = "A", "<DEL>", "SVLEN=-%d,%d;END" % (rng.randrange(100), rng.randrange(100))
        gts = "\t".join(rng.choice(["0|0", "0|1", "1|1"]) for _ in range(6))
        lines.append("%s\t%d\trs%d\t%s\t%s\t100\tPASS\t%s\tGT\t%s\n" % (chrom, pos, i, ref, alt, info, gts))
    return vcfgen.header(6) + "".join(lines).encode()


def sv_sorted(per_chrom=60, seed=9) -> bytes:
    """The same mix, sorted by chromosome (in the index's order) and position: what the index is meant for."""
    rng = random.Random(seed)
    lines, i = [], 0
    for chrom in ["1", "2", "22", "X", "M"]:
        pos = 500
        for _ in range(per_chrom):
            pos += rng.randrange(1, 400)
            kind = rng.randrange(5)
            if kind == 0:
                ref, alt, info = "A", "<DEL>", "SVTYPE=DEL;END=%d" % (pos + rng.randrange(3000))
            elif kind == 1:
                ref, alt, info = "A", "<DUP>", "SVTYPE=DUP;SVLEN=%d" % rng.randrange(-2000, 2000)
            elif kind == 2:
                ref, alt, info = "ACGTACGT", "A,AC", "AC=1;AF=0.5"
            else:
                ref, alt, info = "A", rng.choice(["C", "G,T", "ACGTT"]), "AC=%d" % rng.randrange(9)
            gts = "\t".join(rng.choice(["0|0", "0|0", "0|1", "1|1", "0|2"]) for _ in range(12))
            lines.append("%s\t%d\trs%d\t%s\t%s\t100\tPASS\t%s\tGT\t%s\n" % (chrom, pos, i, ref, alt, info, gts))
            i += 1
    return vcfgen.header(12) + "".join(lines).encode()


QUERIES = {
    "sv_sorted": ["1:1-2000", "1:3000-3300", "2:1-100000", "22:2000-2600", "X:5000-9000", "M:1-999999", "M:700-800",
                  "1:999999-9999999", "7:1-1000", "Y:1-10", "2", "1:2500-2500", "X:1-400"],
    "sv_mix": ["1:1000-5000", "X:20000-30000", "M:0-99999999", "2:1-100000"],
    "kg_2504x60": ["20:60000-61000", "20:60500-60600", "20:1-50", "20:61500-99999999", "21:1-100"],
}


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    with tempfile.TemporaryDirectory() as wd:
        ip, op = os.path.join(wd, "a.vcf"), os.path.join(wd, "a.vcfc")
        open(ip, "wb").write(sv_mix())
        assert subprocess.run([BIN, "compress", ip, op], capture_output=True).returncode == 0
        open(os.path.join(OUT, "sv_mix.vcfc"), "wb").write(open(op, "rb").read())
        sources = {"sv_mix": open(op, "rb").read()}
        open(ip, "wb").write(sv_sorted())
        assert subprocess.run([BIN, "compress", ip, op], capture_output=True).returncode == 0
        open(os.path.join(OUT, "sv_sorted.vcfc"), "wb").write(open(op, "rb").read())
        sources["sv_sorted"] = open(op, "rb").read()
        queries = {}
        for name in ("refgen_300x40", "kg_2504x60", "edge_8samples"):
            sources[name] = goldenlib.read(name + ".vcfc")
        for name, vcfc in sources.items():
            open(op, "wb").write(vcfc)
            for b in (1, 4, 25):
                r = subprocess.run([BIN, "create-binned-index", str(b), op], capture_output=True)
                assert r.returncode == 0, (name, b, r.stderr[-300:])
                idx = open(op + ".vcfci", "rb").read()
                fn = "%s.bin%d.vcfci" % (name, b)
                open(os.path.join(OUT, fn), "wb").write(idx)
                manifest[fn] = {"source": name + ".vcfc" + ("" if name.startswith("sv_") else " (tests/golden/)"),
                                "entries_per_bin": b, "entries": len(idx) // 13}
                # query-binned-index (main.cpp:2974-3350) with this index
                for q in QUERIES.get(name, []):
                    r = subprocess.run([BIN, "query-binned-index", op, q], capture_output=True)
                    assert r.returncode == 0, (name, b, q)
                    qfn = "%s.bin%d.q%d.out" % (name, b, QUERIES[name].index(q))
                    if len(r.stdout) <= 4096:
                        open(os.path.join(OUT, qfn), "wb").write(r.stdout)
                    queries["%s|%d|%s" % (name, b, q)] = {"sha256": __import__("hashlib").sha256(r.stdout).hexdigest(),
                                                         "bytes": len(r.stdout), "lines": r.stdout.count(b"\n"),
                                                         "file": qfn if len(r.stdout) <= 4096 else None}
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    json.dump(queries, open(os.path.join(OUT, "QUERIES.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(manifest), "index fixtures to", OUT)


if __name__ == "__main__":
    main()
