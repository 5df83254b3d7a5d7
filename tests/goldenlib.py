"""Access to tests/golden/ (fixtures produced by the reference binary, see oracle/make_golden.py)."""
import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    return json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))


def read(name: str) -> bytes:
    p = os.path.join(GOLDEN, name)
    if os.path.exists(p):
        return open(p, "rb").read()
    return gzip.decompress(open(p + ".gz", "rb").read())


def exists(name: str) -> bool:
    p = os.path.join(GOLDEN, name)
    return os.path.exists(p) or os.path.exists(p + ".gz")


def load_all():
    """name -> dict(vcf=..., vcfc=... or None, rt=... (expected round trip), entry=manifest row)."""
    out = {}
    for name, e in manifest().items():
        if name.startswith("_"):
            continue
        vcf = read(name + ".vcf")
        vcfc = read(name + ".vcfc") if exists(name + ".vcfc") else None
        rt = read(name + ".rt") if exists(name + ".rt") else vcf
        out[name] = dict(vcf=vcf, vcfc=vcfc, rt=rt, entry=e)
    return out


def split_header(vcf: bytes):
    """(header region, data-line region): '#' lines come first in every fixture."""
    pos = 0
    while pos < len(vcf) and vcf[pos:pos + 1] == b"#":
        pos = vcf.index(b"\n", pos) + 1
    return vcf[:pos], vcf[pos:]
