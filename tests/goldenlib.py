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


def sparse_digest(path: str) -> dict:
    """Filesystem-independent digest of a holey file: logical size, number of non-zero bytes and a sha256 over
    (offset, bytes) of every maximal run of non-zero bytes, found by walking the data extents (SEEK_DATA / SEEK_HOLE)."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    nz = 0
    size = os.path.getsize(path)
    fd = os.open(path, os.O_RDONLY)
    try:
        pos = 0
        while pos < size:
            try:
                a = os.lseek(fd, pos, os.SEEK_DATA)
            except OSError:
                break
            b = os.lseek(fd, a, os.SEEK_HOLE)
            o = a
            while o < b:
                buf = os.pread(fd, min(b - o, 1 << 24), o)
                if not buf:
                    break
                arr = np.frombuffer(buf, dtype=np.uint8)
                m = np.concatenate(([0], (arr != 0).astype(np.int8), [0]))
                d = np.diff(m)
                for s, e in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1)):
                    h.update(int(o + s).to_bytes(8, "little"))
                    h.update(int(e - s).to_bytes(8, "little"))
                    h.update(buf[s:e])
                    nz += int(e - s)
                o += len(buf)
            pos = b
    finally:
        os.close(fd)
    return {"logical_size": size, "nonzero_bytes": nz, "sha256": h.hexdigest()}
