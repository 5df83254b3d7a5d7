"""ctypes binding of oracle/libvcfc_oracle.so -- the CPU parity oracle (TEST USE ONLY)."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "libvcfc_oracle.so")
REF_BIN = os.path.join(ODIR, "_ref", "main_release")

E_TOOFEW, E_EIGHTCOLS, E_CAP, E_FORMAT, E_TRUNC, E_IO, E_HEADER = -1, -2, -3, -4, -5, -6, -7


def build():
    src = os.path.join(ODIR, "vcfc_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, src])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        sz, u8p, szp = C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t)
        _lib.vcfc_oracle_compress_block.argtypes = [u8p, sz, C.c_void_p, sz, szp, C.c_void_p, szp, szp]
        _lib.vcfc_oracle_compress_block.restype = C.c_int
        _lib.vcfc_oracle_decompress_block.argtypes = [u8p, sz, C.c_uint64, C.c_void_p, sz, szp, szp, szp]
        _lib.vcfc_oracle_decompress_block.restype = C.c_int
        _lib.vcfc_oracle_parse_headers.argtypes = [u8p, sz, C.POINTER(C.c_uint64)]
        _lib.vcfc_oracle_parse_headers.restype = C.c_long
        _lib.vcfc_oracle_compress_file.argtypes = [C.c_char_p, C.c_char_p]
        _lib.vcfc_oracle_build_binned_index.argtypes = [u8p, sz, C.c_uint64, C.c_void_p, sz, szp]
        _lib.vcfc_oracle_build_binned_index.restype = C.c_long
        _lib.vcfc_oracle_decompress_file.argtypes = [C.c_char_p, C.c_char_p]
    return _lib


def compress_block(data: bytes, want_offsets=False):
    """-> (rc, out_bytes, n_lines, err_line[, offsets])"""
    n_max = data.count(b"\n") + 2
    cap = 2 * len(data) + 16 * n_max + 64
    out = C.create_string_buffer(cap)
    offs = (C.c_uint64 * n_max)()
    olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = lib().vcfc_oracle_compress_block(data, len(data), out, cap, C.byref(olen), offs, C.byref(nl), C.byref(el))
    res = (rc, out.raw[:olen.value], nl.value, el.value)
    if want_offsets:
        res += (list(offs[:nl.value]),)
    return res


def decompress_block(data: bytes, sample_count: int, cap=None):
    cap = cap or (len(data) * 520 + 4096)   # a 0x7f token byte expands to 127 * 4 bytes
    out = C.create_string_buffer(cap)
    olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = lib().vcfc_oracle_decompress_block(data, len(data), sample_count, out, cap, C.byref(olen), C.byref(nl), C.byref(el))
    return rc, out.raw[:olen.value], nl.value, el.value


def parse_headers(data: bytes):
    sc = C.c_uint64(0)
    n = lib().vcfc_oracle_parse_headers(data, len(data), C.byref(sc))
    return n, sc.value


def compress_vcf(vcf: bytes):
    """Whole-file compress() semantics on bytes: '#' lines pass through, data lines encoded."""
    out = bytearray()
    pos = 0
    while pos < len(vcf):
        e = vcf.find(b"\n", pos)
        e = len(vcf) if e < 0 else e
        line = vcf[pos:e]
        pos = e + 1
        if not line:
            continue
        if line[:1] == b"#":
            out += line + b"\n"
        else:
            rc, enc, _, _ = compress_block(line + b"\n")
            if rc != 0:
                return rc, bytes(out)
            out += enc
    return 0, bytes(out)


def decompress_vcfc(vcfc: bytes):
    n, sc = parse_headers(vcfc)
    if n < 0:
        return n, b""
    rc, txt, _, _ = decompress_block(vcfc[n:], sc)
    return rc, vcfc[:n] + txt


def have_ref_binary():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def build_binned_index(vcfc: bytes, entries_per_bin: int):
    """Whole .vcfc file -> (rc_or_entry_count, .vcfci bytes); restates create_binned_index4 (main.cpp:1284-1637)."""
    cap = 13 * (vcfc.count(b"\n") + 2)
    out = C.create_string_buffer(cap)
    olen = C.c_size_t(0)
    rc = lib().vcfc_oracle_build_binned_index(vcfc, len(vcfc), entries_per_bin, out, cap, C.byref(olen))
    return rc, out.raw[:olen.value]


# ---- indexed range query: pure-Python restatement (small cases only) of query_binned_index_binarysearch,
#      /root/reference/src/main.cpp:2974-3350, with compare_to_range (main.cpp:108-140) and compute_end_position
#      (main.cpp:763-852).  Pinned by tests/golden/index/QUERIES.json (outputs of the reference binary). ----
_REFS = {**{str(i): i for i in range(1, 23)}, "X": 23, "Y": 24, "M": 25}


def _strtoul_field(b: bytes):
    """str_to_uint64 / str_to_long (utils.cpp:152-175): strtoul must consume the whole field; None on failure."""
    import re
    if b == b"":
        return 0
    m = re.fullmatch(rb"[ \t\n\v\f\r]*([+-]?)([0-9]+)", b)
    if not m:
        return None
    v = min(int(m.group(2)), 2**64 - 1)
    if m.group(1) == b"-" and v != 2**64 - 1:
        v = (-v) % 2**64
    return v - 2**64 if v >= 2**63 else v          # stored in a long


def _kvp(info: bytes):
    d = {}
    for pair in info.split(b";"):
        if not pair:
            continue
        parts = [x for x in pair.split(b"=") if x]
        if len(parts) == 2:
            d[parts[0]] = parts[1]
        elif len(parts) == 1:
            d[parts[0]] = b""
        else:
            raise ValueError("Invalid kvp format")
    return d


def end_position(pos: int, ref: bytes, alt: bytes, info: bytes) -> int:
    if b"<" in alt:
        kv = _kvp(info)
        if b"END" in kv:
            m = 0
            for t in kv[b"END"].split(b","):
                if t:
                    v = _strtoul_field(t)
                    if v is None:
                        raise ValueError("END")
                    m = max(m, v)
            return abs(m)
        if b"SVLEN" in kv:
            m = 0
            for t in kv[b"SVLEN"].split(b","):
                if t:
                    v = _strtoul_field(t)
                    if v is None:
                        raise ValueError("SVLEN")
                    m = max(m, abs(v))
            return pos + m - 1
        return pos
    longest = max((len(a) for a in alt.split(b",")), default=0)
    return pos + max(len(ref), longest) - 1


def query_binned_index(vcfc: bytes, index: bytes, region: str) -> bytes:
    """Text the reference prints for `query-binned-index file REGION` (index entries of 13 bytes)."""
    import struct
    if ":" in region:
        name, rng = region.split(":", 1)
        a, b = rng.split("-", 1)
        q_start, q_end = int(a), int(b)
    else:
        name, q_start, q_end = region, 0, 0
    qidx = _REFS.get(name, 0)
    sc = C.c_uint64(0)
    hlen = lib().vcfc_oracle_parse_headers(vcfc, len(vcfc), C.byref(sc))
    assert hlen >= 0 and len(index) % 13 == 0
    count = len(index) // 13
    if count == 0:
        return b""
    rd = lambda i: struct.unpack_from("<BIQ", index, 13 * i)
    greater = lambda e: e[0] > qidx or (e[0] == qidx and e[1] > q_start)
    less = lambda e: e[0] < qidx or (e[0] == qidx and e[1] < q_start)
    lo, hi = 0, count - 1
    mid = (lo + hi) // 2
    entry = rd(0)                                   # (the reference's struct is uninitialised when count == 1)
    while lo < hi:
        mid = (lo + hi) // 2
        entry = rd(mid)
        if entry[0] == qidx and entry[1] == q_start:
            break
        if greater(entry):
            if mid == 0:
                break
            hi = mid - 1
        elif less(entry):
            lo = mid + 1
    if mid > 0 and greater(entry):
        mid -= 1
        entry = rd(mid)
    pos, hits = entry[2], []
    while len(vcfc) - pos >= 8:
        ll = ((vcfc[pos] & 0x3F) << 24) | (vcfc[pos + 1] << 16) | (vcfc[pos + 2] << 8) | vcfc[pos + 3]
        cols = vcfc[pos + 8: pos + 4 + ll].split(b"\t", 8)
        lpos = _strtoul_field(cols[1])
        info = cols[7] if b"<" in cols[4] else b""
        lend = end_position(lpos, cols[3], cols[4], info)
        lidx = _REFS.get(cols[0].decode("latin1"), 0)
        if lidx < qidx or (lidx == qidx and (lend % 2**64) < q_start):
            pass                                    # before the query
        elif lidx > qidx or (lidx == qidx and lpos > q_end):
            break                                   # behind it
        else:
            hits.append(vcfc[pos: pos + 4 + ll])
        pos += 4 + ll
    if not hits:
        return b""
    rc, txt, _, _ = decompress_block(b"".join(hits), sc.value)
    assert rc == 0
    return txt
