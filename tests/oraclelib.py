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
