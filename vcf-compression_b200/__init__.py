"""vcf-compression_b200 -- ctypes binding of libvcfc_gpu.so (the C ABI in include/vcfc_gpu.h).

The product is the shared library (hand-written sm_100a CUDA behind a C ABI) and the `vcfc` CLI.
This module is the thin host-side mirror used by tests and bench.py: the same calls a C++ host
makes, with the reference's function names for the path kept as method names

    compress_block / decompress_block      <- compress_data_line / decompress2_data_line, block-wise
    compress / decompress2_fd / query      <- src/compress.hpp:17,26 ; src/main.cpp:3777

The directory name carries a hyphen, so import it with
``importlib.import_module("vcf-compression_b200")``.

There is no CPU implementation here: if the library is missing or no sm_100 device is usable,
every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("VCFC_LIB_PATH") or os.path.join(HERE, "libvcfc_gpu.so")   # (override: tuning builds)
CLI_PATH = os.path.join(HERE, "vcfc")
HEADER = os.path.join(ROOT, "include", "vcfc_gpu.h")

OK, E_TOOFEW, E_EIGHTCOLS, E_CAP, E_FORMAT, E_TRUNC, E_IO, E_HEADER, E_CUDA, E_ARG, E_LINE2BIG, E_QUERY = range(12)
PATH_FAST, PATH_GENERIC = 1, 2


class VcfcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vcfc error {code}: {msg}")
        self.code = code


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("reserved", C.c_int32), ("out_len", C.c_uint64),
                ("n_lines", C.c_uint64), ("err_line", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    """Loads the library; raises if it has not been built (python vcf-compression_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VcfcError(E_CUDA, f"{LIB_PATH} is not built; run `python vcf-compression_b200/build.py` (no CPU path exists)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, i = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
    szp, u64p = C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)
    sig = {
        "vcfc_gpu_init": (i, [i, C.POINTER(vp)]),
        "vcfc_gpu_destroy": (None, [vp]),
        "vcfc_strerror": (C.c_char_p, [i]),
        "vcfc_last_cuda_error": (C.c_char_p, [vp]),
        "vcfc_encode_bound": (sz, [sz]),
        "vcfc_encode_block": (i, [vp, vp, sz, vp, sz, szp, vp, sz, szp, u64p]),
        "vcfc_encode_block_dev": (i, [vp, vp, sz, vp, sz, vp, sz, vp, vp]),
        "vcfc_decode_block": (i, [vp, vp, sz, u64, vp, sz, szp, szp, u64p]),
        "vcfc_decode_block_dev": (i, [vp, vp, sz, u64, vp, sz, vp, vp]),
        "vcfc_decode_size_dev": (i, [vp, vp, sz, u64, C.POINTER(Result), vp]),
        "vcfc_fetch_result": (i, [vp, vp, C.POINTER(Result), vp]),
        "vcfc_parse_headers": (i, [vp, sz, szp, u64p]),
        "vcfc_compress_file": (i, [vp, C.c_char_p, C.c_char_p]),
        "vcfc_decompress_file": (i, [vp, C.c_char_p, C.c_char_p]),
        "vcfc_compress_file_multi": (i, [C.POINTER(vp), i, C.c_char_p, C.c_char_p]),
        "vcfc_decompress_file_multi": (i, [C.POINTER(vp), i, C.c_char_p, C.c_char_p]),
        "vcfc_compress_index_file_multi": (i, [C.POINTER(vp), i, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
        "vcfc_query_file": (i, [vp, C.c_char_p, C.c_char_p, i]),
        "vcfc_sparsify_file": (i, [C.c_char_p, C.c_char_p]),
        "vcfc_sparse_query_file": (i, [vp, C.c_char_p, C.c_char_p, i]),
        "vcfc_create_binned_index_file": (i, [vp, C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
        "vcfc_query_binned_index_file": (i, [vp, C.c_char_p, C.c_char_p, i]),
        "vcfc_set_timing": (i, [vp, i]),
        "vcfc_last_kernel_ms": (C.c_float, [vp, i]),
        "vcfc_launch_count": (u64, [vp]),
        "vcfc_last_path": (i, [vp]),
        "vcfc_last_reject_reason": (i, [vp]),
        "vcfc_force_generic": (i, [vp, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def sparsify(vcfc_path: str, sparse_path: str) -> int:
    """.vcfc -> holey file addressed by position (sparsify_file, sparse.cpp:290-580); host only."""
    return lib().vcfc_sparsify_file(vcfc_path.encode(), sparse_path.encode())


def strerror(code: int) -> str:
    return lib().vcfc_strerror(code).decode()


def parse_headers(data: bytes):
    """(rc, header_len, sample_count) -- decompress2_metadata_headers_fd, compress.cpp:1108-1211."""
    hl, sc = C.c_size_t(0), C.c_uint64(0)
    rc = lib().vcfc_parse_headers(data, len(data), C.byref(hl), C.byref(sc))
    return rc, hl.value, sc.value


def _ptr(b):
    """bytes / bytearray / numpy array / int -> void* (no copy)."""
    if isinstance(b, int):
        return C.c_void_p(b)
    if isinstance(b, bytes):
        return C.cast(C.c_char_p(b), C.c_void_p)
    if isinstance(b, bytearray):
        return C.cast((C.c_char * len(b)).from_buffer(b), C.c_void_p)
    return C.c_void_p(b.ctypes.data)     # numpy


def _stream(handle):
    """cudaStream_t for the ABI: None -> the context's own stream (NULL); 0 -> the legacy default stream, which is what
    torch.cuda.current_stream().cuda_stream reports for torch's default stream (cudaStreamLegacy = 1 names it explicitly,
    because a NULL stream argument means "the context's stream" in this ABI); anything else is passed through."""
    if handle is None:
        return C.c_void_p(None)
    return C.c_void_p(1 if handle == 0 else handle)


class Codec:
    """One context = one (process, GPU)."""

    def __init__(self, device: int = 0):
        self._ctx = C.c_void_p(None)
        rc = lib().vcfc_gpu_init(device, C.byref(self._ctx))
        if rc != OK:
            raise VcfcError(rc, f"vcfc_gpu_init(device={device}) failed: {strerror(rc)} -- this library has no CPU path")
        self.device = device

    def close(self):
        if self._ctx:
            lib().vcfc_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- introspection ----
    @property
    def last_path(self) -> int:
        return lib().vcfc_last_path(self._ctx)

    @property
    def launches(self) -> int:
        return lib().vcfc_launch_count(self._ctx)

    @property
    def last_reject_reason(self) -> int:
        return lib().vcfc_last_reject_reason(self._ctx)

    def force_generic(self, on):
        """0/False automatic, 1/True generic kernels only, 2 tile kernels with the span-walking decoder."""
        lib().vcfc_force_generic(self._ctx, int(on))

    def set_timing(self, on: bool):
        lib().vcfc_set_timing(self._ctx, int(on))

    def last_kernel_ms(self, which: int) -> float:
        return lib().vcfc_last_kernel_ms(self._ctx, which)

    def cuda_error(self) -> str:
        return lib().vcfc_last_cuda_error(self._ctx).decode()

    # ---- host-pointer block codecs ----
    def compress_block(self, data, want_offsets: bool = False, out_cap: int | None = None, n_bytes: int | None = None):
        """Data lines -> .vcfc bytes.  Returns (rc, out, n_lines, err_line[, offsets])."""
        n = len(data) if n_bytes is None else n_bytes
        cap = lib().vcfc_encode_bound(n) if out_cap is None else out_cap
        out = C.create_string_buffer(max(cap, 1))
        olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        lcap = n // 18 + 2 if want_offsets else 0
        offs = (C.c_uint64 * lcap)() if want_offsets else None
        rc = lib().vcfc_encode_block(self._ctx, _ptr(data), n, out, cap, C.byref(olen),
                                     offs, lcap, C.byref(nl), C.byref(el))
        if rc == E_CUDA:
            raise VcfcError(rc, self.cuda_error())
        res = (rc, out.raw[:olen.value], nl.value, el.value)
        if want_offsets:
            res += (list(offs[:min(nl.value, lcap)]),)
        return res

    def decompress_block(self, data, sample_count: int, out_cap: int | None = None):
        """.vcfc data lines -> text.  Returns (rc, out, n_lines, err_line)."""
        cap = out_cap if out_cap is not None else len(data) * 520 + 4096   # a 0x7f token expands to 508 bytes
        out = C.create_string_buffer(max(cap, 1))
        olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        rc = lib().vcfc_decode_block(self._ctx, _ptr(data), len(data), sample_count, out, cap,
                                     C.byref(olen), C.byref(nl), C.byref(el))
        if rc == E_CUDA:
            raise VcfcError(rc, self.cuda_error())
        return rc, out.raw[:olen.value], nl.value, el.value

    # ---- raw pointer forms (pinned host / device memory owned by the caller) ----
    def encode_host_ptr(self, in_ptr: int, in_len: int, out_ptr: int, out_cap: int):
        olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        rc = lib().vcfc_encode_block(self._ctx, C.c_void_p(in_ptr), in_len, C.c_void_p(out_ptr), out_cap,
                                     C.byref(olen), None, 0, C.byref(nl), C.byref(el))
        if rc == E_CUDA:
            raise VcfcError(rc, self.cuda_error())
        return rc, olen.value, nl.value, el.value

    def decode_host_ptr(self, in_ptr: int, in_len: int, sample_count: int, out_ptr: int, out_cap: int):
        olen, nl, el = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
        rc = lib().vcfc_decode_block(self._ctx, C.c_void_p(in_ptr), in_len, sample_count, C.c_void_p(out_ptr), out_cap,
                                     C.byref(olen), C.byref(nl), C.byref(el))
        if rc == E_CUDA:
            raise VcfcError(rc, self.cuda_error())
        return rc, olen.value, nl.value, el.value

    def encode_dev(self, d_in: int, in_len: int, d_out: int, out_cap: int, d_result: int, stream: int = 0,
                   d_offsets: int = 0, line_cap: int = 0):
        rc = lib().vcfc_encode_block_dev(self._ctx, C.c_void_p(d_in), in_len, C.c_void_p(d_out), out_cap,
                                         C.c_void_p(d_offsets or None), line_cap, C.c_void_p(d_result),
                                         _stream(stream))
        if rc != OK:
            raise VcfcError(rc, self.cuda_error() if rc == E_CUDA else strerror(rc))

    def decode_dev(self, d_in: int, in_len: int, sample_count: int, d_out: int, out_cap: int, d_result: int,
                   stream: int = 0):
        rc = lib().vcfc_decode_block_dev(self._ctx, C.c_void_p(d_in), in_len, sample_count, C.c_void_p(d_out), out_cap,
                                         C.c_void_p(d_result), _stream(stream))
        if rc != OK:
            raise VcfcError(rc, self.cuda_error() if rc == E_CUDA else strerror(rc))

    def decode_size_dev(self, d_in: int, in_len: int, sample_count: int, stream: int = 0) -> Result:
        r = Result()
        rc = lib().vcfc_decode_size_dev(self._ctx, C.c_void_p(d_in), in_len, sample_count, C.byref(r),
                                        _stream(stream))
        if rc != OK:
            raise VcfcError(rc, self.cuda_error() if rc == E_CUDA else strerror(rc))
        return r

    def fetch_result(self, d_result: int, stream: int = 0) -> Result:
        r = Result()
        rc = lib().vcfc_fetch_result(self._ctx, C.c_void_p(d_result), C.byref(r), _stream(stream))
        if rc != OK:
            raise VcfcError(rc, self.cuda_error())
        return r

    # ---- file drivers (reference verbs) ----
    def compress(self, in_path: str, out_path: str) -> int:
        return lib().vcfc_compress_file(self._ctx, in_path.encode(), out_path.encode())

    def decompress2_fd(self, in_path: str, out_path: str) -> int:
        return lib().vcfc_decompress_file(self._ctx, in_path.encode(), out_path.encode())

    @staticmethod
    def compress_multi(codecs, in_path: str, out_path: str) -> int:
        """compress() over several contexts (one worker thread per context, host concatenation by chunk offsets)."""
        arr = (C.c_void_p * len(codecs))(*[c._ctx for c in codecs])
        return lib().vcfc_compress_file_multi(arr, len(codecs), in_path.encode(), out_path.encode())

    @staticmethod
    def compress_index_multi(codecs, in_path: str, out_path: str, index_path: str, entries_per_bin: int):
        """compress() + create_binned_index4() in one pass -> (rc, number of index entries)."""
        arr = (C.c_void_p * len(codecs))(*[c._ctx for c in codecs])
        n = C.c_uint64(0)
        rc = lib().vcfc_compress_index_file_multi(arr, len(codecs), in_path.encode(), out_path.encode(), index_path.encode(),
                                                  entries_per_bin, C.byref(n))
        return rc, n.value

    @staticmethod
    def decompress_multi(codecs, in_path: str, out_path: str) -> int:
        arr = (C.c_void_p * len(codecs))(*[c._ctx for c in codecs])
        return lib().vcfc_decompress_file_multi(arr, len(codecs), in_path.encode(), out_path.encode())

    def query(self, in_path: str, region: str, out_fd: int) -> int:
        return lib().vcfc_query_file(self._ctx, in_path.encode(), region.encode(), out_fd)

    def sparse_query(self, sparse_path: str, region: str, out_fd: int) -> int:
        """REF:START-END on a sparsified file (query_sparse_file_fd, main.cpp:235-582)."""
        return lib().vcfc_sparse_query_file(self._ctx, sparse_path.encode(), region.encode(), out_fd)

    def query_binned_index(self, vcfc_path: str, region: str, out_fd: int) -> int:
        """Indexed range query (query_binned_index_binarysearch, main.cpp:2974-3350); the index is vcfc_path + '.vcfci'."""
        return lib().vcfc_query_binned_index_file(self._ctx, vcfc_path.encode(), region.encode(), out_fd)

    def create_binned_index(self, vcfc_path: str, index_path: str, entries_per_bin: int):
        """-> (rc, number of entries); writes the .vcfci file (create_binned_index4, main.cpp:1284-1637)."""
        n = C.c_uint64(0)
        rc = lib().vcfc_create_binned_index_file(self._ctx, vcfc_path.encode(), index_path.encode(), entries_per_bin, C.byref(n))
        return rc, n.value

    # ---- whole-file semantics on bytes (tests) ----
    def compress_vcf(self, vcf: bytes):
        """compress() on an in-memory file: '#' lines pass through (+'\\n'), data regions are encoded."""
        out = bytearray()
        pos, n = 0, len(vcf)
        while pos < n:
            if vcf[pos:pos + 1] == b"\n":
                pos += 1
                continue
            if vcf[pos:pos + 1] == b"#":
                e = vcf.find(b"\n", pos)
                e = n if e < 0 else e
                out += vcf[pos:e] + b"\n"
                pos = e + 1
                continue
            h = vcf.find(b"\n#", pos)
            end = n if h < 0 else h + 1
            rc, enc, _, _ = self.compress_block(vcf[pos:end])
            out += enc
            if rc != OK:
                return rc, bytes(out)
            pos = end
        return OK, bytes(out)

    def decompress_vcfc(self, vcfc: bytes):
        rc, hl, sc = parse_headers(vcfc)
        if rc != OK:
            return rc, b""
        rc, txt, _, _ = self.decompress_block(vcfc[hl:], sc)
        return rc, vcfc[:hl] + txt
