"""CPU suite: the C-ABI library builds, loads and exports every symbol include/vcfc_gpu.h declares.
No compute call is made here (there is no GPU in the build container and no CPU path in the library)."""
import ctypes
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("vcf-compression_b200")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(pkg.LIB_PATH):
        build = importlib.import_module("vcf-compression_b200.build")
        build.build()
    return pkg.LIB_PATH


def declared_symbols():
    src = open(pkg.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcfc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_block_api():
    syms = declared_symbols()
    for s in ("vcfc_gpu_init", "vcfc_encode_block", "vcfc_encode_block_dev", "vcfc_decode_block",
              "vcfc_decode_block_dev", "vcfc_compress_file", "vcfc_decompress_file", "vcfc_query_file",
              "vcfc_parse_headers", "vcfc_encode_bound"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_binding_loads_and_pure_host_calls_work(built):
    L = pkg.lib()
    assert L.vcfc_strerror(0) == b"ok"
    assert pkg.strerror(pkg.E_TOOFEW)
    assert L.vcfc_encode_bound(1000) >= 1000 + 8
    hdr = b"##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\n"
    assert pkg.parse_headers(hdr + b"x") == (0, len(hdr), 3)
    assert pkg.parse_headers(hdr)[0] == pkg.E_HEADER          # no data lines: the reference throws
    assert pkg.parse_headers(b"#CHROM\tPOS\nx")[0] == pkg.E_HEADER
    assert pkg.parse_headers(b"1\t2\n")[0] == pkg.E_HEADER


def test_no_cpu_fallback_without_a_device(built):
    """Without a usable sm_100 device the context cannot be created and nothing computes."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.VcfcError):
        pkg.Codec(0)


def test_product_does_not_link_the_oracle(built):
    out = subprocess.run(["ldd", built], capture_output=True, text=True).stdout
    assert "oracle" not in out
    srcs = os.path.join(ROOT, "vcf-compression_b200")
    for dp, _, fs in os.walk(srcs):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "vcfc_oracle" not in txt and "oraclelib" not in txt, f


def test_cli_exists_and_reports_missing_gpu(built):
    if not os.path.exists(pkg.CLI_PATH):
        pytest.skip("CLI not built")
    p = subprocess.run([pkg.CLI_PATH, "frobnicate"], capture_output=True, text=True)
    assert "Unknown action name: frobnicate" in p.stdout


def test_index_column_parsing_host_build_matches_oracle(tmp_path):
    """vcfc_index.cuh (the END / chromosome-index code of k_index_lines and of the indexed query) compiled for the host
    gives the oracle's answer on tricky columns: signs, blanks, overflow, empty terms, duplicate keys, malformed pairs."""
    import ctypes as C
    import itertools
    import subprocess
    import oraclelib as O
    so = str(tmp_path / "index_host_shim.so")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "index_host_shim.cu")
    # host build with g++: the header only needs the CUDA function-space qualifiers to vanish
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-D__host__=", "-D__device__=", "-o", so, src])
    shim = C.CDLL(so)
    shim.shim_line_index_fields.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_longlong), C.POINTER(C.c_uint8)]
    orc = O.lib()
    orc.vcfc_oracle_line_index_fields.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_long), C.POINTER(C.c_uint8)]
    chroms = ["1", "9", "10", "22", "23", "0", "X", "Y", "M", "MT", "chr1", "", "x", "2 "]
    poss = ["100", "0", "", " 7", "\t".strip() or "5", "+12", "-3", "12a", "a", "99999999999999999999999", "0012", "1 ", "18446744073709551615"]
    refs = ["A", "ACGT", "", "ACGTACGTACGT"]
    alts = ["C", "G,T", "ACGTT,A", ",", "", "<DEL>", "<CN0>,<CN2>", "A,<DEL>", "AC<"]
    infos = ["AC=1", "END=500", "END=500,900", "END=900,500;X", "SVLEN=-40", "SVLEN=40,-70;END", "END=", "END", "SVLEN=;X=1",
             ";;AC=2;", "END=12x", "SVLEN=abc", "A=B=C", "=", "END= 77", "END=+8", "END=-5", "END=5;END=9", "END=9;END=5",
             "SVLEN=-5;END=77", "ENDX=5", "XEND=5;SVLEN=3", "END=99999999999999999999999", "SVLEN=,,4,", "k==v", "=v;END=3", ""]
    n = bad = 0
    for chrom, pos, ref, alt, info in itertools.product(chroms[:6] + chroms[6::2], poss, refs[:2] + refs[3:], alts, infos):
        line = ("%s\t%s\tid\t%s\t%s\tq\tf\t%s\tGT\t" % (chrom, pos, ref, alt, info)).encode()
        e1, r1, e2, r2 = C.c_longlong(0), C.c_uint8(0), C.c_long(0), C.c_uint8(0)
        rc1 = shim.shim_line_index_fields(line, len(line), C.byref(e1), C.byref(r1))
        rc2 = orc.vcfc_oracle_line_index_fields(line, len(line), C.byref(e2), C.byref(r2))
        n += 1
        if (rc1 != 0) != (rc2 != 0) or (rc1 == 0 and (e1.value != e2.value or r1.value != r2.value)):
            bad += 1
            assert False, (line, rc1, e1.value, r1.value, rc2, e2.value, r2.value)
    assert n > 20000 and bad == 0
