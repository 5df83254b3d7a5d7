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
