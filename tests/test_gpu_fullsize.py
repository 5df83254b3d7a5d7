"""GPU suite (-m gpu), BASELINE.json shapes at sizes the oracle cannot sweep in seconds: parity through
size-independent properties (device round trip, oracle on line windows, offsets/headers consistency),
on data generated on the device by tests/vcfsynth.py.

config 2: 1000G-chr20-shaped, 2504 samples          config 3: biobank-shaped, 100k samples (400 KB lines)
config 4: position-range query on the 2504-sample file
"""
import importlib
import os

import pytest
import torch

import oraclelib as O
import vcfsynth

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("vcf-compression_b200")


@pytest.fixture(scope="module")
def codec():
    c = pkg.Codec(0)
    yield c
    c.close()


def _device_roundtrip(codec, kind, n_lines, n_samples, seed, windows=3, win_lines=400, decode_fast=True):
    dev = torch.device("cuda:0")
    d_in, lens = vcfsynth.generate(kind, n_lines, n_samples, seed=seed, device=dev)
    n_in = d_in.numel()
    cap = n_in // 2 + (1 << 20)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    d_offs = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    codec.encode_dev(d_in.data_ptr(), n_in, d_out.data_ptr(), cap, d_res.data_ptr(), st, d_offs.data_ptr(), n_lines + 1)
    r = codec.fetch_result(d_res.data_ptr(), st)
    assert r.status == 0 and r.n_lines == n_lines
    assert codec.last_path == pkg.PATH_FAST, codec.last_reject_reason   # these shapes must be served by the tile kernels
    n_out = int(r.out_len)
    offs = d_offs[:n_lines]
    # every line offset points at a length header whose value reaches the next offset (checksum of the structure)
    nxt = torch.cat([offs[1:], torch.tensor([n_out], device=dev)])
    hdr = d_out[offs.unsqueeze(1) + torch.arange(4, device=dev)].to(torch.int64)
    assert bool(((hdr[:, 0] >> 6) == 3).all())
    ll = ((hdr[:, 0] & 0x3F) << 24) | (hdr[:, 1] << 16) | (hdr[:, 2] << 8) | hdr[:, 3]
    assert bool((offs + 4 + ll == nxt).all())
    # oracle on windows of whole lines
    starts = torch.zeros(n_lines + 1, dtype=torch.int64, device=dev)
    starts[1:] = torch.cumsum(lens, 0)
    w = min(win_lines, n_lines)
    for lo in sorted({0, (n_lines - w) // 2, n_lines - w})[:windows]:
        a, b = int(starts[lo]), int(starts[lo + w])
        oa, ob = int(offs[lo]), (int(offs[lo + w]) if lo + w < n_lines else n_out)
        orc, oout, onl, _ = O.compress_block(bytes(d_in[a:b].cpu().numpy()))
        assert orc == 0 and onl == w
        assert oout == bytes(d_out[oa:ob].cpu().numpy())
    # decode(encode(x)) == x on the device
    sz = codec.decode_size_dev(d_out.data_ptr(), n_out, n_samples, st)
    assert sz.status == 0 and sz.out_len == n_in and sz.n_lines == n_lines
    d_txt = torch.empty(n_in + 64, dtype=torch.uint8, device=dev)
    codec.decode_dev(d_out.data_ptr(), n_out, n_samples, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), st)
    r2 = codec.fetch_result(d_res.data_ptr(), st)
    assert r2.status == 0 and r2.out_len == n_in and r2.n_lines == n_lines
    if decode_fast:
        assert codec.last_path == pkg.PATH_FAST
    assert torch.equal(d_txt[:n_in], d_in)
    # encode is deterministic: a second pass writes the same bytes
    d_out2 = torch.empty(n_out + 64, dtype=torch.uint8, device=dev)
    codec.encode_dev(d_in.data_ptr(), n_in, d_out2.data_ptr(), n_out + 64, d_res.data_ptr(), st)
    r3 = codec.fetch_result(d_res.data_ptr(), st)
    assert r3.status == 0 and r3.out_len == n_out and torch.equal(d_out2[:n_out], d_out[:n_out])
    return n_in, n_out


def test_config2_shape_2504_samples(codec):
    n_in, n_out = _device_roundtrip(codec, "kg", 200_000, 2504, seed=20)
    assert 10 < n_in / n_out < 30


def test_config1_distribution_dense(codec):
    n_in, n_out = _device_roundtrip(codec, "random", 60_000, 2504, seed=5)
    assert 6 < n_in / n_out < 10             # the reference's 7.84x on random_vcf.py data


def test_config3_shape_100k_samples(codec):
    _device_roundtrip(codec, "kg", 1500, 100_000, seed=3, win_lines=5)


def test_config4_range_query(codec, tmp_path):
    """query REF:START-END prints exactly the data lines whose POS is in range (main.cpp:3777-3929)."""
    d_in, lens = vcfsynth.generate("kg", 20_000, 2504, seed=9, device="cuda:0")
    data = bytes(d_in.cpu().numpy())
    vcf = vcfsynth.header(2504) + data
    ip, op = str(tmp_path / "a.vcf"), str(tmp_path / "a.vcfc")
    open(ip, "wb").write(vcf)
    assert codec.compress(ip, op) == 0
    lines = data.split(b"\n")[:-1]
    pos = [int(l.split(b"\t", 2)[1]) for l in lines]
    lo, hi = pos[5000], pos[5000] + 140_000           # ~140 kb window as in compare-query.sh:10-11
    want = b"".join(l + b"\n" for l, p in zip(lines, pos) if lo <= p <= hi)
    outp = str(tmp_path / "q.out")
    fd = os.open(outp, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
    assert codec.query(op, f"20:{lo}-{hi}", fd) == 0
    os.close(fd)
    assert open(outp, "rb").read() == want and len(want) > 0
    fd = os.open(outp, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
    assert codec.query(op, "21", fd) == 0             # other chromosome: nothing
    os.close(fd)
    assert open(outp, "rb").read() == b""
    rp = str(tmp_path / "a.rt")
    assert codec.decompress2_fd(op, rp) == 0
    assert open(rp, "rb").read() == vcf


def test_config1_exact_file(codec, tmp_path):
    """BASELINE.json configs[0] / SURVEY.md 7 step 4: the reference generator's exact 10k x 2504 file (tests/config1gen.py,
    input sha256 c7c9e4e3...) through the file drivers: the compressed file has the sha256 the unmodified reference binary
    produced (580246c9...), the round trip is the input, `query 1:12000-14000` prints the 1001 lines of the
    window, and the index / indexed query agree with the oracle."""
    import hashlib
    import config1gen
    vcf = config1gen.generate()
    assert hashlib.sha256(vcf).hexdigest() == config1gen.INPUT_SHA256
    ip, op, rp = (str(tmp_path / x) for x in ("c1.vcf", "c1.vcfc", "c1.rt"))
    open(ip, "wb").write(vcf)
    rc, n_ent = pkg.Codec.compress_index_multi([codec], ip, op, op + ".vcfci", 150)
    assert rc == 0
    vcfc = open(op, "rb").read()
    assert len(vcfc) == config1gen.VCFC_LEN and hashlib.sha256(vcfc).hexdigest() == config1gen.VCFC_SHA256
    assert codec.decompress2_fd(op, rp) == 0
    assert hashlib.sha256(open(rp, "rb").read()).hexdigest() == config1gen.INPUT_SHA256
    lines = vcf[vcf.index(b"\n1\t") + 1:].split(b"\n")[:-1]
    want = b"".join(l + b"\n" for l in lines if 12000 <= int(l.split(b"\t", 2)[1]) <= 14000)
    assert want.count(b"\n") == 1001
    for verb in (codec.query, codec.query_binned_index):
        outp = str(tmp_path / "q.out")
        fd = os.open(outp, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
        assert verb(op, "1:12000-14000", fd) == 0
        os.close(fd)
        assert open(outp, "rb").read() == want
    on, oidx = O.build_binned_index(vcfc, 150)
    assert n_ent == on and open(op + ".vcfci", "rb").read() == oidx
