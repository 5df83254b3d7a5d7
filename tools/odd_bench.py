#!/usr/bin/env python3
"""GPU box: device-resident encode / decode throughput of blocks the 4-byte sample grid does not take as they are --
(a) a config-2-shaped block with 1 % irregular lines (a few odd-width samples each), (b) an all-GT:DP:GQ block --
on the default dispatch and on the generic kernels, with the bytes of the two checked against each other and the decode against
the input (parity with the oracle is the test suite's business: tests/test_gpu_parity.py::test_odd_*, ::test_gt_dp_gq_block).

    python tools/odd_bench.py [--lines 200000]
"""
import argparse, importlib, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import vcfsynth
pkg = importlib.import_module("vcf-compression_b200")


def timed(codec, d_in, n_in, d_out, cap, d_res, stream, steps=5):
    codec.encode_dev(d_in.data_ptr(), n_in, d_out.data_ptr(), cap, d_res.data_ptr(), stream)
    r = codec.fetch_result(d_res.data_ptr(), stream)
    assert r.status == 0, r.status
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        codec.encode_dev(d_in.data_ptr(), n_in, d_out.data_ptr(), cap, d_res.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, int(r.out_len), int(r.n_lines), codec.last_path


def timed_dec(codec, d_c, n_c, samples, d_txt, d_res, stream, steps=3):
    codec.decode_dev(d_c.data_ptr(), n_c, samples, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), stream)
    r = codec.fetch_result(d_res.data_ptr(), stream)
    assert r.status == 0, r.status
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        codec.decode_dev(d_c.data_ptr(), n_c, samples, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, int(r.out_len), codec.last_path


def run(name, text: bytes, samples: int, codec, dev, with_generic=True):
    n = len(text)
    pad = (-n) % 16
    d_in = torch.frombuffer(bytearray(text + b"\0" * pad), dtype=torch.uint8).to(dev)
    cap = int(n * 1.6) + (1 << 20)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_out2 = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_txt = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {"workload": name, "bytes": n}
    codec.force_generic(0)
    codec.set_timing(True)
    ms, olen, nl, path = timed(codec, d_in, n, d_out, cap, d_res, st)
    res["encode"] = {"gbs": n / ms / 1e6, "ms": ms, "kernel_ms": codec.last_kernel_ms(0), "path": path, "out_bytes": olen, "lines": nl}
    codec.set_timing(False)
    dms, tlen, dpath = timed_dec(codec, d_out, olen, samples, d_txt, d_res, st)
    res["decode"] = {"gbs": n / dms / 1e6, "ms": dms, "path": dpath, "round_trip": bool(tlen == n and torch.equal(d_txt[:n], d_in[:n]))}
    if with_generic:
        codec.force_generic(1)
        gms, golen, gnl, gpath = timed(codec, d_in, n, d_out2, cap, d_res, st, steps=2)
        res["encode_generic"] = {"gbs": n / gms / 1e6, "ms": gms, "path": gpath,
                                 "same_bytes": bool(golen == olen and torch.equal(d_out[:olen], d_out2[:olen]))}
        gdms, _, gdpath = timed_dec(codec, d_out, olen, samples, d_txt, d_res, st, steps=1)
        res["decode_generic"] = {"gbs": n / gdms / 1e6, "ms": gdms, "path": gdpath}
        codec.force_generic(0)
    print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=200000)
    ap.add_argument("--every", type=int, default=100, help="one line in this many carries odd-width samples")
    ap.add_argument("--only-mixed", action="store_true")
    ap.add_argument("--only-gtdp", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    codec = pkg.Codec(0)
    rng = random.Random(1)
    odd = [b"10|0", b"0|10", b".", b"0", b"1", b"0|1:35:99", b"./.:."]
    if not args.only_gtdp:
        mixed_part(args, codec, dev, rng, odd)
    if args.only_mixed:
        return
    gtdp_part(args, codec, dev)


def mixed_part(args, codec, dev, rng, odd):
    # (a) config-2 shape, 1 % of the lines carry a few odd-width samples
    d, lens = vcfsynth.generate("kg", args.lines, 2504, seed=20, device=dev)
    text = d.cpu().numpy().tobytes()
    starts = np.concatenate([[0], np.cumsum(lens.cpu().numpy())])
    pieces, prev = [], 0
    for li in range(0, args.lines, args.every):
        a, b = int(starts[li]), int(starts[li + 1])
        cols = text[a:b - 1].split(b"\t")
        for _ in range(5):
            cols[rng.randrange(9, len(cols))] = rng.choice(odd)
        pieces.append(text[prev:a]); pieces.append(b"\t".join(cols) + b"\n"); prev = b
    pieces.append(text[prev:])
    mixed = b"".join(pieces)
    run("regular (config 2 shape)", text, 2504, codec, dev, with_generic=False)
    os.environ["VCFC_ENC_FORCE_ODD"] = "1"          # the same block on the encoder instantiation that carries the term walkers
    codec_odd = pkg.Codec(0)
    del os.environ["VCFC_ENC_FORCE_ODD"]
    run("regular, on the encoder instantiation with the term walkers", text, 2504, codec_odd, dev, with_generic=False)
    del codec_odd
    run("config 2 shape, 1 line in %d with 5 odd-width samples" % args.every, mixed, 2504, codec, dev, with_generic=not args.only_mixed)


def gtdp_part(args, codec, dev):
    # (b) all GT:DP:GQ
    L, S = max(1000, args.lines // 4), 1000
    g = np.random.default_rng(3)
    gt = np.array([list(b"0|0"), list(b"0|1"), list(b"1|0"), list(b"1|1")], dtype=np.uint8)[g.integers(0, 4, (L, S))]
    cell = np.empty((L, S, 10), dtype=np.uint8)
    cell[:, :, 0:3] = gt; cell[:, :, 3] = ord(":")
    dp = g.integers(10, 100, (L, S)); cell[:, :, 4] = 48 + dp // 10; cell[:, :, 5] = 48 + dp % 10; cell[:, :, 6] = ord(":")
    gq = g.integers(10, 100, (L, S)); cell[:, :, 7] = 48 + gq // 10; cell[:, :, 8] = 48 + gq % 10; cell[:, :, 9] = 9
    cell[:, S - 1, 9] = 10
    rows = [b"3\t%d\t.\tC\tT\t.\tPASS\tDP=100\tGT:DP:GQ\t" % (1000 + i) for i in range(L)]
    body = cell.reshape(L, S * 10)
    gtdp = b"".join(r + body[i].tobytes() for i, r in enumerate(rows))
    run("all GT:DP:GQ (1000 samples per line)", gtdp, S, codec, dev, with_generic=not args.only_gtdp)
    del gtdp, body, cell
    # (c) chrX outside the pseudo-autosomal regions: the male columns (a fixed half of the samples) are haploid calls
    L, S = max(1000, args.lines // 2), 2504
    male = g.random(S) < 0.5
    width = np.where(male, 2, 4)
    start = np.concatenate([[0], np.cumsum(width)])[:-1]                  # the layout is the same for every line
    af = (1.0 / (2 * S)) * (0.5 * 2 * S) ** g.random(L)
    body = np.empty((L, int(width.sum())), dtype=np.uint8)
    body[:, start] = (g.random((L, S)) < af[:, None]).astype(np.uint8) + 48
    fem = start[~male]
    body[:, fem + 1] = ord("|")
    body[:, fem + 2] = (g.random((L, len(fem))) < af[:, None]).astype(np.uint8) + 48
    body[:, start + width - 1] = 9
    body[:, -1] = 10
    lines = [b"X\t%d\t.\tC\tT\t.\tPASS\tDP=100\tGT\t" % (2700000 + 37 * i) + body[i].tobytes() for i in range(L)]
    run("chrX shape: half of the 2504 sample columns haploid", b"".join(lines), S, codec, dev, with_generic=not args.only_gtdp)


main()
