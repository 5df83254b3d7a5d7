#!/usr/bin/env python3
"""Debug aid (GPU box): encode seeded blocks on the CUDA path and on the oracle, print where they first differ.
    python tests/encdiff.py
"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oraclelib as O
import vcfgen
pkg = importlib.import_module("vcf-compression_b200")


def first_diff(data, codec, name):
    orc, oout, onl, oel, ooffs = O.compress_block(data, want_offsets=True)
    rc, out, nl, el, offs = codec.compress_block(data, want_offsets=True)
    if rc == 0 and out == oout:
        print(f"{name}: ok ({len(data)} -> {len(out)} bytes, path {codec.last_path}, reject {codec.last_reject_reason})")
        return True
    print(f"{name}: MISMATCH rc {rc} vs {orc}, len {len(out)} vs {len(oout)}, path {codec.last_path}, reject {codec.last_reject_reason}")
    lines = data.split(b"\n")
    starts = [0]
    for l in lines[:-1]:
        starts.append(starts[-1] + len(l) + 1)
    n = min(len(out), len(oout))
    i = next((k for k in range(n) if out[k] != oout[k]), n)
    # which line
    li = max(k for k, o in enumerate(ooffs) if o <= i) if ooffs else 0
    print(f"  first diff at out byte {i} (line {li}, +{i - ooffs[li]} into it); ours {out[i:i+8].hex()} oracle {oout[i:i+8].hex()}")
    line = lines[li]
    s0 = starts[li]
    cols = line.split(b"\t")
    req = sum(len(c) + 1 for c in cols[:9])
    print(f"  line starts at in byte {s0} (mod 16 = {s0 % 16}), first sample at {s0 + req} (mod 4 = {(s0 + req) % 4}, mod 16 = {(s0 + req) % 16}, mod 32768 = {(s0+req) % 32768}), {len(cols) - 9} samples, line len {len(line)}")
    # decode oracle tokens of that line up to the diff to find the sample index
    o = ooffs[li] + 8 + req
    k = 0
    while o < i:
        b = oout[o]
        if b & 0x80 == 0: k += b; o += 1
        elif (b & 0xE0) == 0xE0:
            o += 1
            while oout[o] not in (9, 10): o += 1
            o += 1; k += 1
        else: k += b & 0x1F; o += 1
    print(f"  diff token begins at sample {k}: in byte {s0 + req + 4 * k} (window mod 4096 = {(4 * k + ((s0 + req) & 15)) % 4096}); samples around: {cols[9 + max(0, k - 3):9 + k + 4]}")
    return False


def main():
    codec = pkg.Codec(0)
    ok = True
    cases = [("kg300", vcfgen.kg_like, dict(n_lines=300, n_samples=2504, seed=17)),
             ("rnd64", vcfgen.random_vcf_like, dict(n_lines=64, n_samples=2504, seed=12)),
             ("rnd333", vcfgen.random_vcf_like, dict(n_lines=200, n_samples=333, seed=11)),
             ("dense", vcfgen.random_vcf_like, dict(n_lines=40, n_samples=2504, seed=13, probs=(0.5, 0.4, 0.1))),
             ("wide", vcfgen.random_vcf_like, dict(n_lines=3, n_samples=100000, seed=16, probs=(0.98, 0.02, 0.0))),
             ("s7", vcfgen.random_vcf_like, dict(n_lines=3000, n_samples=7, seed=14)),
             ("kg1000", vcfgen.kg_like, dict(n_lines=100, n_samples=1000, seed=18))]
    import goldenlib
    for name, g in goldenlib.load_all().items():
        if g["entry"]["compress_rc"] != 0:
            continue
        vcf = g["vcf"]
        # data-line region as compress_vcf sees it: skip '#' lines
        pos = 0
        while pos < len(vcf) and vcf[pos:pos + 1] == b"#":
            pos = vcf.index(b"\n", pos) + 1
        ok &= first_diff(vcf[pos:], codec, "golden:" + name)
    for name, mk, args in cases:
        _, data = mk(**args)
        ok &= first_diff(data, codec, name)
    sys.exit(0 if ok else 1)


main()
