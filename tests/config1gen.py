"""BASELINE.json configs[0] / SURVEY.md 8(d) config 1: the file the reference's other/random_vcf.py writes with
sample_count = 2504, variant_count = 10000 (random.seed(5), Python 3.12 `random` stream) -- 100,566,566 bytes,
sha256 c7c9e4e3...; the unmodified reference binary compresses it to 12,821,634 bytes, sha256 580246c9....

This is a restatement of that generator (/root/reference/other/random_vcf.py:1-75), not a copy: the same draws in the same
order from the same Mersenne Twister stream (random.choice / random.shuffle per line, two random.random() per sample),
with the per-sample Python loop replaced by numpy on a list of the raw draws.  The sha256 of its output is checked
against the recorded one, which proves the port exact; the GPU path's compressed bytes are then checked against the
reference binary's sha256 (tests/test_gpu_fullsize.py::test_config1_exact_file)."""
import hashlib
import math
import random

import numpy as np

INPUT_SHA256 = "c7c9e4e34b025aa40f9c6526be0e3274d61c4645b7dccf7e01752feb14645918"
INPUT_LEN = 100_566_566
VCFC_SHA256 = "580246c98d81910b2ed55269d55ee91d9dccf8f67d4f078973cab92fcbad7e94"
VCFC_LEN = 12_821_634


def generate(sample_count: int = 2504, variant_count: int = 10000) -> bytes:
    rng = random.Random()          # same algorithm and seeding as the module-level functions the reference uses
    rng.seed(5)
    bases = ['A', 'T', 'G', 'C']
    probs = [0.90, 0.08, 0.02]
    s = sum(probs)
    c, cdist = 0, []
    for p in probs:
        c += p
        cdist.append(c)
    out = [b'##fileformat=VCFv4.1\n', b'##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n', b'##fileDate=20150218\n']
    digits = int(math.ceil(math.log10(sample_count)))
    hdr = ['CHROM', 'POS', 'ID', 'REF', 'ALT', 'QUAL', 'FILTER', 'INFO', 'FORMAT'] + [('HG%0' + str(digits) + 'd') % j for j in range(sample_count)]
    out.append(('#' + '\t'.join(hdr) + '\n').encode())
    pos = 10000
    rnd = rng.random
    n2 = 2 * sample_count
    digit = np.frombuffer(b"012", dtype=np.uint8)
    for i in range(variant_count):
        ref = rng.choice(bases)
        alts = [b for b in bases if b != ref]
        rng.shuffle(alts)
        alts = alts[:2]
        req = '\t'.join(['1', str(pos), 'var' + str(i), ref, ','.join(alts), '100', 'PASS', 'INFO', 'GT']) + '\t'
        pos += 2
        r = np.array([rnd() for _ in range(n2)], dtype=np.float64) * s
        a = (r >= cdist[0]).astype(np.uint8) + (r >= cdist[1]).astype(np.uint8)
        if (r >= cdist[2]).any():
            raise RuntimeError("a draw beyond the last cumulative probability (the reference would print None)")
        row = np.empty((sample_count, 4), dtype=np.uint8)
        row[:, 0] = digit[a[0::2]]
        row[:, 1] = ord('|')
        row[:, 2] = digit[a[1::2]]
        row[:, 3] = 9
        row[-1, 3] = 10
        out.append(req.encode())
        out.append(row.tobytes())
    return b"".join(out)


if __name__ == "__main__":
    import sys
    import time
    t = time.time()
    data = generate()
    print(len(data), hashlib.sha256(data).hexdigest(), f"{time.time() - t:.1f}s")
    if len(sys.argv) > 1:
        open(sys.argv[1], "wb").write(data)
