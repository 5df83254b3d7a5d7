#!/usr/bin/env python3
"""Tuning aid (GPU box): file -> file throughput of the pipeline for a few thread / chunk settings."""
import importlib, os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, vcfsynth
pkg = importlib.import_module("vcf-compression_b200")
d, lens = vcfsynth.generate("kg", 300000, 2504, seed=20, device="cuda:0")
ip, op, rp = "/dev/shm/fs_in.vcf", "/dev/shm/fs_out.vcfc", "/dev/shm/fs_rt.vcf"
with open(ip, "wb") as f:
    f.write(vcfsynth.header(2504)); f.write(d.cpu().numpy().tobytes())
n = os.path.getsize(ip)
del d
codec = pkg.Codec(0)
def best(fn, k=3):
    ts = []
    for _ in range(k):
        t = time.perf_counter(); rc = fn(); ts.append(time.perf_counter() - t); assert rc == 0
    return min(ts)
for twins, readers, writers, chunk, dchunk in [(0, 8, 3, 16, 4), (1, 8, 3, 16, 4), (1, 8, 3, 32, 4), (1, 12, 3, 16, 4), (1, 12, 3, 32, 8), (1, 16, 4, 32, 4), (0, 12, 3, 32, 4)]:
    os.environ.update(VCFC_FILE_TWINS=str(twins), VCFC_READERS=str(readers), VCFC_WRITERS=str(writers), VCFC_FILE_CHUNK_MB=str(chunk), VCFC_FILE_DCHUNK_MB=str(dchunk))
    codec.compress(ip, op)
    tc = best(lambda: codec.compress(ip, op))
    codec.decompress2_fd(op, rp)
    td = best(lambda: codec.decompress2_fd(op, rp), 2)
    print(f"twins {twins} readers {readers:2d} writers {writers:2d} chunk {chunk:3d} MB dchunk {dchunk:2d} MB: compress {n/tc/1e9:6.2f} GB/s  decompress {n/td/1e9:6.2f} GB/s", flush=True)
# plain file copy speed of this box's tmpfs (one thread, read + write), for scale
t = time.perf_counter(); subprocess.run(["cp", ip, rp]); print(f"cp on tmpfs: {n/(time.perf_counter()-t)/1e9:.2f} GB/s; nproc {os.cpu_count()}")
for x in (ip, op, rp): os.remove(x)
