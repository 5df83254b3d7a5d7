"""Builds libvcfc_gpu.so (sm_100a) and the `vcfc` CLI in-tree with nvcc / g++.

    python vcf-compression_b200/build.py [--force] [-v]

nvcc cross-compiles without a GPU.  Outputs stay inside the package directory so that they
travel to the GPU box with the repo snapshot (they are git-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libvcfc_gpu.so")
CLI = os.path.join(HERE, "vcfc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CU = ["vcfc_api.cu", "vcfc_files.cu", "vcfc_generic.cu", "vcfc_encode_fast.cu", "vcfc_decode_fast.cu", "vcfc_index.cu", "vcfc_pipeline.cu", "vcfc_sparse.cu"]
FLAGS = (["-DVCFC_DEBUG"] if os.environ.get("VCFC_DEBUG") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in CU]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "vcfc_gpu.h"))
    if force or _stale(LIB, deps):
        objs = []
        procs = []
        os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
        for s in srcs:
            o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
            objs.append(o)
            if force or _stale(o, deps):
                cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
                procs.append((cmd, subprocess.Popen(cmd)))
        for cmd, p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xcompiler", "-pthread", "-lpthread"])
    main = os.path.join(HERE, "host", "vcfc_main.cpp")
    if os.path.exists(main) and (force or _stale(CLI, [main, LIB])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), main,
                               "-o", CLI, "-L", HERE, "-lvcfc_gpu", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
