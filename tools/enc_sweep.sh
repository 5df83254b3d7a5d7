#!/bin/bash
# Tuning aid (GPU box): tools/enc_sweep.sh LINES V1 V2 ... -- bench.py (encode only, parity gate included) for each variant lib, kg and random
lines=$1; shift
for v in "$@"; do
  lib=$PWD/vcf-compression_b200/variants/libvcfc_gpu_$v.so
  [ "$v" = default ] && lib=$PWD/vcf-compression_b200/libvcfc_gpu.so
  for kind in kg random; do
    VCFC_LIB_PATH=$lib python bench.py --lines $lines --kind $kind --steps 10 --warmup 3 --no-e2e --no-cpu --no-decode > gpurun_out/var_${v}_$kind.json 2> gpurun_out/var_${v}_$kind.err || echo "fail $v $kind"
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${v}_$kind.json")); print("$v $kind value %.0f GB/s  kernel_ms %.3f  step_ms %.3f frac %.3f" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["roofline"]["frac"]))
except Exception as e: print("$v $kind: no result", e)
PY
  done
done
