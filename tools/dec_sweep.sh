#!/bin/bash
# Tuning aid (GPU box): tools/dec_sweep.sh LINES V1 V2 ... -- bench.py (encode + decode, parity gate included) for each variant lib
lines=$1; shift
for v in "$@"; do
  lib=$PWD/vcf-compression_b200/variants/libvcfc_gpu_$v.so
  [ "$v" = default ] && lib=$PWD/vcf-compression_b200/libvcfc_gpu.so
  for kind in kg random; do
    VCFC_LIB_PATH=$lib python bench.py --lines $lines --kind $kind --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/dvar_${v}_$kind.json 2> gpurun_out/dvar_${v}_$kind.err || echo "fail $v $kind"
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dvar_${v}_$kind.json")); e=d["decode"]; print("$v $kind enc %.0f GB/s k %.3f ms frac %.3f | dec %.0f GB/s kernel_ms %.3f step_ms %.3f frac %.3f" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], e["value"], e["roofline"]["kernel_ms"], e["ms_per_step"], e["roofline"]["frac"]))
except Exception as ex: print("$v $kind: no result", ex)
PY
  done
done
