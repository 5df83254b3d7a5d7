#!/bin/bash
# Tuning aid: builds libvcfc_gpu.so with extra -D flags into vcf-compression_b200/variants/libvcfc_gpu_<name>.so
#   tools/build_variant.sh NAME "-DVCFC_ENC_SSTAGE=4096 ..."      then: VCFC_LIB_PATH=.../variants/libvcfc_gpu_NAME.so python bench.py
set -e
name=$1; shift; defs="$*"
root=$(cd "$(dirname "$0")/.." && pwd); src=$root/vcf-compression_b200/csrc; out=$root/vcf-compression_b200/variants; mkdir -p $out/obj_$name
for f in vcfc_api vcfc_files vcfc_generic vcfc_encode_fast vcfc_decode_fast vcfc_index vcfc_pipeline vcfc_sparse; do
  nvcc $defs -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wno-unused-function --expt-relaxed-constexpr -c $src/$f.cu -o $out/obj_$name/$f.o 2>/dev/null &
done
wait
nvcc -shared -o $out/libvcfc_gpu_$name.so $out/obj_$name/*.o -cudart static -Xcompiler -pthread -lpthread 2>/dev/null
rm -rf $out/obj_$name
echo $out/libvcfc_gpu_$name.so
