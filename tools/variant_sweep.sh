for v in A B C E F; do
  for kind in kg random; do
    VCFC_LIB_PATH=$PWD/vcf-compression_b200/variants/libvcfc_gpu_$v.so python bench.py --lines 400000 --kind $kind --steps 10 --warmup 3 --no-e2e --no-cpu --no-decode > gpurun_out/var_${v}_$kind.json 2> gpurun_out/var_${v}_$kind.err || echo "fail $v $kind"
  done
done
