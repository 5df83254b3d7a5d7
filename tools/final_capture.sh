#!/bin/bash
# GPU box: the round's closing measurements into gpurun_out/ (tag = $1): bench lines (kg full size, random), the ncu launch
# list and one --set full capture of the four largest kernels on a 400k-line run, the odd-width blocks.
tag=${1:-r2f}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1 || echo "smoke failed"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || echo "bench failed"
python bench.py --kind random --no-cpu > gpurun_out/${tag}_bench_random.json 2> gpurun_out/${tag}_bench_random.err || echo "bench random failed"
python tools/odd_bench.py > gpurun_out/${tag}_odd.jsonl 2> gpurun_out/${tag}_odd.err || echo "odd bench failed"
A="--lines 400000 --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py $A > gpurun_out/${tag}_small.json 2> gpurun_out/${tag}_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/${tag}_launches.csv python bench.py $A > gpurun_out/${tag}_launches.log 2>&1
B="--lines 400000 --steps 1 --warmup 0 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:"k_encode_stream|k_gather_tiles|k_dec_sizes|k_dec_expand_grid" -c 4 -o gpurun_out/prof_${tag} -f python bench.py $B > gpurun_out/${tag}_ncu_full.log 2>&1
cuobjdump -sass vcf-compression_b200/libvcfc_gpu.so 2>/dev/null | grep -E "UBLKCP|SYNCS|UTMA" | sort | uniq -c > gpurun_out/${tag}_sass_bulk.txt
tail -c 600 gpurun_out/${tag}_bench.json; echo; tail -2 gpurun_out/${tag}_ncu_full.log
