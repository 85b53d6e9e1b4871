#!/bin/bash
# SAO shape variants: per-kind timings
mkdir -p gpurun_out
for so in build_variants/libhdb_*.so; do
	HDB_LIBRARY=$PWD/$so timeout 300 python tools/sao_time.py --iters 5 ${SAO_ARGS} 2>&1 | tail -1
done | tee gpurun_out/r02w_sao_variants.txt
