#!/bin/bash
# SAO shape variants: parity of the default build, then per-kind timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sao.py -q 2>&1 | tail -3
for so in build_variants/libhdb_*.so; do
	HDB_LIBRARY=$PWD/$so timeout 300 python tools/sao_time.py --iters 5 2>&1 | tail -1
done | tee gpurun_out/r02w_sao_variants.txt
