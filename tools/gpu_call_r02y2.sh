#!/bin/bash
# library variants under build_variants/: the main workload's kernel time (checksums must agree)
mkdir -p gpurun_out
for so in build_variants/libhdb_*.so; do
	HDB_LIBRARY=$PWD/$so timeout 300 python tools/quick_time.py --n 1024 --iters 10 2>&1 | tail -1
done | tee gpurun_out/r02y2_variants.txt
