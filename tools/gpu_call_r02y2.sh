#!/bin/bash
mkdir -p gpurun_out
for so in build_variants/libhdb_*.so; do
for wh in "1080 1920 256" "1008 1000 512" "360 288 2048"; do
	set -- $wh
	HDB_LIBRARY=$PWD/$so timeout 300 python tools/quick_time.py --w $1 --h $2 --n $3 --iters 10 2>&1 | tail -1
done; done | tee gpurun_out/r02y2_unroll.txt
