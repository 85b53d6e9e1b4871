#!/bin/bash
# Development aid: time every library under build_variants/ with tools/quick_time.py (run on the GPU box).
#   tools/run_variants.sh [frames] [ENV=VALUE ...]   -- extra environment applies to every run
cd "$(dirname "$0")/.."
N=${1:-256}; shift
for so in build_variants/libhdb_*.so; do
	env "$@" HDB_LIBRARY=$PWD/$so timeout 300 python tools/quick_time.py --n $N --iters ${ITERS:-10} 2>&1 | tail -1
done
