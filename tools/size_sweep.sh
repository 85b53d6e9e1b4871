#!/bin/bash
# Development aid (GPU box): kernel-only time against batch size (fixed per-launch costs show as the intercept).
cd "$(dirname "$0")/.."
for n in ${SIZES:-16 64 128 256 512}; do python tools/quick_time.py --n $n --iters 20 "$@" 2>&1 | tail -1; done
