#!/bin/bash
# Development aid (GPU box): dynamic instruction mix and pipe utilisation of DeblockTiledKernel per library variant.
cd "$(dirname "$0")/.."
for tag in "$@"; do
	so=$PWD/build_variants/libhdb_$tag.so
	HDB_LIBRARY=$so python tools/quick_time.py --n 64 --iters 3 > gpurun_out/mix_$tag.plain.log 2>&1 || { echo "plain run failed for $tag"; continue; }
	HDB_LIBRARY=$so ncu --section InstructionStats --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats \
		--clock-control none -k regex:DeblockTiled -s 3 -c 1 -f -o gpurun_out/mix_$tag python tools/quick_time.py --n 64 --iters 3 > gpurun_out/mix_$tag.log 2>&1
	echo "$tag done rc=$?"
done
