#!/bin/bash
# Build tuning variants of the library: tools/build_variants.sh "RB STAGES WARPS MIN_CTAS" ...
# -> build_variants/libhdb_rb<RB>_s<S>_w<W>_c<C>.so   (select with HDB_LIBRARY=...)
set -e
cd "$(dirname "$0")/../gpu_video_codec_b200/csrc"
mkdir -p ../../build_variants
for v in "$@"; do
	set -- $v
	tag="rb$1_s$2_w$3_c$4"
	d=../../build_variants/obj_$tag
	mkdir -p $d
	for f in deblock_kernels deblock_api deblock_multi; do
		nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
			-DHDB_RB=$1 -DHDB_STAGES=$2 -DHDB_WARPS=$3 -DHDB_MIN_CTAS=$4 -c $f.cu -o $d/$f.o &
	done
	wait
	nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../build_variants/libhdb_$tag.so $d/*.o -lpthread
	rm -rf $d
	echo built $tag
done
