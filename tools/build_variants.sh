#!/bin/bash
# Build tuning variants of the library (development aid):
#   tools/build_variants.sh "tag -DHDB_X=.. -DHDB_Y=.." ...
# -> build_variants/libhdb_<tag>.so   (select at run time with HDB_LIBRARY=...)
set -e
cd "$(dirname "$0")/../gpu_video_codec_b200/csrc"
mkdir -p ../../build_variants
for v in "$@"; do
	set -- $v
	tag=$1; shift
	d=../../build_variants/obj_$tag
	mkdir -p $d
	(
	for f in deblock_kernels sao_kernels deblock_api deblock_multi; do
		nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v \
			"$@" -c $f.cu -o $d/$f.o > $d/$f.log 2>&1 &
	done
	wait
	nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../build_variants/libhdb_$tag.so $d/*.o -lpthread
	echo "built $tag: $(grep -A2 'DeblockTiledKernel' $d/deblock_kernels.log | grep -E 'spill|Used' | tr '\n' ' ')"
	rm -rf $d
	) &
done
wait
