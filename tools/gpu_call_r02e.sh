#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -12 gpurun_out/r02e_pytest.log
timeout 600 python bench.py --workload 4k_edges512 > gpurun_out/r02e_bench_edges.json 2> gpurun_out/r02e_bench_edges.err; echo "bench edges exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_edges.json')); print(d['derive_kernel'], d['derive_plus_filter'], d['roofline']['frac'])"
timeout 600 python bench.py --workload 4k_sao256 > gpurun_out/r02e_bench_sao.json 2> gpurun_out/r02e_bench_sao.err; echo "bench sao exit $?"; tail -3 gpurun_out/r02e_bench_sao.err
cat gpurun_out/r02e_bench_sao.json | head -c 2500
timeout 120 python tools/quick_time.py --w 1080 --h 1920 --n 256 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:DeblockTiled -s 3 -c 1 -o gpurun_out/prof_r02e_ldgsts python tools/quick_time.py --w 1080 --h 1920 --n 256 --iters 2 > gpurun_out/r02e_ncu.log 2>&1
tail -2 gpurun_out/r02e_ncu.log
timeout 300 python bench.py --impl reference_gpu --steps 5 --warmup 2 > gpurun_out/r02e_bench_refgpu.json 2> gpurun_out/r02e_bench_refgpu.err; echo "refgpu exit $?"; cat gpurun_out/r02e_bench_refgpu.json | head -c 1800
timeout 300 python bench.py --impl reference_gpu --workload image2_single --steps 5 --warmup 2 > gpurun_out/r02e_bench_refgpu_image2.json 2>> gpurun_out/r02e_bench_refgpu.err; head -c 600 gpurun_out/r02e_bench_refgpu_image2.json
timeout 300 python tools/quick_time.py --n 1024 --iters 10 2>&1 | tail -1
ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "nvcuvid|nvidia-encode|nvjpeg" | head
tools/run_variants.sh 1024 > gpurun_out/r02e_variants.log 2>&1; cat gpurun_out/r02e_variants.log
