#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -12 gpurun_out/r02f_pytest.log
timeout 600 python bench.py --workload 4k_sao256 > gpurun_out/r02f_bench_sao.json 2> gpurun_out/r02f_bench_sao.err; echo "bench sao exit $?"; tail -2 gpurun_out/r02f_bench_sao.err; head -c 2600 gpurun_out/r02f_bench_sao.json; echo
timeout 600 python bench.py --workload 4k_edges512 > gpurun_out/r02f_bench_edges.json 2> gpurun_out/r02f_bench_edges.err; echo "bench edges exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_edges.json')); print(d['derive_kernel'], d['derive_plus_filter'], d['roofline']['frac'])"
for sz in "1080 1920" "1008 1000" "360 288"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02f_ldgsts.log 2>&1
done
timeout 120 python tools/quick_time.py --w 1080 --h 1920 --n 1024 >> gpurun_out/r02f_ldgsts.log 2>&1
cat gpurun_out/r02f_ldgsts.log
HDB_REFGPU_LOG=gpurun_out/r02f_refgpu_4k.log timeout 300 python bench.py --impl reference_gpu --steps 5 --warmup 2 > gpurun_out/r02f_bench_refgpu.json 2> gpurun_out/r02f_bench_refgpu.err; echo "refgpu exit $?"; head -c 1500 gpurun_out/r02f_bench_refgpu.json; echo; head -c 1500 gpurun_out/r02f_refgpu_4k.log
HDB_REFGPU_LOG=gpurun_out/r02f_refgpu_1080.log timeout 300 python bench.py --impl reference_gpu --workload 1080p_batch256 --steps 5 --warmup 2 > gpurun_out/r02f_bench_refgpu_1080.json 2>> gpurun_out/r02f_bench_refgpu.err; head -c 800 gpurun_out/r02f_bench_refgpu_1080.json; echo
