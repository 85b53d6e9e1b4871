#!/bin/bash
# SAO kernels: parity, per-kind timings at 8 and 10 bits, the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sao.py -q -x > gpurun_out/r02u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02u_pytest.log; tail -6 gpurun_out/r02u_pytest.log
timeout 300 python tools/sao_time.py --iters 5 2>&1 | tail -1 | tee gpurun_out/r02u_sao_kinds.txt
timeout 300 python tools/sao_time.py --iters 5 --n 128 --depth 10 2>&1 | tail -1 | tee -a gpurun_out/r02u_sao_kinds.txt
timeout 600 python bench.py --workload 4k_sao256 > gpurun_out/r02u_bench_4k_sao256.json 2> gpurun_out/r02u_bench_sao.err; echo "bench sao exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02u_bench_4k_sao256.json')); print(d['roofline']['frac'], d['deblock_alone'], d['sao_alone'], d['deblock_plus_sao'])"
