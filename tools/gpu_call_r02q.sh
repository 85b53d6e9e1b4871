#!/bin/bash
# after the SAO rewrite and the cp.async change: the whole GPU suite, a SAO-only soak, a general soak
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest_gpu.log; tail -3 gpurun_out/r02q_pytest_gpu.log
timeout 400 python tools/soak.py 11 150 13 > gpurun_out/r02q_soak_sao.txt 2>&1; tail -2 gpurun_out/r02q_soak_sao.txt
timeout 400 python tools/soak.py 12 200 > gpurun_out/r02q_soak.txt 2>&1; tail -2 gpurun_out/r02q_soak.txt
