#!/bin/bash
# the whole GPU suite, smoke, a SAO-only soak (8-bit and 16-bit containers), a general soak
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest_gpu.log; tail -3 gpurun_out/r02q_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02q_smoke.log
timeout 400 python tools/soak.py 31 120 13,14 > gpurun_out/r02q_soak_sao.txt 2>&1; tail -2 gpurun_out/r02q_soak_sao.txt
timeout 400 python tools/soak.py 32 150 > gpurun_out/r02q_soak.txt 2>&1; tail -2 gpurun_out/r02q_soak.txt
