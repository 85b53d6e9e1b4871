#!/bin/bash
# cp.async variants after a change: parity, then timings of the widths off the 32 grid
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nv12.py -q -x > gpurun_out/r02y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02y_pytest.log; tail -4 gpurun_out/r02y_pytest.log
for wh in "1080 1920 256" "1080 1920 1024" "1008 1000 512" "1000 1000 512" "360 288 2048" "1920 1080 256"; do
	set -- $wh
	timeout 300 python tools/quick_time.py --w $1 --h $2 --n $3 --iters 10 2>&1 | tail -1
done | tee gpurun_out/r02y_cp_async_timings.txt
