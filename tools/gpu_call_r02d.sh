#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
tail -8 gpurun_out/r02d_pytest.log
for sz in "1080 1920" "1008 1000" "1000 1000" "360 288"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02d_ldgsts.log 2>&1
done
cat gpurun_out/r02d_ldgsts.log
timeout 600 python bench.py --workload 4k_yuv420p10_512 > gpurun_out/r02d_bench_p10.json 2> gpurun_out/r02d_bench_p10.err; echo "bench p10 exit $?"; tail -3 gpurun_out/r02d_bench_p10.err
head -c 1500 gpurun_out/r02d_bench_p10.json
