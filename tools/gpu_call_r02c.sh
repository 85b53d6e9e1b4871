#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
for sz in "1080 1920" "1008 1000" "1000 1000"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02c_ldgsts.log 2>&1
done
cat gpurun_out/r02c_ldgsts.log
timeout 120 python tools/quick_time.py --w 1080 --h 1920 --n 256 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:DeblockTiled -s 3 -c 2 -o gpurun_out/prof_r02c_ldgsts python tools/quick_time.py --w 1080 --h 1920 --n 256 --iters 2 > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
