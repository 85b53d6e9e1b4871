#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
for sz in "1080 1920" "360 288" "1000 1000" "1008 1000" "1920 1080"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02b_ldgsts.log 2>&1
done
cat gpurun_out/r02b_ldgsts.log
