#!/bin/bash
# round 2, first GPU call: the whole GPU suite, the default bench line, the strip stream on logical shards,
# kernel-shape variants and the cp.async loader against the padded-scratch kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --workload 8k_strips --gpus 8 --steps 2 > gpurun_out/r02a_strips_logical.json 2> gpurun_out/r02a_strips_logical.err; echo "strips exit $?"
tools/run_variants.sh 256 > gpurun_out/r02a_variants.log 2>&1
tools/run_variants.sh 1024 >> gpurun_out/r02a_variants.log 2>&1
for sz in "1080 1920" "360 288" "1000 1000"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02a_ldgsts.log 2>&1
  HDB_NO_LDGSTS=1 timeout 120 python tools/quick_time.py --w $1 --h $2 --n 256 >> gpurun_out/r02a_ldgsts.log 2>&1
done
timeout 120 python tools/quick_time.py --w 1920 --h 1080 --n 256 >> gpurun_out/r02a_ldgsts.log 2>&1
cat gpurun_out/r02a_variants.log gpurun_out/r02a_ldgsts.log
