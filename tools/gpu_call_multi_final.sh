#!/bin/bash
# usage: gpu_call_multi_final.sh <N>  -- sanity of the multi-GPU legs on the final library (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02n_pytest_multi_n$N.log 2>&1; echo "pytest multi exit $?" >> gpurun_out/r02n_pytest_multi_n$N.log
tail -4 gpurun_out/r02n_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/r02n_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02n_bench_reference_n$N.json 2> gpurun_out/r02n_bench_reference_n$N.err; echo "reference arm N=$N exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/r02n_bench_n$N.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "link_probe")})
print(json.dumps(d.get("strips_8k"), indent=0)[:1500])
r = json.load(open("gpurun_out/r02n_bench_reference_n$N.json"))
print({k: r.get(k) for k in ("impl", "value", "unit", "cpu_baseline")})
PY
