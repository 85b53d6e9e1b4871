#!/bin/bash
# usage: gpu_call_multi.sh <N>  -- the multi-GPU legs on a box with N GPUs (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02m_topo_n$N.txt 2>&1
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv >> gpurun_out/r02m_topo_n$N.txt 2>&1
lscpu | head -20 >> gpurun_out/r02m_topo_n$N.txt 2>&1; numactl -H >> gpurun_out/r02m_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02m_pytest_multi_n$N.log 2>&1; echo "pytest multi exit $?" >> gpurun_out/r02m_pytest_multi_n$N.log
tail -6 gpurun_out/r02m_pytest_multi_n$N.log
: > gpurun_out/r02m_pcie_probe_n$N.txt
for g in 1 2 4 8; do
  if [ $g -le $N ]; then timeout 300 python tools/pcie_probe_multi.py $g >> gpurun_out/r02m_pcie_probe_n$N.txt 2>&1; fi
done
cat gpurun_out/r02m_pcie_probe_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02m_bench_n$N.json 2> gpurun_out/r02m_bench_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/r02m_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02m_bench_n$N.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "link_probe")})
    print(json.dumps(d.get("strips_8k"), indent=0)[:2500])
except Exception as e:
    print("no bench line:", e)
PY
timeout 600 python bench.py --workload 8k_strips --gpus $N --steps 3 > gpurun_out/r02m_strips_n$N.json 2> gpurun_out/r02m_strips_n$N.err; echo "strips exit $?"; tail -2 gpurun_out/r02m_strips_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_strips_n$N.json')); print(json.dumps(d['strips_8k'], indent=0)[:2500])"
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8 > gpurun_out/r02m_nvlink_sample_n$N.txt
