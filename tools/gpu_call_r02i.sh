#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
tail -12 gpurun_out/r02i_pytest.log
for wl in 4k_sao256 4k_p010_512 4k_yuv420p10_512; do
  timeout 600 python bench.py --workload $wl > gpurun_out/r02i_bench_$wl.json 2> gpurun_out/r02i_bench_$wl.err; echo "bench $wl exit $?"; tail -2 gpurun_out/r02i_bench_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_$wl.json')); print({k:d.get(k) for k in ('value','ms_per_step','deblock_alone','sao_alone','deblock_plus_sao')}); print(d['roofline']['frac'], d['roofline']['kernel'], d['e2e'].get('value') if 'e2e' in d else None, d['parity'])"
done
timeout 200 python tools/soak.py 11 120 > gpurun_out/r02i_soak.log 2>&1; tail -3 gpurun_out/r02i_soak.log
HDB_REFGPU_LOG=gpurun_out/r02i_refgpu_4k.log timeout 300 python bench.py --impl reference_gpu --steps 5 --warmup 2 > gpurun_out/r02i_bench_refgpu.json 2> gpurun_out/r02i_bench_refgpu.err; echo "refgpu exit $?"; head -c 1200 gpurun_out/r02i_bench_refgpu.json; echo
