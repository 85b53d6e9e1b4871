#!/bin/bash
# round 2, the record: bench lines of every workload, both reference arms, launch list, ncu --set full captures
mkdir -p gpurun_out
R=gpurun_out/r02z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit --format=csv > ${R}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > ${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${R}_pytest_gpu.log
tail -6 ${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${R}_smoke.log 2>&1; tail -2 ${R}_smoke.log
timeout 900 python bench.py > ${R}_bench_n1.json 2> ${R}_bench_n1.err; echo "bench exit $?"; tail -2 ${R}_bench_n1.err
timeout 600 python bench.py --impl reference > ${R}_bench_reference_arm.json 2> ${R}_bench_reference_arm.err; echo "reference arm exit $?"
for wl in 4k_nv12_512 4k_edges512 4k_yuv420p10_512 4k_p010_512 4k_sao256; do
  timeout 600 python bench.py --workload $wl > ${R}_bench_$wl.json 2> ${R}_bench_$wl.err; echo "bench $wl exit $?"; tail -1 ${R}_bench_$wl.err
done
HDB_REFGPU_LOG=${R}_refgpu_4k.log timeout 300 python bench.py --impl reference_gpu > ${R}_bench_reference_gpu_4k.json 2> ${R}_refgpu.err
HDB_REFGPU_LOG=${R}_refgpu_1080p.log timeout 300 python bench.py --impl reference_gpu --workload 1080p_batch256 > ${R}_bench_reference_gpu_1080p.json 2>> ${R}_refgpu.err
HDB_REFGPU_LOG=${R}_refgpu_image2.log timeout 300 python bench.py --impl reference_gpu --workload image2_single > ${R}_bench_reference_gpu_image2.json 2>> ${R}_refgpu.err
: > ${R}_ldgsts.txt
for sz in "1080 1920 256" "1080 1920 1024" "360 288 256" "1000 1000 256" "1008 1000 256" "1920 1080 256"; do set -- $sz
  timeout 120 python tools/quick_time.py --w $1 --h $2 --n $3 >> ${R}_ldgsts.txt 2>&1
done
HDB_NO_LDGSTS=1 timeout 120 python tools/quick_time.py --w 1080 --h 1920 --n 256 >> ${R}_ldgsts.txt 2>&1
timeout 300 python tools/soak.py 21 150 > ${R}_soak.txt 2>&1; tail -2 ${R}_soak.txt
ls -la gpurun_out | grep r02z
