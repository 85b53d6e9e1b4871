#!/bin/bash
# source-level counters of the cp.async variant on portrait 1080x1920
mkdir -p gpurun_out
R=gpurun_out/r02x
timeout 300 python tools/quick_time.py --w 1080 --h 1920 --n 256 --iters 5 > ${R}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:DeblockTiled -s 3 -c 1 -f -o ${R}_prof_ldgsts python tools/quick_time.py --w 1080 --h 1920 --n 256 --iters 2 > ${R}_ncu.log 2>&1
ncu -i ${R}_prof_ldgsts.ncu-rep --page source --csv > ${R}_source.csv 2>> ${R}_ncu.log
python tools/ncu_summary.py ${R}_prof_ldgsts.ncu-rep > ${R}_ncu_full_summary_ldgsts.json 2>> ${R}_ncu.log
rm -f ${R}_prof_ldgsts.ncu-rep
tail -2 ${R}_plain.log; du -sm gpurun_out
