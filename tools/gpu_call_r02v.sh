#!/bin/bash
# ncu capture of the SAO kernel alone (warm launch of tools/profile_variants.py)
mkdir -p gpurun_out
R=gpurun_out/r02v
N=128 timeout 600 python tools/profile_variants.py > gpurun_out/plain3.log 2>&1 && \
N=128 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:Sao" -f -o ${R}_prof_sao python tools/profile_variants.py > ${R}_ncu_sao.log 2>&1
python tools/ncu_summary.py ${R}_prof_sao.ncu-rep > ${R}_ncu_full_summary_sao.json 2>> ${R}_ncu_sao.log
tail -3 ${R}_ncu_sao.log; du -sm gpurun_out
