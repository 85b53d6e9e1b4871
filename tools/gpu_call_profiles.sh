#!/bin/bash
# round 2, the profiles: launch list of the default bench, ncu --set full of the hot kernel on the full batch and of
# every kernel of the library on 256 frames; summaries are extracted on the box (the reports are large)
mkdir -p gpurun_out
R=gpurun_out/r02z
# every launch of the default bench with its device time (shares, not absolutes)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-sub > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sub > ${R}_ncu_launches.log 2>&1
# the hot kernel on the full batch, and every kernel of the library on 256 frames
timeout 300 python tools/quick_time.py --n 1024 --iters 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:DeblockTiled -s 3 -c 1 -f -o ${R}_prof_main_n1024 python tools/quick_time.py --n 1024 --iters 2 > ${R}_ncu_main.log 2>&1
N=256 timeout 600 python tools/profile_variants.py > gpurun_out/plain3.log 2>&1 && \
N=256 ncu --set full --clock-control none --profile-from-start off -k "regex:Deblock|Derive|Sao" -f -o ${R}_prof_variants python tools/profile_variants.py > ${R}_ncu_variants.log 2>&1
python tools/ncu_summary.py ${R}_prof_main_n1024.ncu-rep > ${R}_ncu_full_summary_main_n1024.json 2>> ${R}_ncu_main.log
python tools/ncu_summary.py ${R}_prof_variants.ncu-rep > ${R}_ncu_full_summary_variants.json 2>> ${R}_ncu_variants.log
du -sm gpurun_out; if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f ${R}_prof_variants.ncu-rep; echo "variants report dropped (size)"; fi
tail -2 ${R}_ncu_variants.log
ls -la gpurun_out | grep r02z
