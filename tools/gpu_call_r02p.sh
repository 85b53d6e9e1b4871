#!/bin/bash
# evidence refresh after the SAO rewrite and the cp.async change: the default bench line, the ncu summary of one warm
# launch of every kernel of the library (256 frames), smoke
mkdir -p gpurun_out
R=gpurun_out/r02p
timeout 900 python bench.py > ${R}_bench_n1.json 2> ${R}_bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('${R}_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_link'))"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${R}_smoke.log 2>&1; echo "smoke exit $?"
N=256 timeout 600 python tools/profile_variants.py > gpurun_out/plain3.log 2>&1 && \
N=256 ncu --set full --clock-control none --profile-from-start off -k "regex:Deblock|Derive|Sao" -f -o ${R}_prof_variants python tools/profile_variants.py > ${R}_ncu_variants.log 2>&1
python tools/ncu_summary.py ${R}_prof_variants.ncu-rep > ${R}_ncu_full_summary_variants.json 2>> ${R}_ncu_variants.log
rm -f ${R}_prof_variants.ncu-rep
tail -2 ${R}_ncu_variants.log; du -sm gpurun_out
