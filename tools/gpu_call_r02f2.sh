#!/bin/bash
# closing run on the library at HEAD: the whole GPU suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
R=gpurun_out/r02f
timeout 2400 python -m pytest tests -m gpu -q > ${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${R}_pytest_gpu.log; tail -3 ${R}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${R}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 ${R}_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > ${R}_bench_reference_arm.json 2> ${R}_bench_reference_arm.err; echo "reference arm exit $?"
timeout 900 python bench.py > ${R}_bench_n1.json 2> ${R}_bench_n1.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("${R}_bench_n1.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "frac_of_link", d["e2e"].get("frac_of_link"), "launches", d.get("gpu_launches"))
print("parity", d.get("parity"))
print("other", [(o.get("config", {}).get("workload"), o.get("roofline", {}).get("frac")) for o in d.get("other_configs", [])])
r = json.load(open("${R}_bench_reference_arm.json"))
print("reference", r.get("value"), r.get("unit"), r.get("cpu_baseline", {}).get("cores"))
PY
