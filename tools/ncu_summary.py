"""Extract of an ncu report for profiles/: per profiled launch the metrics the roofline discussion uses.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_summary.json   (runs `ncu -i ... --page raw --csv`)"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_ldgsts.sum", "lts__t_sectors_srcunit_tex_aperture_peer.sum", "lts__t_sectors_srcunit_tex_aperture_peer_op_read.sum",
        "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
STALL = "smsp__average_warp_latency_issue_stalled_"
STALL2 = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    e = {"id": d.get("ID"), "kernel": d.get("Kernel Name")}
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            try:
                e[k] = float(d[k].replace(",", ""))
            except ValueError:
                e[k] = d[k]
            if u.get(k):
                e.setdefault("_units", {})[k] = u[k]
    stalls = {}
    for k, v in d.items():
        if (k.startswith(STALL) or k.startswith(STALL2)) and k.endswith("_per_issue_active.ratio") or (k.startswith(STALL2) and k.endswith(".ratio") and "not_issued" not in k):
            try:
                f = float(v)
            except ValueError:
                continue
            if f >= 0.1:
                stalls[k.replace(STALL, "").replace(STALL2, "").replace("_per_issue_active.ratio", "").replace(".ratio", "")] = round(f, 3)
    e["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    if "dram__bytes_read.sum" in e and "dram__bytes_write.sum" in e:
        e["dram_bytes_total"] = e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
    res.append(e)
json.dump({"report": rep.split("/")[-1], "how": "ncu --set full --clock-control none --import-source on; durations are cold-cache and serialised: never a timing source", "launches": res}, sys.stdout, indent=1)
