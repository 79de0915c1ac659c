"""Selected metrics of every launch in one or more .ncu-rep files -> JSON (profiles/).  Usage: python tools/ncu_summary.py out.json a.ncu-rep [b.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
out = {}
for rep in sys.argv[2:]:
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        rec = {k: f"{r[idx[k]]} {units[idx[k]]}".strip() for k in KEYS if k in idx}
        top = sorted(((float(r[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0]) for h in stall if r[idx[h]]), reverse=True)[:5]
        rec["top_stalls_per_issue"] = {n: round(v, 2) for v, n in top}
        rec["report"] = rep.split("/")[-1]
        out.setdefault(name, []).append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print({k: len(v) for k, v in out.items()})
