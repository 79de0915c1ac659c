"""Per-launch summary (time, occupancy, issue utilisation, top stall reasons) from an .ncu-rep.  Usage: python tools/ncu_stalls.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio")]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "")[:28]
    vals = " ".join(f"{k.split('__')[1].split('.')[0][:16]}={r[idx[k]]}" for k in keys if k in idx)
    v = sorted([(float(r[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0]) for h in st if r[idx[h]]], reverse=True)
    print(r[idx["ID"]].rjust(3), name.ljust(28), vals)
    print("      stalls/issue:", ", ".join(f"{n} {x:.2f}" for x, n in v[:7]))
