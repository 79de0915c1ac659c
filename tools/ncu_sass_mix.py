"""Executed-instruction mix of one kernel from an .ncu-rep (source page, SASS view).
Usage: python tools/ncu_sass_mix.py report.ncu-rep kernel_regex [launch_index]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout.splitlines()
blocks, cur = [], None
for row in csv.reader(out):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and "hdr" in cur:
        cur["rows"].append(row)
b = blocks[which]
h = {k: i for i, k in enumerate(b["hdr"])}
ops, samples, wf = Counter(), Counter(), Counter()
tot = 0
for r in b["rows"]:
    sass = r[h["Source"]].strip()
    toks = sass.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "RED", "ATOM", "SHFL", "MUFU")) and "." in op else "")
    n = int(r[h["Instructions Executed"]])
    ops[op] += n
    samples[op] += int(r[h["# Samples"]])
    wf[op] += int(r[h["L1 Wavefronts Shared"]])
    tot += n
print(b["name"][:100])
print(f"total warp instructions {tot}")
stot = sum(samples.values())
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * samples[op] / max(stot, 1):5.1f}%   smem wavefronts {wf[op]}")
