"""Executed warp instructions per CUDA source line for one kernel: joins the SASS page of an .ncu-rep with nvdisasm line info of the
matching cubin.  Usage: python tools/ncu_lines.py report.ncu-rep lib.so kernel_substring [launch_index] [top_n]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

rep, lib, kname = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    sect, cur, chain, fresh = None, None, [], True
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect = m.group(1)
            lines.setdefault(sect, [])
            continue
        if ln.startswith("\t.section") or ln.startswith("//-----"):
            if not ln.startswith("//-----"):
                sect = None
            continue
        if sect is None:
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
        if m:
            fr = (os.path.basename(m.group(1)), int(m.group(2)))
            if fresh:
                chain = []
                fresh = False
            chain.append(fr)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines[sect].append((int(m.group(1), 16), m.group(2).strip(), tuple(chain)))
            fresh = True
sect = [k for k in lines if kname.split("<")[0].split("(")[0] in k and lines[k]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
blocks, curb = [], None
for row in csv.reader(out):
    if not row:
        continue
    if row[0] == "Kernel Name":
        curb = {"name": row[1], "rows": []}
        blocks.append(curb)
    elif row[0] == "Address":
        curb["hdr"] = row
    elif curb is not None and "hdr" in curb:
        curb["rows"].append(row)
cands = [b for b in blocks if kname in b["name"].replace("<unnamed>::", "")]
b = cands[which]
h = {k: i for i, k in enumerate(b["hdr"])}
n = len(b["rows"])
match = [s for s in sect if len(lines[s]) == n]
if not match:
    print("no cubin section with", n, "instructions; candidates:", [(s[-60:], len(lines[s])) for s in sect])
    sys.exit(1)
info = lines[match[0]]
per_line, per_outer, tot = Counter(), Counter(), 0
for r, (off, sass, chain) in zip(b["rows"], info):
    c = int(r[h["Instructions Executed"]])
    tot += c
    inner = chain[0] if chain else ("?", 0)
    outer = chain[-1] if chain else ("?", 0)
    per_line[inner] += c
    per_outer[outer] += c
print(b["name"][:90], "total", tot)
rng = os.environ.get("LINES")  # e.g. LINES=engine.cu:599-698 -> inclusive counts per line of that range (first frame of the chain inside it)
if rng:
    f0, r = rng.split(":")
    lo, hi = [int(v) for v in r.split("-")]
    inc, smp, lsb, ssb = Counter(), Counter(), Counter(), Counter()
    for r_, (off, sass, chain) in zip(b["rows"], info):
        c = int(r_[h["Instructions Executed"]])
        for (f, l) in chain:
            if f == f0 and lo <= l <= hi:
                inc[l] += c
                smp[l] += int(r_[h["# Samples"]])
                lsb[l] += int(r_[h["stall_long_sb"]])
                ssb[l] += int(r_[h["stall_short_sb"]]) + int(r_[h["stall_mio"]])
                break
    stot = sum(smp.values())
    src = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f0)).read().splitlines() if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f0)) else None
    for l in sorted(inc):
        if inc[l] * 200 >= tot or smp[l] * 100 >= stot:
            print(f"  {l:5d} {inc[l]:11d} {100 * inc[l] / tot:5.1f}% smp {100 * smp[l] / max(stot, 1):5.1f}% (long {100 * lsb[l] / max(stot, 1):4.1f} short/mio {100 * ssb[l] / max(stot, 1):4.1f})  {src[l - 1].strip()[:90] if src else ''}")
    sys.exit(0)
print("-- by innermost source line")
for (f, l), c in per_line.most_common(topn):
    print(f"  {f}:{l:<5d} {c:11d} {100 * c / tot:5.1f}%")
print("-- by outermost (kernel body) line")
for (f, l), c in per_outer.most_common(topn):
    print(f"  {f}:{l:<5d} {c:11d} {100 * c / tot:5.1f}%")
