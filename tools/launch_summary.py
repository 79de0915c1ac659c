"""Totals per kernel name of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_summary.py file.csv [top]"""
import csv, re, sys
rows = {}
order = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    us = v / 1000 if u.startswith("ns") else v if u.startswith("us") else v * 1000
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
    name = re.sub(r"^void ", "", name)[:90]
    if name not in rows:
        rows[name] = [0.0, 0]
    rows[name][0] += us; rows[name][1] += 1
tot = sum(v[0] for v in rows.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for name, (us, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-92s %10.1f us %6d x %8.2f  %5.1f %%" % (name, us, n, us / n, 100 * us / tot))
print("total %.1f us over %d launches" % (tot, sum(v[1] for v in rows.values())))
