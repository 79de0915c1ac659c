"""Pivot an `ncu --csv --metrics ...` log into one line per kernel launch.  Usage: python tools/ncu_csv.py log.csv"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
d = OrderedDict()
for r in rows[h + 1:]:
    rec = dict(zip(hdr, r))
    d.setdefault((rec["ID"], rec["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")[:24]), {})[rec["Metric Name"]] = rec["Metric Value"]
short = {"gpu__time_duration.sum": "ns", "smsp__inst_executed.sum": "inst", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wf",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "conflicts", "sm__cycles_active.avg": "cyc_active",
         "smsp__inst_executed_op_shared_ld.sum": "lds"}
for k, v in d.items():
    print(k[0].rjust(3), k[1].ljust(24), "  ".join(f"{short.get(a, a)}={b}" for a, b in v.items()))
