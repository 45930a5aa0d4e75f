"""Executed-instruction histogram by SASS opcode from an ncu report (usage: sass_hist.py report.ncu-rep)."""
import csv, re, subprocess, sys
from collections import defaultdict
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
ops, tot = defaultdict(float), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break
        continue
    try:
        e = float(r[ie])
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    op = (m.group(2) if m else r[ia]).split(".")[0]
    ops[op] += e
    tot += e
print("executed warp-instructions: %.0f" % tot)
for k in sorted(ops, key=ops.get, reverse=True)[:24]:
    print("%-10s %12.0f %5.1f%%" % (k, ops[k], 100 * ops[k] / tot))
