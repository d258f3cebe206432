"""Split one kernel's SASS (ncu --page source --print-source sass --csv) at BAR.SYNC instructions and print, per region,
executed warp instructions, stall samples by reason and the instruction mix -- regions map to the FFT passes of a kernel.

usage: sass_regions.py src.csv [invocation_index]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# the csv holds one table per profiled launch: split on the "Kernel Name" rows
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
t = tables[which]
hdr = t["rows"][0]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(t["name"])
regions = [[]]
for r in t["rows"][1:]:
    if len(r) < len(hdr):
        continue
    regions[-1].append(r)
    if r[ix["Source"]].strip().startswith("BAR"):
        regions.append([])
tot_i = sum(int(r[ix["Instructions Executed"]]) for reg in regions for r in reg)
tot_s = sum(int(r[ix["# Samples"]]) for reg in regions for r in reg)
print("total warp instr %d  samples %d" % (tot_i, tot_s))
for k, reg in enumerate(regions):
    ni = sum(int(r[ix["Instructions Executed"]]) for r in reg)
    ns = sum(int(r[ix["# Samples"]]) for r in reg)
    if ni == 0:
        continue
    st = Counter()
    ops = Counter()
    for r in reg:
        for s in stalls:
            st[s[6:]] += int(r[ix[s]] or 0)
        op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
        ops[op[0].split(".")[0] if op else "?"] += int(r[ix["Instructions Executed"]])
    top = ", ".join("%s %.0f%%" % (a, 100 * b / max(ns, 1)) for a, b in st.most_common(5))
    mix = ", ".join("%s %.0f%%" % (a, 100 * b / ni) for a, b in ops.most_common(6))
    print("region %2d: %4d sass  inst %5.1f%%  samples %5.1f%%  (samples/inst x%.2f) | %s | %s" %
          (k, len(reg), 100 * ni / tot_i, 100 * ns / tot_s, (ns / tot_s) / (ni / tot_i), top, mix))
