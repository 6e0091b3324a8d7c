#!/usr/bin/env python3
"""Top SASS lines by warp-stall samples from `ncu --page source --csv --print-source sass` (one block per kernel).
    python tools/ncu_source_top.py file.source.csv [kernel_index] [top_n]"""
import csv, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
for row in csv.reader(open(path)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = row
        else:
            cur["rows"].append(row)
b = blocks[which]
h = b["hdr"]
print(b["name"][:150])
ia, isrc, isamp = h.index("Address"), h.index("Source"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
iexec = h.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in b["rows"])
print("total samples", tot, " instructions executed", sum(int(r[iexec] or 0) for r in b["rows"]))
agg = {}
for r in b["rows"]:
    for i, c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
rows = sorted(b["rows"], key=lambda r: -int(r[isamp] or 0))[:topn]
for r in sorted(rows, key=lambda r: int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia] or 0)):
    st = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols if int(r[i] or 0)), reverse=True)[:3]
    print("%6s %5d %8s  %-70s %s" % (r[ia][-5:], int(r[isamp] or 0), r[iexec], r[isrc][:70], st))
