#!/usr/bin/env python
"""Dynamic SASS opcode mix + stall samples of one kernel from `ncu -i rep --page source --csv`.
usage: ncu -i prof.ncu-rep --page source --csv | python tools/ncu_opmix.py [cells_substeps]"""
import collections, csv, re, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); samples = collections.Counter(); total = 0; tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", r[ci["Source"]])
    if not m: continue
    op = m.group(1)
    n = int(float(r[ci["Instructions Executed"]] or 0)); s = int(float(r[ci["# Samples"]] or 0))
    ops[op] += n; samples[op] += s; total += n; tot_s += s
norm = float(sys.argv[1]) if len(sys.argv) > 1 else None
print(f"total warp instructions {total:.4g}, samples {tot_s}")
for op, n in ops.most_common(30):
    extra = f"  {32*n/norm:7.2f}/cell-substep" if norm else ""
    print(f"{op:10s} {n:14d} {100*n/total:5.1f}%  samples {100*samples[op]/max(tot_s,1):5.1f}%{extra}")
