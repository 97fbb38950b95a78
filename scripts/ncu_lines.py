#!/usr/bin/env python
"""Per-source-line instruction / stall summary of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 4 and r[0] == "Line No")
hdr = rows[hi]
ii = hdr.index("Instructions Executed"); ws = hdr.index("Warp Stall Sampling (All Samples)")
items = []
for r in rows[hi + 1:]:
    if r and r[0].isdigit():
        try: items.append((int(r[ii]), int(r[ws]), int(r[0]), r[1][:110]))
        except ValueError: pass
tot = sum(i[0] for i in items); tw = sum(i[1] for i in items)
print("total warp-instr", tot, "stall samples", tw)
print("-- by instructions"); [print("%10d %5.1f%%  stall %5.1f%%  L%-4d %s" % (n, 100*n/tot, 100*w/tw, l, s)) for n, w, l, s in sorted(items, reverse=True)[:top]]
print("-- by stall samples"); [print("%10d %5.1f%%  stall %5.1f%%  L%-4d %s" % (n, 100*n/tot, 100*w/tw, l, s)) for n, w, l, s in sorted(items, key=lambda x: -x[1])[:top]]
