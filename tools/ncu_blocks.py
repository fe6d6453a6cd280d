"""Executed warp instructions of one kernel by contiguous SASS block (same execution count), from an .ncu-rep source page.
Usage: python tools/ncu_blocks.py report.ncu-rep"""
import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
blocks = []
for i, r in enumerate(data):
    ex = int(r[ix['Instructions Executed']]); th = int(r[ix['Thread Instructions Executed']])
    if blocks and blocks[-1][2] == ex: blocks[-1][1] = i; blocks[-1][3] += th; blocks[-1][4] += int(r[ix['# Samples']])
    else: blocks.append([i, i, ex, th, int(r[ix['# Samples']])])
tot = sum((b[1]-b[0]+1)*b[2] for b in blocks)
print("total warp-inst", tot)
for b in blocks:
    n = b[1]-b[0]+1
    if n*b[2] > tot*0.004:
        print("%5d-%5d n=%4d exec/inst=%6d warp-inst=%8d (%.1f%%) avg-threads=%.1f samples=%d  %s" % (b[0], b[1], n, b[2], n*b[2], 100.0*n*b[2]/tot, b[3]/max(1,n*b[2]), b[4], data[b[0]][ix['Source']].strip()[:40]))
