"""Executed FP64-pipe instructions of a kernel from the SOURCE page of an .ncu-rep captured with --import-source on
(`--set full` has the pipe's utilisation but not its instruction count): sums "Instructions Executed" over the SASS lines
whose opcode is DFMA / DMUL / DADD / DSETP / DMNMX, per captured launch, and appends the row
  sm__inst_executed_pipe_fp64.sum,inst,<launch0>,<launch1>,...   (warp-level, derived)
to the summary CSV that bench.py reads.   python tools/ncu_fp64_count.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
counts = []
for b in blocks:
    hdr = b[0]
    i_s, i_e = hdr.index("Source"), hdr.index("Instructions Executed")
    n = 0
    for r in b[1:]:
        if len(r) <= i_e:
            continue
        tok = r[i_s].split()
        if not tok:
            continue
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0].rstrip(";")
        if op in FP64:
            n += int(r[i_e])
    counts.append(n)
counts = counts[::2] if len(counts) > 1 and counts[0::2] == counts[1::2] else counts   # ncu prints each launch's page twice
rows = [r for r in csv.reader(open(sys.argv[2])) if r and r[0] != "sm__inst_executed_pipe_fp64.sum"]
rows.append(["sm__inst_executed_pipe_fp64.sum", "inst (warp-level; derived from the source page: " + "/".join(FP64) + ")"] + [str(c) for c in counts])
csv.writer(open(sys.argv[2], "w", newline="")).writerows(rows)
print(counts)
