"""Per-instruction stall samples of one kernel from an .ncu-rep (ncu --set full --import-source on): totals by stall reason and the
top instructions.  Usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, sys, subprocess
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data), 'executed warp-inst', sum(int(r[ix['Instructions Executed']]) for r in data))
for st in ['stall_long_sb','stall_wait','stall_math','stall_not_selected','stall_selected','stall_short_sb','stall_branch_resolving','stall_no_inst','stall_dispatch','stall_lg','stall_mio','stall_barrier','stall_membar','stall_drain']:
    v = sum(int(r[ix[st]]) for r in data)
    if v: print('  %-24s %6d  %.1f%%' % (st, v, 100.0 * v / tot))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:ntop]
for i in sorted(top):
    r = data[i]
    st = {k: int(r[ix[k]]) for k in ix if k.startswith('stall_') and '(' not in k and int(r[ix[k]]) > 0}
    print(i, r[ix['Source']].strip()[:64], '| samples', r[ix['# Samples']], 'exec', r[ix['Instructions Executed']], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
