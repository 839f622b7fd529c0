"""Summarise where a kernel spends its warp-time from an .ncu-rep (source page, SASS view):
    ncu -i rep.ncu-rep --page source --csv --print-source sass | python tools/ncu_sass_profile.py [nchunks]
Prints, for consecutive chunks of the SASS, the share of stall samples, executed instructions, the dominant
opcodes and the dominant stall reasons."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r)]
h = rows[hi[0]]
si, src, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
names = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_mio', 'stall_math', 'stall_selected',
         'stall_not_selected', 'stall_dispatch', 'stall_branch_resolving', 'stall_no_inst', 'stall_barrier']
idx = [h.index(n) for n in names]
data = []
for r in rows[hi[0] + 1:]:
    try:
        data.append((int(r[si]), r[src], int(r[ie]), [int(r[i]) for i in idx]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
n = len(data)
print('total samples', tot, 'SASS instructions', n, 'warp instructions executed', sum(d[2] for d in data))
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 36
chunk = max(1, n // nch)
for c in range(0, n, chunk):
    seg = data[c:c + chunk]
    s = sum(d[0] for d in seg)
    ie_ = sum(d[2] for d in seg)
    ops = {}
    for v, t, _, _ in seg:
        sp = t.split()
        op = sp[1] if sp and sp[0].startswith('@') and len(sp) > 1 else (sp[0] if sp else '')
        ops[op] = ops.get(op, 0) + 1
    st = [sum(d[3][k] for d in seg) for k in range(len(names))]
    top = sorted(ops.items(), key=lambda x: -x[1])[:3]
    ts = sorted(zip(st, names), reverse=True)[:3]
    print(f'{c:5d} {100 * s / max(tot, 1):5.1f}% ie={ie_ / 1e6:7.1f}M {top} | {[(n_[6:], round(100 * v / max(s, 1))) for v, n_ in ts]}')
