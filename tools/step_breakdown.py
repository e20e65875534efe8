"""Aggregate one steady-state step from an ncu launch list (steps are delimited by the two pack_nchw launches)."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID") + 1
L = [(r[4], int(r[-1])) for r in rows[start:] if len(r) > 10]
idx = [i for i in range(len(L) - 1) if 'pack_nchw' in L[i][0] and 'pack_nchw' in L[i + 1][0]]
step = L[idx[which]:idx[which + 1]]
tot = sum(t for _, t in step)
agg = OrderedDict()
for n, t in step:
    k = n.replace('void ', '').replace('<unnamed>::', '').split('(')[0][:70]
    a = agg.setdefault(k, [0, 0])
    a[0] += t
    a[1] += 1
print(f"step launches {len(step)}  total {tot/1e6:.3f} ms (ncu: cold cache, serialised -- compare shares)\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if t / tot > 0.001:
        print(f"| `{k}` | {c} | {t/1e6:.3f} | {100*t/tot:.1f} % |")
