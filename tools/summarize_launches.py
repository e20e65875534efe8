"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one forward step's kernels and shares.
    python tools/summarize_launches.py gpurun_out/launches.csv <first launch index of a step> <launches per step>
"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID") + 1
L = [(r[4], int(r[-1])) for r in rows[start:] if len(r) > 10]
if len(sys.argv) > 2 and not sys.argv[2].startswith("auto"):
    first = int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else len(L)
else:
    # auto: one steady-state step = the launches between two consecutive occurrences of the step's LAST kernel
    # (default: the fused optimizer update of a training step), taking the 4th such interval (after the warm-up steps)
    marker = sys.argv[2].split(":", 1)[1] if len(sys.argv) > 2 and ":" in sys.argv[2] else "train_step_update"
    idx = [i for i, (nm, _) in enumerate(L) if marker in nm]
    if len(idx) >= 5:
        first, n = idx[3] + 1, idx[4] - idx[3]
    elif len(idx) >= 2:
        first, n = idx[-2] + 1, idx[-1] - idx[-2]
    else:
        first, n = 0, len(L)
step = L[first:first + n]
tot = sum(t for _, t in step)
agg = OrderedDict()
for name, t in step:
    short = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    a = agg.setdefault(short, [0, 0])
    a[0] += t
    a[1] += 1
print(f"launches {len(step)}  total {tot/1e6:.3f} ms (ncu: cold cache, serialised -- compare shares, not absolutes)\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {c} | {t/1e6:.3f} | {100*t/tot:.1f} % |")
print("\nper launch, in order:\n")
for i, (name, t) in enumerate(step):
    short = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    print(f"{i:3d} {t/1e3:9.1f} us  {short}")
