"""Summarise the whole-step `ncu --set full` raw-page CSV written by tools/prof_round2.sh (read here, no GPU needed):
    python tools/ncu_step_summary.py gpurun_out/r02k_ncu_full_train_step.raw.csv.gz [--per-launch]
One row per kernel (template instantiation): launches, time, DRAM bytes and achieved DRAM GB/s, DRAM %, tensor-pipe %,
registers, achieved warps.  ncu runs each kernel alone, cold cache, ~40 replays: compare SHARES, not absolutes."""
import csv
import gzip
import io
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
per_launch = "--per-launch" in sys.argv
f = gzip.open(path) if path.endswith(".gz") else open(path, "rb")
rows = list(csv.reader(io.TextIOWrapper(f)))
h, u = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def val(r, k):
    v = r[col[k]]
    return float(v.replace(",", "")) * SCALE.get(u[col[k]], 1.0) if v not in ("", "n/a") else 0.0


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    return name.split("(")[0][:70]


T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
DP, TP = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
RG, WA = "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"
agg = OrderedDict()
total = 0.0
for r in rows[2:]:
    n = short(r[col["Kernel Name"]])
    a = agg.setdefault(n, dict(n=0, t=0.0, b=0.0, dp=0.0, tp=0.0, regs=0, wa=0.0, launches=[]))
    t = val(r, T)
    a["n"] += 1; a["t"] += t; a["b"] += val(r, RD) + val(r, WR)
    a["dp"] += val(r, DP) * t; a["tp"] += val(r, TP) * t; a["wa"] += val(r, WA) * t
    a["regs"] = int(val(r, RG))
    a["launches"].append((t, val(r, RD) + val(r, WR), val(r, DP), val(r, TP)))
    total += t
print(f"{len(rows) - 2} launches, {total / 1e3:.3f} ms in total (each kernel alone, cold cache)\n")
print("| kernel | launches | ms | share | DRAM GB | GB/s | DRAM % | tensor pipe % | regs | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"]
    print(f"| `{n}` | {a['n']} | {t / 1e3:.3f} | {100 * t / total:.1f} % | {a['b'] / 1e9:.2f} | {a['b'] / (t * 1e-6) / 1e9:.0f} | "
          f"{a['dp'] / t:.0f} | {a['tp'] / t:.0f} | {a['regs']} | {a['wa'] / t:.0f} |")
    if per_launch and a["n"] > 1:
        for (lt, lb, ldp, ltp) in a["launches"]:
            print(f"| &nbsp;&nbsp;· | | {lt / 1e3:.3f} | | {lb / 1e9:.2f} | {lb / (lt * 1e-6) / 1e9:.0f} | {ldp:.0f} | {ltp:.0f} | | |")
