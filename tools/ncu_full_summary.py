"""Summarise `ncu --set full` reports or their `--page raw --csv` exports (read here, no GPU needed):
    python tools/ncu_full_summary.py rep1.ncu-rep x.raw.csv ...
Prints a markdown table of the metrics the roofline discussion uses and a JSON dict of DRAM bytes per launch."""
import csv
import io
import json
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__cycles_elapsed.max", "sm cycles"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
out = {}
print("| report | kernel | " + " | ".join(k for _, k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    txt = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, zip(u, v)))
        name = d["Kernel Name"][1].split("(")[0].replace("void ", "")
        cells = []
        for k, _ in KEYS:
            unit, val = d.get(k, ("", ""))
            cells.append(f"{float(val):.4g} {unit}" if val else "-")
        print(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(cells) + " |")
        tot = sum(float(d[k][1]) * SCALE.get(d[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out[rep.split("/")[-1]] = {"kernel": name, "dram_bytes_per_launch": tot, "time_us": float(d["gpu__time_duration.sum"][1])}
print()
print(json.dumps(out, indent=1))
