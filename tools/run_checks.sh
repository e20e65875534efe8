#!/bin/bash
# correctness groups (one process each) + layer timings; prints compact lines
fmt_ok='
import sys, json
for l in sys.stdin:
    try:
        r = json.loads(l)
        print(r["case"], "OK" if r.get("ok") else "FAIL " + json.dumps(r))
    except Exception:
        print(l.rstrip()[:300])
'
fmt_t='
import sys, json
for l in sys.stdin:
    try:
        r = json.loads(l)
        print(r["case"], r.get("tune"), "ms=%.4f tflops=%.1f" % (r.get("ms", -1), r.get("tflops", -1)), r.get("error", ""), r.get("pairs_per_s", ""))
    except Exception:
        print(l.rstrip()[:300])
'
for g in ${GROUPS_OK-tap halo misc model}; do echo "=== $g"; timeout 300 python tools/gpu_check.py $g 2>&1 | python -c "$fmt_ok" | grep -v " OK$"; done
for g in ${GROUPS_T-bench fwd}; do echo "=== $g"; timeout 300 python tools/gpu_check.py $g 2>&1 | python -c "$fmt_t"; done
