# development check: e2e leg with different sub-batch sizes
for C in 8 16 32 64; do timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 8 --e2e-chunk $C | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk', $C, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'raw', round(d['e2e_raw_uint16']['value']))"; done
