#!/bin/bash
for g in 8 4 6 12; do
HIPGP_E2E_GROUP=$g timeout 300 python bench.py --quick --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('group', $g, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['e2e']['frac_of_value'],4), 'ms', round(d['e2e']['ms_per_step'],2))"
done
