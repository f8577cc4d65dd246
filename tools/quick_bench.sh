#!/bin/bash
# usage: tools/quick_bench.sh [label]   (env vars pass through); prints the key numbers of one short bench run
python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-run-simulation 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'shots/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'stage_ms', {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if k!='note'}, 'LER', round(d['logical_error_rate'],4), 'e2e', round(d['e2e']['value']))
    else: print(l.strip()[:300])
"
