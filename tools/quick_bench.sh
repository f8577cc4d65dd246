#!/bin/bash
# usage: tools/quick_bench.sh [label]   (env vars pass through); prints the key numbers of one short bench run
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'shots/s', round(d['value']), 'stage_ms', {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'LER', d['logical_error_rate'], 'Gedge/s', round(d['roofline_smem']['edge_messages_per_s']/1e9,1), 'e2e', round(d['e2e']['value']))
    else: print(l.strip()[:300])
"
