#!/bin/bash
for o in "20=3" "20=0" "20=1" "20=2"; do
for c in 1024; do
echo "== B200_OPTS=$o clients $c"
B200_OPTS=$o timeout 600 python bench.py --no-cpu-baseline --steps 10 --clients $c --waterfall-skip 0 --pcm16 --e2e-raw u8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), 'e2e', round(d['e2e']['value']), 'raw', round((d.get('e2e_raw') or {}).get('value',0)))
"
done
done
