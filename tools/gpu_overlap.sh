#!/bin/bash
# which client kernel bounds a step: bench legs with the demodulation and / or the tails masked out
for c in 1024 1; do
for o in "20=0" "20=1" "20=2" "20=3"; do
echo "== clients $c B200_OPTS=$o (bit0 demod, bit1 tails)"
B200_OPTS=$o timeout 300 python bench.py --clients $c --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), 'MS/s', round(d['ms_per_step']*1000/64,2), 'us/frame', json.dumps(d.get('breakdown', {}))[:200])
"
done
done
