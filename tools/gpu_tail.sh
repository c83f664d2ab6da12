#!/bin/bash
# tail pipeline: parity tests, stage profile, timing alone and inside a step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clients.py tests/test_gpu_variants.py::test_chunked_demodulation_equals_sequential tests/test_gpu_waterfall.py tests/test_gpu_stream_blocks.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/tailprof.py 2>&1 | tail -12
for o in "" "24=224" "24=160"; do
echo "== B200_OPTS=$o"
B200_OPTS=$o timeout 300 python tools/cliprobe.py 2>&1 | grep "tails"
B200_OPTS=$o timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], json.dumps(d.get('breakdown', {}))[:300])
"
done
B200_OPTS=$o timeout 300 python bench.py --clients 1 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('1 client', d['value'], d['ms_per_step'], json.dumps(d.get('breakdown', {}))[:300])
"
