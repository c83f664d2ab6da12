#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_variants.py::test_raw_sample_formats_on_the_tma_path tests/test_golden.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -12
python tools/rawprobe.py 64
timeout 600 python bench.py --no-cpu-baseline --steps 20 --waterfall-skip 0 --pcm16 --e2e-raw s16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), 'e2e', round(d['e2e']['value']), 'raw', round((d.get('e2e_raw') or {}).get('value',0)))
"
timeout 600 python bench.py --no-cpu-baseline --steps 20 --waterfall-skip 0 --pcm16 --e2e-raw u8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), 'e2e', round(d['e2e']['value']), 'raw', round((d.get('e2e_raw') or {}).get('value',0)))
"
