#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_blocks.py tests/test_gpu_bench_config.py tests/test_gpu_host_driver.py -m gpu -x -q 2>&1 | tail -4
for a in "" "--waterfall-skip 0 --pcm16" "--waterfall-skip 0 --pcm16 --e2e-raw s16" "--waterfall-skip 0 --pcm16 --e2e-raw u8"; do
echo "== $a"
timeout 600 python bench.py --no-cpu-baseline --steps 30 $a 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], 'raw', (d.get('e2e_raw') or {}).get('value'), (d.get('e2e_raw') or {}).get('h2d_bytes_per_step'))
"
done
