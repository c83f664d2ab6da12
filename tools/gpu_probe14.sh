#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clients.py tests/test_gpu_variants.py tests/test_gpu_waterfall.py tests/test_gpu_stream_blocks.py tests/test_golden.py tests/test_gpu_bench_config.py -m gpu -x -q > gpurun_out/r2_pytest_tail4.txt 2>&1
tail -5 gpurun_out/r2_pytest_tail4.txt
timeout 300 python tools/cliprobe.py > gpurun_out/r2_cliprobe2.txt 2>&1; cat gpurun_out/r2_cliprobe2.txt
timeout 300 python tools/tailprof.py > gpurun_out/r2_tailprof2.txt 2>&1; cat gpurun_out/r2_tailprof2.txt
for o in "" "22=16"; do
B200_OPTS=$o timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d.get('breakdown', {}))[:400])
"
done
