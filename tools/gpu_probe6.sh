#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clients.py tests/test_gpu_stream_blocks.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2_pytest_tail2.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_tail2.txt
timeout 200 python tools/tailprof.py 64 2>&1 | tee gpurun_out/r2_tailprof.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_tail2.json 2> gpurun_out/r2_bench_tail2.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_tail2.json'));print(d['value'],d['ms_per_step'],d['breakdown'])"
