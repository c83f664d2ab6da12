#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_bench_config.py tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2_pytest_var.txt 2>&1; tail -12 gpurun_out/r2_pytest_var.txt
timeout 600 python bench.py --config cfg3 --no-cpu-baseline --steps 40 > gpurun_out/r2_bench_cfg3c.json 2> gpurun_out/r2_bench_cfg3c.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg3c.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['breakdown'])"
