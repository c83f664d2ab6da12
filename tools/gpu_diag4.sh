#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
for b in 16 32 64; do
  timeout 300 python bench.py --steps 30 --warmup 3 --batch $b --no-cpu-baseline --no-e2e > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python -c "import json;d=json.load(open('gpurun_out/bench_b$b.json'));print($b, d['value'], d['roofline']['frac'], d['breakdown'])"
done
