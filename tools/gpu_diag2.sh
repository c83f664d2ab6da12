#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
for b in 16 64; do timeout 200 python tools/fwdprobe2.py $b > gpurun_out/fwdprobe2_$b.txt 2>&1; cat gpurun_out/fwdprobe2_$b.txt; done
