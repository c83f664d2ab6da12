#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:client_' -s 8 -c 2 \
    -f -o gpurun_out/cli python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cli.log 2>&1
tail -3 gpurun_out/ncu_cli.log
ls -la gpurun_out/cli.ncu-rep
