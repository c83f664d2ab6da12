#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
timeout 900 python -m pytest tests/test_gpu_bench_config.py -x -q -s > gpurun_out/r2_pytest_benchcfg.txt 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_benchcfg.txt
