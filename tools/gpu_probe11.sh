#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py::test_chunked_demodulation_equals_sequential tests/test_gpu_clients.py tests/test_gpu_stream_blocks.py tests/test_golden.py tests/test_gpu_bench_config.py::test_cfg2_iq_2p20_1024_clients_batch64 -x -q -m gpu > gpurun_out/r2_pytest_demod2.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_demod2.txt
timeout 300 python tools/cliprobe.py 1024 64 2>&1 | tee gpurun_out/r2_cliprobe.txt
timeout 300 python tools/cliprobe.py 8192 64 2>&1 | tee -a gpurun_out/r2_cliprobe.txt
