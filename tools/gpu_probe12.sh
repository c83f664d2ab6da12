#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py::test_chunked_demodulation_equals_sequential tests/test_gpu_clients.py tests/test_gpu_stream_blocks.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:client_demod_warp' -s 1 -c 1 -f -o gpurun_out/r2_demod python tools/cliprobe.py 1024 64 > gpurun_out/r2_ncu_demod.log 2>&1; tail -2 gpurun_out/r2_ncu_demod.log
