#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_variants.py -x -q 2>&1 | tail -6
timeout 200 python tools/fwdprobe2.py 64 2>&1 | tail -8
