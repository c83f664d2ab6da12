#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:client_tail2' -s 3 -c 1 \
    -f -o gpurun_out/r2_tail2 python tools/tailprof.py 64 > gpurun_out/r2_ncu_tail2.log 2>&1
tail -3 gpurun_out/r2_ncu_tail2.log
