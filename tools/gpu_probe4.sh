#!/bin/bash
mkdir -p gpurun_out
B200_OPTS=6=4 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fwd_stream' -s 1 -c 1 \
    -f -o gpurun_out/r2_stream python tools/fwdonce.py 64 2 > gpurun_out/r2_ncu_stream.log 2>&1
tail -3 gpurun_out/r2_ncu_stream.log
