#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:client_demod_warp' -s 30 -c 1 -f -o gpurun_out/r2_demod360 python tools/cliprobe.py 1024 64 > gpurun_out/r2_ncu_demod360.log 2>&1; tail -2 gpurun_out/r2_ncu_demod360.log
