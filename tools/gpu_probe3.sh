#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fwdprobe.py 64 "stream" > gpurun_out/r2_probe_stream64.txt 2>&1; cat gpurun_out/r2_probe_stream64.txt
timeout 200 python tools/fwdprobe.py 16 "stream (tma 4)" > gpurun_out/r2_probe_stream16.txt 2>&1; cat gpurun_out/r2_probe_stream16.txt
