#!/bin/bash
# round-2 call 1: microbenchmarks, the GPU test-suite as it stands, forward variants that never ran on hardware
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_nvsmi.txt 2>&1
timeout 300 tools/ubench/ubench > gpurun_out/r2_ubench.txt 2>&1; echo "ubench rc=$?"; cat gpurun_out/r2_ubench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu0.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu0.txt
timeout 300 python tools/fwdprobe.py 64 "scalar" "tma3 fused lag1" > gpurun_out/r2_probe_a.txt 2>&1; cat gpurun_out/r2_probe_a.txt
timeout 200 python tools/fwdprobe.py 64 "table" > gpurun_out/r2_probe_table.txt 2>&1; cat gpurun_out/r2_probe_table.txt
timeout 200 python tools/fwdprobe.py 64 "fused12" > gpurun_out/r2_probe_fused12.txt 2>&1; cat gpurun_out/r2_probe_fused12.txt
timeout 200 python tools/fwdprobe.py 16 "fused12" "sub-batches" "lanes" > gpurun_out/r2_probe_16.txt 2>&1; cat gpurun_out/r2_probe_16.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo "bench rc=$?"; cat gpurun_out/r2_bench0.json
