#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
for b in 16 64; do timeout 200 python tools/fwdprobe2.py $b > gpurun_out/fwdprobe2_$b.txt 2>&1; cat gpurun_out/fwdprobe2_$b.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:fft_|pyramid|client_' -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/launches.csv
