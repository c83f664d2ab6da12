#!/bin/bash
# diagnostic pass over the current build: stage timings, ncu full capture of the forward kernels, bench line, launch list
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
for b in 16 4 2; do timeout 200 python tools/fwdprobe.py $b > gpurun_out/fwdprobe$b.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:fft_pass|pyramid_kernel' -s 6 -c 3 \
    -f -o gpurun_out/fwd python tools/fwdonce.py 16 3 > gpurun_out/ncu_fwd.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/fwdprobe16.txt gpurun_out/fwdprobe4.txt gpurun_out/fwdprobe2.txt
cat gpurun_out/bench.json
