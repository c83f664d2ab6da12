#!/bin/bash
mkdir -p gpurun_out
KRE='regex:fft_pass|pyramid|client_|radix_split|waterfall_gather|flag_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fft_pass1_tma|fft_pass2_tma3|pyramid_kernel' -s 6 -c 3 -f \
  -o gpurun_out/r2_fwd python tools/fwd_once.py 4 > gpurun_out/r2_ncu_fwd.log 2>&1; tail -2 gpurun_out/r2_ncu_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fft_pass1_tma|fft_pass2_tma3|pyramid_kernel' -s 6 -c 3 -f \
  -o gpurun_out/r2_fwd_r2c python tools/fwd_once.py 4 real > gpurun_out/r2_ncu_fwd_r2c.log 2>&1; tail -2 gpurun_out/r2_ncu_fwd_r2c.log
