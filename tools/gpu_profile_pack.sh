#!/bin/bash
# round-2 profile pack: launch list of the bench command, one --set full capture of every hot kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
tail -c 600 gpurun_out/r2_launches_bench.log
# forward group: one launch of each kernel at 64 frames per launch (the bench configuration)
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fft_pass1_tma|fft_pass2_tma3|pyramid_kernel' -s 6 -c 3 -f \
  -o gpurun_out/r2_fwd python tools/fwdprobe.py 64 zzz > gpurun_out/r2_ncu_fwd.log 2>&1; tail -2 gpurun_out/r2_ncu_fwd.log
# clients: chunked demodulation (360-point plan) and the tail pipeline, 1024 clients x 64 frames
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:client_demod_warp|client_tail2' -s 40 -c 1 -f \
  -o gpurun_out/r2_demod360 python tools/cliprobe.py 1024 64 > gpurun_out/r2_ncu_cli.log 2>&1; tail -2 gpurun_out/r2_ncu_cli.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:client_tail2' -s 4 -c 1 -f \
  -o gpurun_out/r2_tail2b python tools/cliprobe.py 1024 64 > gpurun_out/r2_ncu_tail.log 2>&1; tail -2 gpurun_out/r2_ncu_tail.log
ls -la gpurun_out/*.ncu-rep
