#!/bin/bash
# round-2 profile pack: launch list of the bench command, one --set full capture of every hot kernel (exported to CSV on
# the box: the reports themselves are ~20 MB each and gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
KRE='regex:fft_pass|pyramid|client_|radix_split|waterfall_gather|flag_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench.log
cap() {  # name, kernel regex, skip, count, command...
  name=$1; kre=$2; skip=$3; cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "$kre" -s $skip -c $cnt -f -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/${name}_ncu.log
}
cap r2_fwd 'regex:fft_pass1_tma|fft_pass2_tma3|pyramid_kernel' 6 3 python tools/fwd_once.py 4
cap r2_fwd_r2c 'regex:fft_pass1_tma|fft_pass2_tma3|pyramid_kernel' 6 3 python tools/fwd_once.py 4 real
cap r2_demod360 'regex:client_demod_warp' 30 1 python tools/cliprobe.py 1024 64
cap r2_tail2 'regex:client_tail2' 4 1 python tools/cliprobe.py 1024 64
ls -la gpurun_out/
