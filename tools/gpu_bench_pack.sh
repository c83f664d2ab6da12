#!/bin/bash
# round-2 bench lines (one JSON line each)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; echo "$name rc=$? $(head -c 300 gpurun_out/r2_bench_$name.json)"; }
run reference --impl reference
run cfg2
run cufft --impl cufft
run cfg3 --config cfg3
run cfg1 --config cfg1
run cadence_pcm16 --waterfall-skip 0 --pcm16 --no-cpu-baseline
run raw_s16 --e2e-raw s16 --waterfall-skip 0 --pcm16 --no-cpu-baseline
run raw_u8 --e2e-raw u8 --waterfall-skip 0 --pcm16 --no-cpu-baseline
