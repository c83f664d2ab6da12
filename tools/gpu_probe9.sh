#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_a.err; cat gpurun_out/r2_bench_a.json
timeout 300 python bench.py --impl cufft; timeout 300 python bench.py --impl cufft --config cfg3
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 --waterfall-skip 0 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/r2_bench_cfg3.err; cat gpurun_out/r2_bench_cfg3.json
timeout 600 python bench.py --steps 10 --warmup 3 --waterfall-skip 0 --pcm16 --e2e-raw s16 --no-cpu-baseline > gpurun_out/r2_bench_n3.json 2> gpurun_out/r2_bench_n3.err; echo "n3 rc=$?"; tail -3 gpurun_out/r2_bench_n3.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n3.json'));print('N3/N2:',d['value'],d['e2e'],d.get('e2e_raw'))"
