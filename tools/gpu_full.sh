#!/bin/bash
# whole GPU suite + smoke + default bench (the round-end sequence)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full.txt 2>&1; tail -5 gpurun_out/r2_pytest_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -3 gpurun_out/r2_smoke.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cat gpurun_out/r2_bench_default.json
