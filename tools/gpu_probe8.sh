#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clients.py tests/test_gpu_stream_blocks.py tests/test_golden.py tests/test_gpu_bench_config.py::test_cfg2_iq_2p20_1024_clients_batch64 tests/test_gpu_host_driver.py -x -q -m gpu > gpurun_out/r2_pytest_tail3.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_tail3.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_tail3.json 2> gpurun_out/r2_bench_tail3.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_tail3.json'));print(d['value'],d['ms_per_step'],d['breakdown'],d['e2e']['value'])"; tail -3 gpurun_out/r2_bench_tail3.err
