#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config.py::test_cfg3_real_2p21_256_fm_clients_batch64 tests/test_gpu_stream_blocks.py tests/test_gpu_waterfall.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2_pytest_r2c.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_r2c.txt
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3b.json 2> gpurun_out/r2_bench_cfg3b.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/r2_bench_cfg3b.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_cfg3b.json'));print(d['value'],d['ms_per_step'],d['breakdown'],d['e2e']['value'],d['roofline']['frac'])"
