#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest_multi.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2_bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ingest_msps'],d['ms_per_step'],d['e2e'],d.get('mgpu'))"
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
