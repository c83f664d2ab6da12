#!/bin/bash
# N GPUs of one box: the bench line exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1]);print(d['value'],d['ingest_msps'],d['ms_per_step'],d['e2e']['value'],d['mgpu']['modes_weak'],d['mgpu']['strong']['ingest_msps'],d['mgpu'].get('exchange_parity'))"
