#!/bin/bash
# usage: tools/mgpu_cmp.sh N  -> one line per mode: n_gpus value ingest ms_per_step
N=$1
for mode in ${MODES:-scatter-dma scatter spectrum}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --mgpu-mode $mode 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$mode', d['n_gpus'], round(d['value']), round(d['value']/d['n_gpus']), d['ms_per_step'])
"
done
