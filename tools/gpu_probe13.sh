#!/bin/bash
# step time against forward grid sizing while the tail kernel holds SMs
mkdir -p gpurun_out
for o in "" "21=116" "21=112" "21=100" "22=4" "22=8" "22=16" "21=116,22=8" "21=116,22=16" "21=116,22=32" "21=112,22=16" "21=116,22=16,19=0"; do
  echo "== B200_OPTS=$o"
  B200_OPTS=$o timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d.get('breakdown', {}))[:400])
"
done > gpurun_out/probe13.txt 2>&1
python tools/fwdprobe.py 64 zzz >> gpurun_out/probe13.txt 2>&1
for s in 4 8 16 32; do echo "== split $s"; B200_OPTS=22=$s python tools/fwdprobe.py 64 zzz; done >> gpurun_out/probe13.txt 2>&1
cat gpurun_out/probe13.txt
