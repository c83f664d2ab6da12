#!/bin/bash
mkdir -p gpurun_out
ok=0; bad=0
for i in 1 2 3 4 5 6 7 8; do
  if timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_default.json 2> gpurun_out/b_default.err; then ok=$((ok+1)); else bad=$((bad+1)); fi
done
echo "default: ok=$ok bad=$bad"
timeout 300 python -m pytest tests -m gpu -x -q -k "forward or golden" > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "full rc=$?"
cat gpurun_out/bench.json
grep -v "^\[W" gpurun_out/bench.err | tail -3
