#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r$i.json 2> gpurun_out/bench_r$i.err; echo "run $i rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_r$i.json'));print(d['value'], d['roofline']['frac'], d['breakdown'])"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
for i in 4 5; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r$i.json 2> gpurun_out/bench_r$i.err; echo "run $i (with nvidia-smi polling) rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_r$i.json'));print(d['value'], d['roofline']['frac'], d['breakdown'])"
done
kill $SMI
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "full rc=$?"
cat gpurun_out/bench.json
tail -3 gpurun_out/bench.err
