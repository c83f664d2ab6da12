#!/bin/bash
mkdir -p gpurun_out
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  ok=0; bad=0
  for i in 1 2 3 4; do
    if env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err; then ok=$((ok+1)); else bad=$((bad+1)); fi
  done
  echo "$name: ok=$ok bad=$bad"
}
run default "X=1"
run tma1 "B200_OPTS=6=1"
run fuse2 "B200_OPTS=5=2"
run banks1 "X=1" --banks 1
run batch16 "X=1" --batch 16
timeout 500 compute-sanitizer --tool memcheck --print-limit 3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/memcheck.txt 2>&1
grep -v "^\[W" gpurun_out/memcheck.txt | head -60
