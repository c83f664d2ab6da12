#!/bin/bash
# One round of GPU evidence for profiles/: GPU tests, the bench line, the ncu launch list of the same command and one
# ncu --set full capture of the forward kernels. Run as: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:fft_|pyramid|client_|radix' -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:fft_pass|pyramid_kernel' -s 3 -c 3 \
    -f -o gpurun_out/fwd python tools/fwdonce.py 64 2 > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log
