#!/bin/bash
# One gpurun visit: GPU parity tests + micro benches.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/bench_match.py 2000 64 > gpurun_out/bench_match.log 2>&1
python tools/bench_match.py 4000 32 >> gpurun_out/bench_match.log 2>&1
python tools/bench_match.py 500 64 >> gpurun_out/bench_match.log 2>&1
cat gpurun_out/bench_match.log
