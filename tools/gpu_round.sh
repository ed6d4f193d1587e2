#!/bin/bash
# One gpurun visit: GPU parity tests + micro benches.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for s in "$@"; do timeout 600 python $s >> gpurun_out/extra.log 2>&1; done
[ -f gpurun_out/extra.log ] && tail -40 gpurun_out/extra.log
