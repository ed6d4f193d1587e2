#!/bin/bash
# Multi-GPU visit (gpurun --gpus N -- bash tools/gpu_multi.sh N): device-side multi-GPU BA tests + the N-rank bench line.
N=${1:-2}
mkdir -p gpurun_out
[ -x tools/micro/minmax_pipes ] && tools/micro/minmax_pipes > gpurun_out/minmax_pipes.txt 2>&1; cat gpurun_out/minmax_pipes.txt
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_ba_multi_gpu.py tests/test_ba_sharded_gpu.py -x -q -m gpu > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_${N}gpu.log
tail -15 gpurun_out/pytest_multi_${N}gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err; tail -c 6000 gpurun_out/bench_${N}gpu.json
exit 0
