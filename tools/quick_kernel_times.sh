#!/bin/bash
# quick per-kernel times of the headline workload (bench.py without its secondary blocks):
#   gpurun -- bash tools/quick_kernel_times.sh
timeout 500 python bench.py --steps 10 --warmup 3 --no-ba --no-sgbm --no-cfg3 --no-street --cpu-seconds 1 > gpurun_out/bench_kt.json 2> gpurun_out/bench_kt.err; tail -3 gpurun_out/bench_kt.err
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_kt.json"))
print("value", round(d["value"]), "e2e", d["e2e"].get("one_call_at_a_time"), d["e2e"].get("two_batches_in_flight"), "ms/step", round(d["ms_per_step"],3))
print(d["roofline"]["kernel_ms_per_step"]); print(d["clocks"])
EOF
