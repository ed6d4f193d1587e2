#!/bin/bash
# One gpurun visit: GPU parity tests + smoke + bench (+ optional ncu launch list).  Output lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/extra.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --pairs 32 --cpu-seconds 1 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?"; tail -3 gpurun_out/launches.csv
fi
if [ "$2" == "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fast_kernel|blur_kernel|describe_kernel|hamming_argmin|harris_select|resize_level|triangulate|crosscheck" -s 15 -c 15 -f -o gpurun_out/full python tools/profile_frontend.py 32 2 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sgbm|speckle" -s 9 -c 9 -f -o gpurun_out/sgbm_full python tools/profile_sgbm.py 8 2 > gpurun_out/ncu_sgbm.log 2>&1
  echo "ncu sgbm rc=$?"; tail -2 gpurun_out/ncu_sgbm.log; ls -la gpurun_out/*.ncu-rep
fi
if [ "$3" == "ba" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ba_lm_kernel|pnp_|anms_points|ba_dense" -c 14 -f -o gpurun_out/ba_pnp_full python tools/profile_ba_pnp.py > gpurun_out/ncu_ba_pnp.log 2>&1
  echo "ncu ba/pnp rc=$?"; tail -4 gpurun_out/ncu_ba_pnp.log
fi
exit 0
