#!/bin/bash
# e2e with / without the ramped chunk schedule
for r in 0 1; do
  if [ $r == 1 ]; then export VSLAM_FRONT_RAMP=1; fi
  python tools/e2e_trace.py 256 2>&1 | grep "call ms" | tail -2
done
