#!/bin/bash
# Parameter sweep of the pool kernel's resolve policy on the C1 workload.
mkdir -p gpurun_out; : > gpurun_out/sweep.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for o in 1 8 16 24 32; do for r in 4 8 16; do
  echo "other_min=$o resolve_min=$r" >> gpurun_out/sweep.log
  RTPBR_OTHER_MIN=$o RTPBR_RESOLVE_MIN=$r timeout 120 python tools/profile_step.py --passes 4 >> gpurun_out/sweep.log 2>&1
done; done
tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep.log
