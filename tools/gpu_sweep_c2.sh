#!/bin/bash
# Resolve-threshold / pool-size sweep for the neural-bunny kernels (C2 at 64 spp).
mkdir -p gpurun_out
for rm in 8 4 2 1; do for slots in 64 96; do
  echo -n "resolve_min=$rm slots=$slots: "
  RTPBR_RESOLVE_MIN=$rm RTPBR_POOL_SLOTS=$slots timeout 300 python tools/profile_step.py --scene bunny_glass --bounces 16 --passes 3 --spp 64 2>&1 | tail -1
done; done | tee gpurun_out/sweep_c2.log
