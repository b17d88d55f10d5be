#!/bin/bash
# Tuning knobs of the specialised march loop (Cornell, 64 spp).
mkdir -p gpurun_out
for n in 1 2; do
  echo -n "march_unroll=$n: "
  RTPBR_MARCH_UNROLL=$n timeout 300 python tools/profile_step.py --passes 4 --spp 64 2>&1 | tail -1
done | tee gpurun_out/sweep_pairs.log
RTPBR_MARCH_UNROLL=2 timeout 600 python -m pytest tests -m gpu -x -q -k "c0 or ragged or jit or golden" 2>&1 | tail -2
