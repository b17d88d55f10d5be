#!/bin/bash
# Packed-form knobs of the specialised march step (Cornell, 64 spp).
mkdir -p gpurun_out
for pc in 0 1 2 3; do
  echo -n "pack_clamps=$pc: "
  RTPBR_PACK_CLAMPS=$pc timeout 300 python tools/profile_step.py --passes 4 --spp 64 2>&1 | tail -1
done | tee gpurun_out/sweep_pairs.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
