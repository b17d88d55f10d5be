#!/bin/bash
# Packed-form / pipe-balance knobs of the specialised march step (Cornell, 64 spp).
mkdir -p gpurun_out
for n in 0 1 2; do
  echo -n "alu_clamp_pairs=$n: "
  RTPBR_ALU_CLAMPS=$n timeout 300 python tools/profile_step.py --passes 4 --spp 64 2>&1 | tail -1
done | tee gpurun_out/sweep_pairs.log
RTPBR_ALU_CLAMPS=1 timeout 600 python -m pytest tests -m gpu -x -q -k "c0 or ragged or jit" 2>&1 | tail -2
