#!/bin/bash
# Experiment: path regeneration in its own batches (RTPBR_REGEN_MIN = slots that must wait before a regeneration batch runs).
mkdir -p gpurun_out
RTPBR_REGEN_MIN=16 timeout 150 python -m pytest tests -m gpu -x -q -k "c0 or ragged or golden or progressive or chunking or shards or jit_and_aot" 2>&1 | tail -2
for n in 0 8 16 24; do
  echo -n "regen_min=$n: "
  RTPBR_REGEN_MIN=$n timeout 100 python tools/profile_step.py --passes 4 --spp 64 2>&1 | tail -1
done | tee gpurun_out/sweep_regen.log
