#!/bin/bash
# Round-2 sweep on the family-B Cornell scenes: does the fast region pay with the final scheduling (drop-outs and camera
# prologues in regeneration batches)?
mkdir -p gpurun_out; : > gpurun_out/sweep_r02b.log
run() { echo "$*" >> gpurun_out/sweep_r02b.log; env "${@:2}" timeout 60 python tools/profile_step.py --passes 3 $1 2>&1 | tail -1 >> gpurun_out/sweep_r02b.log; }
F="RTPBR_JIT_FAST=1 RTPBR_JIT_BBOX=1 RTPBR_REGEN_MIN=28 RTPBR_REGEN_IDLE=8 RTPBR_FIN_MIN=6 RTPBR_POOL_SLOTS=88 RTPBR_POOL_MIN_BLOCKS=3"
for S in "--scene cornell_box" "--scene cornell_box_v3" "--scene cornell_box_v2"; do
run "$S" A=0
run "$S" $F
run "$S" RTPBR_JIT_FAST=1 RTPBR_JIT_BBOX=1 RTPBR_REGEN_MIN=24 RTPBR_REGEN_IDLE=8 RTPBR_FIN_MIN=1 RTPBR_POOL_SLOTS=80 RTPBR_POOL_MIN_BLOCKS=3
done
cat gpurun_out/sweep_r02b.log
