#!/bin/bash
# Round-2 sweep on the family-B scenes without the fast region: finish threshold / regeneration batches / t_stop.
mkdir -p gpurun_out; : > gpurun_out/sweep_r02b.log
run() { echo "$*" >> gpurun_out/sweep_r02b.log; env "${@:2}" timeout 60 python tools/profile_step.py --passes 3 $1 2>&1 | tail -1 >> gpurun_out/sweep_r02b.log; }
for S in "--scene cornell_box" "--scene cornell_box_v3" "--scene tokyo_ibl"; do
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0 RTPBR_FIN_MIN=4
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0 RTPBR_FIN_MIN=6
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0 RTPBR_FIN_MIN=4 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
run "$S" RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0 RTPBR_FIN_MIN=4 RTPBR_POOL_SLOTS=80 RTPBR_POOL_MIN_BLOCKS=3
done
run "--scene tokyo_ibl" RTPBR_FIN_MIN=4
run "--scene tokyo_ibl" A=0
B="--scene bunny_glass --spp 32 --bounces 16"
run "$B" A=0
run "$B" RTPBR_JIT_BBOX=0
run "$B" RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
cat gpurun_out/sweep_r02b.log
