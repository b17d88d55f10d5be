#!/bin/bash
# Round-2 sweep on tokyo_ibl at 3 CTAs/SM: scheduling knobs again.
mkdir -p gpurun_out; : > gpurun_out/sweep_r02b.log
run() { echo "$*" >> gpurun_out/sweep_r02b.log; env "${@:2}" timeout 60 python tools/profile_step.py --passes 3 $1 2>&1 | tail -1 >> gpurun_out/sweep_r02b.log; }
S="--scene tokyo_ibl"
run "$S" A=0
run "$S" RTPBR_FIN_MIN=3
run "$S" RTPBR_FIN_MIN=5
run "$S" RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
run "$S" RTPBR_REGEN_MIN=24 RTPBR_REGEN_IDLE=8 RTPBR_FIN_MIN=4
run "$S" RTPBR_POOL_SLOTS=72
run "$S" RTPBR_POOL_SLOTS=56
run "$S" RTPBR_RESOLVE_MIN=16
run "$S" RTPBR_JIT_BBOX=0
cat gpurun_out/sweep_r02b.log
