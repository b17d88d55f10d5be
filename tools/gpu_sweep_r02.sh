#!/bin/bash
# Round-2 sweep (last): fine-tuning around the fast-region defaults (regeneration 24 / idle 8 / finish 6 / 3 CTAs x 80 slots).
mkdir -p gpurun_out; : > gpurun_out/sweep_r02.log
run() { echo "$*" >> gpurun_out/sweep_r02.log; env "$@" timeout 40 python tools/profile_step.py --passes 3 2>&1 | tail -1 >> gpurun_out/sweep_r02.log; }
run A=0
run RTPBR_RESOLVE_MIN=12
run RTPBR_RESOLVE_MIN=16
run RTPBR_RESOLVE_MIN=24
run RTPBR_FIN_MIN=8
run RTPBR_FIN_MIN=5
run RTPBR_REGEN_MIN=20
run RTPBR_REGEN_MIN=28
run RTPBR_REGEN_IDLE=12
run RTPBR_POOL_SLOTS=72
run RTPBR_POOL_SLOTS=88
run RTPBR_POOL_SLOTS=96 RTPBR_REGEN_MIN=32
run RTPBR_RESOLVE_MIN=16 RTPBR_FIN_MIN=8
cat gpurun_out/sweep_r02.log
