#!/bin/bash
# Round-2 sweep (last): fine-tuning around the fast-region defaults after drop-outs moved to regeneration batches.
mkdir -p gpurun_out; : > gpurun_out/sweep_r02.log
run() { echo "$*" >> gpurun_out/sweep_r02.log; env "$@" timeout 40 python tools/profile_step.py --passes 3 2>&1 | tail -1 >> gpurun_out/sweep_r02.log; }
run A=0
run RTPBR_FIN_MIN=4
run RTPBR_FIN_MIN=8
run RTPBR_REGEN_MIN=16
run RTPBR_REGEN_MIN=28
run RTPBR_REGEN_MIN=32
run RTPBR_REGEN_IDLE=4
run RTPBR_REGEN_IDLE=16
run RTPBR_POOL_SLOTS=72
run RTPBR_POOL_SLOTS=88
run RTPBR_REGEN_MIN=28 RTPBR_POOL_SLOTS=88
run RTPBR_MARCH_UNROLL=2 RTPBR_FIN_MIN=1
cat gpurun_out/sweep_r02.log
