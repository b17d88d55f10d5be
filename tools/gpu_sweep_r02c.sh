#!/bin/bash
# Round-2 sweep, C2 (bunny): block sizes that keep >= 109 registers per thread with more warps per SM.
mkdir -p gpurun_out; : > gpurun_out/sweep_r02c.log
run() { echo "$*" >> gpurun_out/sweep_r02c.log; env "$@" timeout 100 python tools/profile_step.py --scene bunny_glass --bounces 16 --spp 64 --passes 2 2>&1 | tail -1 >> gpurun_out/sweep_r02c.log; }
run A=0
run RTPBR_POOL_BLOCK=192 RTPBR_POOL_MIN_BLOCKS_BUNNY=3
run RTPBR_POOL_BLOCK=288 RTPBR_POOL_MIN_BLOCKS_BUNNY=2
run RTPBR_POOL_BLOCK=160 RTPBR_POOL_MIN_BLOCKS_BUNNY=3
run RTPBR_POOL_BLOCK=128 RTPBR_POOL_MIN_BLOCKS_BUNNY=4
run RTPBR_POOL_BLOCK=96 RTPBR_POOL_MIN_BLOCKS_BUNNY=6
cat gpurun_out/sweep_r02c.log
