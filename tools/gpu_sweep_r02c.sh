#!/bin/bash
# Round-2 sweep, C2 (bunny): CTAs per SM / block size / slots.
mkdir -p gpurun_out; : > gpurun_out/sweep_r02c.log
run() { echo "$*" >> gpurun_out/sweep_r02c.log; env "$@" timeout 100 python tools/profile_step.py --scene bunny_glass --bounces 16 --spp 64 --passes 2 2>&1 | tail -1 >> gpurun_out/sweep_r02c.log; }
run A=0
run RTPBR_POOL_MIN_BLOCKS_BUNNY=3
run RTPBR_POOL_BLOCK=192 RTPBR_POOL_MIN_BLOCKS_BUNNY=4
run RTPBR_POOL_BLOCK=384 RTPBR_POOL_MIN_BLOCKS_BUNNY=2
run RTPBR_POOL_BLOCK=320 RTPBR_POOL_MIN_BLOCKS_BUNNY=2
run RTPBR_POOL_BLOCK=128 RTPBR_POOL_MIN_BLOCKS_BUNNY=5
run RTPBR_POOL_SLOTS=96
run RTPBR_POOL_SLOTS=48
run RTPBR_RESOLVE_MIN=16
cat gpurun_out/sweep_r02c.log
