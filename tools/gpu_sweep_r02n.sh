#!/bin/bash
# Round 2, last sweep: leftovers after the configuration literals (each run under its own short timeout).
mkdir -p gpurun_out
L=gpurun_out/sweep_r02n.log
: > $L
run() {   # scene spp bounces env...
  echo "== $1 spp=$2 b=$3 ${*:4}" >> $L
  env "${@:4}" timeout 120 python tools/profile_step.py --scene $1 --spp $2 --bounces $3 --passes 3 2>&1 | tail -1 >> $L
}
run cornell_box 64 8 A=0
run cornell_box 64 8 RTPBR_JIT_FAST=1
run cornell_box_v2 64 8 A=0
run cornell_box_v2 64 8 RTPBR_JIT_FAST=1
run bunny_glass 32 16 A=0
run bunny_glass 32 16 RTPBR_RESOLVE_OOL=1
run cornell_box_shortest 64 8 A=0
run cornell_box_shortest 64 8 RTPBR_RESOLVE_OOL=1
run tokyo_ibl 64 8 A=0
run tokyo_ibl 64 8 RTPBR_POOL_MIN_BLOCKS=4
run tokyo_ibl 64 8 RTPBR_POOL_SLOTS=72
run tokyo_ibl 64 8 RTPBR_POOL_SLOTS=80
run tokyo_ibl 64 8 RTPBR_FIN_MIN=3
run tokyo_ibl 64 8 RTPBR_RESOLVE_OOL=0
run cornell_box_v3 64 8 A=0
run cornell_box_v3 64 8 RTPBR_POOL_SLOTS=80
run src_scene 64 8 A=0
run src_scene 64 8 RTPBR_POOL_MIN_BLOCKS=4
cat $L
