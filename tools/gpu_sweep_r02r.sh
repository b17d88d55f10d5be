#!/bin/bash
# Round 2: regeneration batches (now also the place where missed rays end) for the PBR kernels with a sky.
mkdir -p gpurun_out
L=gpurun_out/sweep_r02r.log
: > $L
run() {   # scene spp bounces env...
  echo "== $1 spp=$2 b=$3 ${*:4}" >> $L
  env "${@:4}" timeout 100 python tools/profile_step.py --scene $1 --spp $2 --bounces $3 --passes 3 2>&1 | tail -1 >> $L
}
run tokyo_ibl 64 8 A=0
run tokyo_ibl 64 8 RTPBR_REGEN_MIN=28 RTPBR_REGEN_IDLE=12
run tokyo_ibl 64 8 RTPBR_REGEN_MIN=28 RTPBR_REGEN_IDLE=12 RTPBR_POOL_SLOTS=88
run tokyo_ibl 64 8 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
run tokyo_ibl 64 8 RTPBR_REGEN_MIN=28 RTPBR_REGEN_IDLE=12 RTPBR_POOL_SLOTS=88 RTPBR_FIN_MIN=4
run bunny_glass 32 16 RTPBR_REGEN_MIN=28 RTPBR_REGEN_IDLE=12
cat $L
