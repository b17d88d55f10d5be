#!/bin/bash
# What makes the per-rank kernel 5.5 % slower once torch.distributed(NCCL, 2 ranks) is up?  NCCL transport knobs.
mkdir -p gpurun_out
: > gpurun_out/multi_probe2.log
run() {
  echo "== $*" >> gpurun_out/multi_probe2.log
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_probe2.py 2>&1 | grep "^rank 0" | head -2 >> gpurun_out/multi_probe2.log
}
run NCCL_DEBUG=WARN
run NCCL_P2P_DISABLE=1
run NCCL_NVLS_ENABLE=0
run NCCL_CUMEM_ENABLE=0
run NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1
cat gpurun_out/multi_probe2.log
