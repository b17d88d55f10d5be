#!/bin/bash
# Multi-GPU check: parity of the sharded render + NCCL tile reduce, then the scaling bench at N ranks.
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 3 > gpurun_out/bench_ref_$N.json 2>> gpurun_out/bench_$N.err
tail -n 5 gpurun_out/multi_check_$N.log; cat gpurun_out/bench_$N.json; tail -n 5 gpurun_out/bench_$N.err; cat gpurun_out/bench_ref_$N.json
