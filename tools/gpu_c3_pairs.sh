#!/bin/bash
# C3 (tokyo_ibl) with and without packed sphere pairs, after the parity tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for sp in 1 0; do echo -n "sphere_pairs=$sp: "; RTPBR_SPHERE_PAIRS=$sp timeout 600 python bench.py --workload c3 --steps 3 2>/dev/null | grep -o '"value": [0-9.]*'; done
