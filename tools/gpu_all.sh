#!/bin/bash
# Parity tests, smoke, headline bench and the informational C2 / C3 workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --workload c3 --steps 3 > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --workload c2 --steps 2 > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_c3.json; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench.err
