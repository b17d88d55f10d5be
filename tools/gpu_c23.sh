set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c2 --steps 2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c3 --steps 3 > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c3.json
