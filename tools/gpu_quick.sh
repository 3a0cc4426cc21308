#!/bin/bash
# quick gpurun call: gpu tests + two bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
python bench.py --steps 3 --warmup 3 --workload qft --qubits 30 --no-cpu > gpurun_out/bench_qft30.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_qft30.json
tail -3 gpurun_out/bench.err
