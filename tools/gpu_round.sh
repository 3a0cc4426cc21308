#!/bin/bash
# one gpurun call: gpu tests + bench line + ncu launch list (+ optional full capture)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
python bench.py --steps 3 --warmup 3 --workload qft --qubits 30 --no-cpu > gpurun_out/bench_qft30.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_qft30.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/bench.err
