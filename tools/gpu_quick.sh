#!/bin/bash
# quick gpurun call: gpu tests + smoke + two bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(round(d['value']), 'gates/s', round(d['roofline']['avg_launch_ms'],3), 'ms/pass frac', round(d['roofline']['frac'],3))"
python bench.py --steps 3 --warmup 3 --workload qft --qubits 30 --no-cpu "$@" > gpurun_out/bench_qft30.json 2>> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_qft30.json'));print('qft30', round(d['value']), 'gates/s', round(d['roofline']['avg_launch_ms'],3), 'ms/pass frac', round(d['roofline']['frac'],3))"
tail -3 gpurun_out/bench.err
