#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --durations=8 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "^FAILED|^E  |passed|failed|s call" gpurun_out/pytest_gpu.log | head -16
run() { # tb cb nbuf
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --tile-bits $1 --chunk-bits $2 --nbuf $3 $4 $5 $6 > gpurun_out/b.json 2> gpurun_out/b.err
  python -c "
import json;d=json.load(open('gpurun_out/b.json'));print('T=$1 L=$2 nbuf=$3 $4 $5 $6', round(d['value']), 'gates/s', round(d['ms_per_step']), 'ms/step', round(d['roofline']['avg_launch_ms'],3), 'ms/pass frac', round(d['roofline']['frac'],3), 'passes', d['roofline']['passes_per_step'])" | tee -a gpurun_out/exp.txt
}
run 12 7 1
run 12 4 1
run 11 4 1
run 11 4 4
run 12 4 1 --workload qft --qubits 30
run 11 4 4 --workload qft --qubits 30
run 12 7 1 --workload qft --qubits 30
