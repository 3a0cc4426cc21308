#!/bin/bash
# one gpurun call: gpu tests + bench lines + ncu launch list + full capture of the tile pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
python bench.py --steps 3 --warmup 3 --workload qft --qubits 30 --no-cpu > gpurun_out/bench_qft30.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_qft30.json
python bench.py --steps 3 --warmup 3 --no-fuse --no-cpu --depth 10 > gpurun_out/bench_unfused.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_unfused.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 4 -c 2 -f -o gpurun_out/tile_full python bench.py --steps 1 --warmup 1 --no-cpu --depth 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/bench.err
ls -la gpurun_out | head -30
