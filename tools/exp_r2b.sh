#!/bin/bash
# round-2 experiment B: full GPU parity run + ncu --set full of the rewritten tile pass (bulk and cp.async loads)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
for t in 1 0; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 6 -c 2 -o gpurun_out/r02b_tile_tma$t -f \
     python bench.py --qubits 30 --steps 1 --warmup 1 --no-cpu --no-check --depth 6 --tma $t > gpurun_out/r02b_ncu_tma$t.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
