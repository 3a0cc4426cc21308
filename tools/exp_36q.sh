#!/bin/bash
mkdir -p gpurun_out
timeout 42 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/run_sharded.py --workload qasm --qubits 36 --layers 4 2>gpurun_out/r02w_err.txt | grep '^{' > gpurun_out/r02w_sharded_36q_qasm_8gpu.json
cat gpurun_out/r02w_sharded_36q_qasm_8gpu.json | cut -c1-1000; tail -3 gpurun_out/r02w_err.txt | cut -c1-200
