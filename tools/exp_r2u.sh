#!/bin/bash
# 8-GPU run: multi-process parity at P = 2/4/8 over NVLink, strong scaling of the 30-qubit circuit, configs[3]/[4] at full size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_process.py -m gpu -q -x --timeout 300 > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
tail -3 gpurun_out/r02u_pytest.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 3 --warmup 2 --no-cpu "${@:2}" 2>>gpurun_out/r02u_err.txt | grep '^{' ; }
run 8 > gpurun_out/r02u_bench_n8_30q.json
run 4 > gpurun_out/r02u_bench_n4_30q.json
for f in gpurun_out/r02u_bench_n*_30q.json; do python -c "
import sys,json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']; pc=d.get('parity_check') or {}
print('$f', f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass passes {r['passes_per_step']} nvlink {d.get('nvlink_gbs_per_gpu',0):.0f} GB/s parity {pc.get('ok')} err {pc.get('max_abs_err')}\")
"; done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/run_sharded.py --workload qasm --qubits 36 --layers 4 2>>gpurun_out/r02u_err.txt | grep '^{' > gpurun_out/r02u_sharded_36q_qasm_8gpu.json
cat gpurun_out/r02u_sharded_36q_qasm_8gpu.json | cut -c1-900
tail -5 gpurun_out/r02u_err.txt | cut -c1-300
