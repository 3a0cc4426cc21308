#!/bin/bash
# 2-GPU run: multi-process / multi-thread parity with remap passes over NVLink + strong scaling of the 30-qubit circuit
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_process.py tests/test_multi_gpu.py -m gpu -q -x --timeout 400 > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -4 gpurun_out/r02g_pytest.log
out=gpurun_out/r02g_sweep.txt; : > $out
run() { echo "== N=$1 ${@:2}" >> $out; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --qubits 30 --steps 2 --warmup 1 --no-cpu "${@:2}" 2>>gpurun_out/r02g_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; pc=d.get('parity_check') or {}
    print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} passes {r['passes_per_step']} e2e {d['e2e']['value']:.0f} nvlink {d.get('nvlink_gbs_per_gpu',0):.0f} GB/s parity {pc.get('ok')} err {pc.get('max_abs_err')}\")
" >> $out; }
run 2
run 2 --no-check --opt double_buffer=0
run 2 --no-check --opt double_buffer=1
run 2 --no-check --tma 1
run 2 --no-check --chunk-bits 3
cat $out
