#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_process.py tests/test_multi_gpu.py tests/test_qasm.py -m gpu -q -x --timeout 400 > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -4 gpurun_out/r02r_pytest.log
out=gpurun_out/r02r_sweep.txt; : > $out
run() { echo "== N=$1 ${@:2}" >> $out; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 2 --warmup 1 --no-cpu "${@:2}" 2>>gpurun_out/r02r_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; pc=d.get('parity_check') or {}
    print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass passes {r['passes_per_step']} nvlink {d.get('nvlink_gbs_per_gpu',0):.0f} GB/s parity {pc.get('ok')} err {pc.get('max_abs_err')}\")
" >> $out; }
run 2
run 2 --no-check --opt peer_tile_bits=11
run 2 --no-check --opt peer_tile_bits=11 double_buffer=2
run 2 --no-check --opt peer_tile_bits=11 peer_chunk_bits=5
run1() { echo "== N=1 $*" >> $out; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-check "$@" 2>>gpurun_out/r02r_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.3f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f}\")
" >> $out; }
run1
run1 --workload qft --qubits 30
cat $out
