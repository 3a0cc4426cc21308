#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02l_sweep.txt; : > $out
run() { echo "== N=$1 ${@:2}" >> $out; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --qubits 30 --steps 2 --warmup 1 --no-cpu --no-check "${@:2}" 2>>gpurun_out/r02l_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass passes {r['passes_per_step']} nvlink {d.get('nvlink_gbs_per_gpu',0):.0f} GB/s\")
" >> $out; }
run 4
run 4 --opt peer_chunk_bits=5
run 4 --opt peer_chunk_bits=6
run 4 --opt peer_chunk_bits=7
run 4 --opt peer_chunk_bits=6 --tile-bits 12
run 4 --opt remap=0
cat $out
