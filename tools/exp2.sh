#!/bin/bash
# 2-GPU sweep of the peer-tile geometry
mkdir -p gpurun_out
: > gpurun_out/exp2.txt
run() {
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu "$@" > gpurun_out/b2.json 2> gpurun_out/b2.err
  python -c "
import json,sys;d=json.loads([l for l in open('gpurun_out/b2.json') if l.startswith('{')][-1]);print(' '.join(sys.argv[1:]), round(d['value']), 'gates/s', round(d['ms_per_step']), 'ms/step', round(d['roofline']['avg_launch_ms'],3), 'ms/pass', d['roofline']['passes_per_step'], 'passes', d.get('peer',''))" "$@" | tee -a gpurun_out/exp2.txt
}
run --chunk-bits 4
run --chunk-bits 6
run --chunk-bits 7
run --tile-bits 12 --chunk-bits 7
run --tile-bits 12 --chunk-bits 5
