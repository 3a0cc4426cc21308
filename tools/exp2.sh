#!/bin/bash
# multi-GPU sweep of the tile geometry: bash tools/exp2.sh N
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/exp2.txt
run() {
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 2 --no-cpu "$@" > gpurun_out/b2.json 2> gpurun_out/b2.err
  python -c "
import json,sys;d=json.loads([l for l in open('gpurun_out/b2.json') if l.startswith('{')][-1]);print('N=$N', ' '.join(sys.argv[1:]), round(d['value']), 'gates/s', round(d['ms_per_step']), 'ms/step', round(d['roofline']['avg_launch_ms'],3), 'ms/pass', d['roofline']['passes_per_step'], 'passes')" "$@" | tee -a gpurun_out/exp2.txt
}
run --tile-bits 11 --chunk-bits 4
run --tile-bits 11 --chunk-bits 6
run --tile-bits 12 --chunk-bits 5
run --tile-bits 12 --chunk-bits 7
run --tile-bits 12 --chunk-bits 4
run --tile-bits 10 --chunk-bits 3
