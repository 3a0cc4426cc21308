#!/bin/bash
# norm check of the random circuit: bash tools/dbg.sh N qubits layers [opts...]
N=$1; Q=$2; LY=$3; shift 3
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/run_sharded.py --workload random --qubits $Q --layers $LY --opt "$@" 2>gpurun_out/rs.err | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N q=$Q layers=$LY $*', 'passes', d['hbm_passes'], 'ms', round(d['apply_ms']), 'norm-1 = %.3e' % (d['checks']['norm_sqr_before_measure']-1))"
