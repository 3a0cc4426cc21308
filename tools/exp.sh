#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp.txt
run() { # tb cb nbuf
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --tile-bits $1 --chunk-bits $2 --nbuf $3 > gpurun_out/b.json 2> gpurun_out/b.err
  python -c "
import json;d=json.load(open('gpurun_out/b.json'));print('T=$1 L=$2 nbuf=$3', round(d['value']), 'gates/s', round(d['ms_per_step']), 'ms/step', round(d['roofline']['avg_launch_ms'],3), 'ms/pass frac', round(d['roofline']['frac'],3), 'passes', d['roofline']['passes_per_step'])" | tee -a gpurun_out/exp.txt
}
run 11 4 0
run 12 4 1
run 12 5 1
run 10 3 0
run 10 4 0
run 11 3 0
run 9 2 0
