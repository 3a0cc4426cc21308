#!/bin/bash
# round-2 experiment A: parity of the rewritten tile pass + load-path / geometry sweep at 30 qubits
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
out=gpurun_out/r02a_sweep.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python bench.py --qubits 30 --steps 2 --warmup 1 --no-cpu "$@" 2>>gpurun_out/r02a_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} passes {r['passes_per_step']} e2e {d['e2e']['value']:.0f}\")
" >> $out; }
run --tma 1
run --tma 0
run --tma 1 --ctas 3
run --tma 1 --ctas 5
run --tma 0 --chunk-bits 3
run --tma 1 --tile-bits 12
run --tma 0 --tile-bits 12
run --tma 1 --tile-bits 12 --chunk-bits 5
run --tma 1 --chunk-bits 5
cat $out
for t in 1 0; do timeout 300 python tools/tile_probe.py --qubits 30 --tile-bits 11 --chunk-bits 4 --tma $t >> gpurun_out/r02a_probe.txt 2>&1; done
cat gpurun_out/r02a_probe.txt
