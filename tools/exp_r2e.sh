#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -8 gpurun_out/r02e_pytest.log
out=gpurun_out/r02e_sweep.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python bench.py --qubits 30 --steps 2 --warmup 1 --no-cpu --no-check "$@" 2>>gpurun_out/r02e_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} passes {r['passes_per_step']} e2e {d['e2e']['value']:.0f}\")
" >> $out; }
run --tma 0 --opt tile_ctas=4 prefetch=0
run --tma 0 --opt tile_ctas=4 prefetch=1
run --tma 0 --opt tile_ctas=3 prefetch=0
run --tma 0 --opt tile_ctas=3 prefetch=1
run --tma 1 --opt tile_ctas=3 prefetch=1
run --tma 0 --opt tile_ctas=3 prefetch=1 --chunk-bits 3
run --tma 0 --opt tile_ctas=3 prefetch=1 --tile-bits 12
cat $out
