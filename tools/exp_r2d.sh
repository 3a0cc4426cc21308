#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -6 gpurun_out/r02d_pytest.log
out=gpurun_out/r02d_sweep.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python bench.py --qubits 30 --steps 2 --warmup 1 --no-cpu --no-check "$@" 2>>gpurun_out/r02d_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.0f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} passes {r['passes_per_step']} e2e {d['e2e']['value']:.0f}\")
" >> $out; }
run --tma 0
run --tma 1
run --tma 0 --tile-bits 12
run --tma 0 --opt tile_ctas=3
cat $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 6 -c 2 -o gpurun_out/r02d_tile_ptx -f \
     python bench.py --qubits 30 --steps 1 --warmup 1 --no-cpu --no-check --depth 6 --tma 0 > gpurun_out/r02d_ncu.log 2>&1
timeout 300 python tools/tile_probe.py --qubits 30 --tile-bits 11 --chunk-bits 4 --tma 0 > gpurun_out/r02d_probe.txt 2>&1; cat gpurun_out/r02d_probe.txt
