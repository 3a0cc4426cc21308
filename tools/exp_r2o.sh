#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -4 gpurun_out/r02o_pytest.log
out=gpurun_out/r02o_sweep.txt; : > $out
run() { echo "== $*" >> $out; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-check "$@" 2>>gpurun_out/r02o_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; c=d.get('clocks') or {}
    print(f\"{d['value']:.0f} gates/s {d['ms_per_step']:.3f} ms/step {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} passes {r['passes_per_step']} e2e {d['e2e']['value']:.0f} power {c.get('power_w_max')} {c.get('reasons')}\")
" >> $out; }
run
run --workload qft --qubits 30
run --opt ptx_ops=0
cat $out
