#!/bin/bash
# A/B on ONE box: the previous commit's library (build_tmp/wt, A) against the current one (B), interleaved
# A = the previous commit built in a worktree:  git worktree add build_tmp/wt <commit> && make -C build_tmp/wt/qvnt_b200/csrc
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log; tail -3 gpurun_out/r02s_pytest.log
timeout 300 python tools/bisect_mix.py > gpurun_out/r02s_bisect.txt 2>&1; cat gpurun_out/r02s_bisect.txt
out=gpurun_out/r02s_ab.txt; : > $out
run() { echo "== $1 ${@:2}" >> $out; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-check "${@:2}" 2>>gpurun_out/r02s_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; c=d.get('clocks') or {}
    print(f\"{d['value']:.0f} gates/s {r['avg_launch_ms']:.3f} ms/pass frac {r['frac']:.3f} power {c.get('power_w_max')} {c.get('reasons')}\")
" >> $out; }
A=$PWD/build_tmp/wt/qvnt_b200/libqvnt_b200.so
QVNT_B200_LIB=$A run A
run B
run B --workload qft --qubits 30
cat $out
