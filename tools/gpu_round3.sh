#!/bin/bash
# Final round-2 evidence on ONE B200 for the kernels that changed after tools/gpu_round2.sh ran: bench lines, the ncu
# launch list of the bench command, ncu --set full of the tile pass (random circuit and QFT), tile probes, timeline.
# (The reference arm and the other kernels are unchanged: their r02_* files from gpu_round2.sh stand.)
mkdir -p gpurun_out; o=gpurun_out/r02_final
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 5 --warmup 3 > ${o}_bench_30q.json 2> ${o}_bench.err
python bench.py --workload qft20 --steps 20 --warmup 5 > ${o}_bench_qft20.json 2>> ${o}_bench.err
python bench.py --workload qft --qubits 30 --steps 5 --warmup 3 --no-cpu > ${o}_bench_qft30.json 2>> ${o}_bench.err
python bench.py --workload qft_h --qubits 32 --steps 3 --warmup 3 --no-cpu --no-check > ${o}_bench_qft_h_32q.json 2>> ${o}_bench.err
python bench.py --qubits 28 --steps 5 --warmup 3 --no-cpu --no-check > ${o}_bench_28q.json 2>> ${o}_bench.err
python bench.py --steps 3 --warmup 2 --no-cpu --no-check --opt ptx_ops=0 > ${o}_bench_30q_cpp_ops.json 2>> ${o}_bench.err
timeout 300 python tools/tile_probe.py --qubits 30 --tile-bits 11 --chunk-bits 4 --tma 0 > ${o}_tile_probe_30q.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches_30q.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-check > /dev/null 2>&1
exp() {   # $1 = report stem: raw csv always, source csv when $2 is given; the report itself is deleted
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  if [ -n "$2" ]; then ncu -i $1.ncu-rep --page source --csv --print-source sass,cuda 2>/dev/null | gzip -9 > $1.source.csv.gz; fi
  rm -f $1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 6 -c 2 -o ${o}_tile_pass -f python bench.py --steps 1 --warmup 1 --no-cpu --no-check --depth 6 > ${o}_ncu_tile.log 2>&1
exp ${o}_tile_pass src
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 4 -c 1 -o ${o}_tile_pass_qft -f python bench.py --workload qft --qubits 30 --steps 1 --warmup 1 --no-cpu --no-check > ${o}_ncu_tile_qft.log 2>&1
exp ${o}_tile_pass_qft src
if [ -f qvnt_b200/libqvnt_b200_trace.so ]; then
  (QVNT_B200_LIB=$PWD/qvnt_b200/libqvnt_b200_trace.so timeout 200 python tools/trace_pass.py --qubits 30 --depth 12; QVNT_B200_LIB=$PWD/qvnt_b200/libqvnt_b200_trace.so timeout 200 python tools/trace_pass.py --qubits 30 --workload qft) > ${o}_trace_1gpu.txt 2>&1
fi
du -sh gpurun_out; ls -la gpurun_out/r02_final* | awk '{print $5, $9}'
python - <<'PY'
import json
for f in ("bench_30q","bench_qft20","bench_qft30","bench_qft_h_32q","bench_28q","bench_30q_cpp_ops"):
    try:
        d=json.loads(open(f"gpurun_out/r02_final_{f}.json").read().strip().splitlines()[-1])
        r=d.get("roofline") or {}; c=d.get("clocks") or {}
        print(f, round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],2), "frac", r.get("frac"), "ms/pass", r.get("avg_launch_ms"), "clk", c.get("sm_mhz"), c.get("reasons"), c.get("power_w_max"), "e2e", round(d["e2e"]["value"],2))
    except Exception as ex: print(f, "ERR", ex)
PY
