#!/bin/bash
# timeline probe (tools/trace_pass.py) on 2 GPUs: the busiest local and remap passes of the 30-qubit circuit
export QVNT_B200_LIB=$PWD/qvnt_b200/libqvnt_b200_trace.so
mkdir -p gpurun_out
out=gpurun_out/r02_trace_2gpu.txt; : > $out
t() { timeout 240 python tools/trace_pass.py "$@" >> $out 2>&1; }
t --qubits 30 --depth 12
t --qubits 30 --depth 12 --gpus 2
t --qubits 30 --depth 12 --gpus 2 --opt peer_tile_bits=11 double_buffer=2
t --qubits 30 --depth 12 --gpus 2 --opt peer_tile_bits=11
cat $out
