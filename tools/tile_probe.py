#!/usr/bin/env python
"""Micro-probes of the fused tile pass on one GPU: memory-phase efficiency (a pass carrying
almost no arithmetic) and arithmetic scaling (k gates per pass), for several tile geometries.
Prints one line per probe: ms per pass, algorithmic GB/s."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import QReg, op  # noqa: E402
from qvnt_b200.op import MultiOp  # noqa: E402


def time_circ(reg, circ, reps=5):
    arr, n = circ.to_c_array()
    reg.apply_raw(arr, n)
    reg.sync()
    reg.stats_reset()
    reg.set_option("profile", 1)
    for _ in range(reps):
        reg.apply_raw(arr, n)
    reg.sync()
    st = reg.stats()
    reg.set_option("profile", 0)
    ms = max(st["ms"][1] / max(1, st["launches"][1]), 1e-9)
    return ms, st["launches"][1] // reps, st["alg_bytes"][1] / max(1, st["launches"][1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--tile-bits", type=int, default=12)
    ap.add_argument("--chunk-bits", type=int, default=7)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--tma", type=int, default=1)
    ap.add_argument("--no-mem", action="store_true")
    ap.add_argument("--only-rx", type=int, default=0, help="run only the rx probe with this k (for ncu)")
    ap.add_argument("--only-rot8", type=int, default=0, help="run only the rot8 probe with this k (for ncu)")
    args = ap.parse_args()
    n = args.qubits
    reg = QReg.new(n)
    hi = [n - 1, n - 2, n - 3, n - 4, n - 5, n - 6, n - 7, n - 8]
    reg.set_option("tile_ctas", args.ctas)
    reg.set_option("tma", args.tma)
    for tb in ((12, 11) if not args.no_mem else ()):
        for cb in (7, 4):
            if tb - cb > 8:
                continue
            reg.set_option("tile_bits", tb)
            reg.set_option("chunk_bits", cb)
            # memory probe: two H on high qubits (forces a tile pass, ~no arithmetic)
            circ = op.h(1 << hi[0]) * op.h(1 << hi[1])
            ms, passes, bpl = time_circ(reg, circ)
            print(f"mem   T={tb} L={cb} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)
    reg.set_option("tile_bits", args.tile_bits)
    reg.set_option("chunk_bits", args.chunk_bits)
    print(f"--- T={args.tile_bits} L={args.chunk_bits} ctas={args.ctas} tma={args.tma}")
    if args.only_rot8:
        circ = MultiOp()
        for i in range(args.only_rot8):
            circ *= (op.rx if i & 1 else op.ry)(0.1 + i, 1 << (4 * ((i // 8) % 3) + i % 4))
        ms, passes, bpl = time_circ(reg, circ, reps=2)
        print(f"rot8  k={args.only_rot8} passes={passes} ms/pass={ms:.3f}")
        return
    if args.only_rx:
        circ = MultiOp()
        for i in range(args.only_rx):
            circ *= op.rx(0.1 + i, 1 << (i % args.tile_bits))
        ms, passes, bpl = time_circ(reg, circ, reps=2)
        print(f"rx    k={args.only_rx} passes={passes} ms/pass={ms:.3f}")
        return
    for k in (2, 4, 8, 16, 32, 64):
        circ = MultiOp()
        for i in range(k):
            circ *= op.rx(0.1 + i, 1 << (i % args.tile_bits))
        ms, passes, bpl = time_circ(reg, circ)
        print(f"rx    k={k} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)
    for k in (8, 16, 24, 32):
        circ = MultiOp()
        for i in range(k):     # 8 rotations per stage of 4 bits (the shape of configs[1])
            circ *= (op.rx if i & 1 else op.ry)(0.1 + i, 1 << (4 * ((i // 8) % 3) + i % 4))
        ms, passes, bpl = time_circ(reg, circ)
        print(f"rot8  k={k} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)
    for k in (8, 32):
        circ = MultiOp()
        for i in range(k):
            circ *= op.rz(0.1 + i, 1 << (13 + i % 12))
        circ *= op.h(1)*op.h(2)
        ms, passes, bpl = time_circ(reg, circ)
        print(f"rz_hi k={k} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)
    for k in (8, 32):
        circ = MultiOp()
        for i in range(k):
            circ *= op.h(1 << (i % 4))
        ms, passes, bpl = time_circ(reg, circ)
        print(f"h_low4 k={k} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)
    for k in (8, 32):
        circ = MultiOp()
        for i in range(k):
            circ *= op.x(1 << (i % 4)).c(1 << (4 + i % 5))
        ms, passes, bpl = time_circ(reg, circ)
        print(f"cx    k={k} passes={passes} ms/pass={ms:.3f} GB/s={bpl / ms / 1e6:.0f}", flush=True)


if __name__ == "__main__":
    main()
