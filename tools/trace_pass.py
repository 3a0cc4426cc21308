#!/usr/bin/env python
"""Timeline of the tile pass from INSIDE the kernel: per-phase cycle counts of the first tiles of the
first CTAs (thread 0's clock64 stamps), for the last local pass and the last remap pass of a run.
Needs the probe build of the library:
    make -C qvnt_b200/csrc VARIANT=_trace EXTRA=-DQV_TRACE
    QVNT_B200_LIB=qvnt_b200/libqvnt_b200_trace.so python tools/trace_pass.py --qubits 30 --depth 8 [--gpus 2]
Columns (cycles of the 1.965 GHz SM clock, mean over CTAs x tiles, first 4 tiles of every CTA dropped):
  wait   loop top -> tile landed in shared memory (+ barrier)
  prep   code patching + next tile's metadata (+ barrier)
  st0/st1  first / second stage: smem -> registers, ops, registers -> smem, barrier
  pre    last stage: smem -> registers, (remap: wait for the peer's ack), barrier, NEXT tile's load issued
  ops    last stage's ops
  store  last stage's stores to HBM issued
  tile   loop top -> next loop top
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import _ffi, workloads  # noqa: E402
from qvnt_b200.register import QReg  # noqa: E402

CTAS, TILES, PTS = 16, 48, 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="random", choices=["random", "qft"])
    ap.add_argument("--opt", nargs="*", default=[])
    a = ap.parse_args()
    lib = _ffi.lib()
    if not hasattr(lib, "qvnt_debug_trace"):
        raise SystemExit("this library has no probe: build with VARIANT=_trace EXTRA=-DQV_TRACE and set QVNT_B200_LIB")
    lib.qvnt_debug_trace.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.qvnt_debug_trace.restype = ctypes.c_int
    n = a.qubits
    circ = workloads.random_layered(n, a.depth) if a.workload == "random" else workloads.qft_full(n)
    reg = QReg.multi(n, 0, a.gpus) if a.gpus > 1 else QReg.with_state(n, 0)
    for kv in a.opt:
        k, v = kv.split("=")
        reg.set_option(k, int(v))
    for _ in range(2):
        reg.reset(0)
        reg.apply(circ)
        reg.sync()
    print(f"# {a.workload} {n} qubits depth {a.depth} gpus {a.gpus} opts {a.opt}")
    for kind, name in ((0, "last LOCAL pass"), (1, "last REMAP pass")):
        st = np.zeros((CTAS, TILES, PTS), dtype=np.uint64)
        info = np.zeros(8, dtype=np.uint32)
        if lib.qvnt_debug_trace(0, kind, st.ctypes.data, info.ctypes.data) != 0:
            raise SystemExit("qvnt_debug_trace failed")
        if info[0] == 0:
            print(f"{name}: none")
            continue
        ns, nops, T, L, nthr, grid, ntiles, fl = (int(x) for x in info)
        per_cta = ntiles // max(grid, 1)
        use = min(TILES, per_cta) - 1
        print(f"{name}: stages {ns} ops {nops} T {T} L {L} threads {nthr} grid {grid} tiles {ntiles} "
              f"bulk {fl & 1} two_buffers {(fl >> 1) & 1} reads_peer {(fl >> 2) & 1}")
        if use < 6:
            print("  too few tiles per CTA")
            continue
        s = st.astype(np.int64)
        lo = 4
        seg = {}
        seg["wait"] = s[:, lo:use, 1] - s[:, lo:use, 0]
        seg["prep"] = s[:, lo:use, 2] - s[:, lo:use, 1]
        prev = s[:, lo:use, 2]
        if ns >= 2:
            seg["st0"] = s[:, lo:use, 3] - prev
            prev = s[:, lo:use, 3]
        if ns >= 3:
            seg["st1"] = s[:, lo:use, 4] - prev
            prev = s[:, lo:use, 4]
        if ns > 3:
            print("  (more than 3 stages: the middle ones are inside `pre`)")
        seg["pre"] = s[:, lo:use, 5] - prev
        seg["ops"] = s[:, lo:use, 6] - s[:, lo:use, 5]
        seg["store"] = s[:, lo:use, 7] - s[:, lo:use, 6]
        seg["tile"] = s[:, lo + 1:use + 1, 0] - s[:, lo:use, 0]
        print("  " + "  ".join(f"{k} {v.mean():8.0f}" for k, v in seg.items()))
        print("  p10/p90 " + "  ".join(f"{k} {np.percentile(v, 10):.0f}/{np.percentile(v, 90):.0f}" for k, v in seg.items()))
        tot = seg["tile"].mean()
        print(f"  one tile per CTA every {tot / 1.965e3:.2f} us; CTAs per SM {grid // 148 if grid >= 148 else 1}")
        # do the CTAs of one SM run in step?  loop-top stamps of CTA 0 over tiles, as offsets
        print("  CTA 0 per-tile cycles: " + " ".join(str(int(x)) for x in (s[0, lo + 1:lo + 13, 0] - s[0, lo:lo + 12, 0])))


if __name__ == "__main__":
    main()
