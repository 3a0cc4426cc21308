#!/usr/bin/env python
"""Per-kernel HBM roofline sweep: every gate class x target bit x
register size on one B200, one SingleOp = one in-place sweep (fuse = 0), CUDA-event time per
launch from the library's profile mode.  Algorithmic bytes = 32 B per amplitude the gate can
change.  Writes a markdown table (stdout) and a CSV (--csv)."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import QReg, op  # noqa: E402
from qvnt_b200.op import MultiOp, single  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_op(reg, sop, reps):
    arr, n = MultiOp([sop]).to_c_array()
    reg.apply_raw(arr, n)
    reg.sync()
    reg.stats_reset()
    reg.set_option("profile", 1)
    for _ in range(reps):
        reg.apply_raw(arr, n)
    reg.sync()
    st = reg.stats()
    reg.set_option("profile", 0)
    cls = 0 if st["launches"][0] else 1
    ms = st["ms"][cls] / max(1, st["launches"][cls])
    return ms, st["alg_bytes"][cls] / max(1, st["launches"][cls])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, nargs="+", default=[30, 32, 33])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--csv", default="")
    args = ap.parse_args()
    pk = peak()
    rows = []
    for n in args.qubits:
        reg = QReg.new(n)
        reg.set_option("fuse", 0)
        reg.apply(op.h((1 << min(n, 12)) - 1))      # a non-trivial state
        bits = sorted({0, 1, 4, 5, 8, 12, n // 2, n - 2, n - 1})
        th = 1.23456
        cases = []
        for b in bits:
            m = 1 << b
            cases += [("h1", b, single.h1(m)), ("rx", b, single.rx(m, th)), ("ry", b, single.ry(m, th)),
                      ("rz", b, single.rz(m, th)), ("x", b, single.x(m)), ("z", b, single.z(m)),
                      ("t", b, single.t(m))]
        for a, b in [(0, 1), (3, 17), (n - 2, n - 1), (5, n - 1)]:
            ab = (1 << a) | (1 << b)
            lbl = f"{a},{b}"
            cases += [("h2", lbl, single.h2(1 << a, 1 << b)), ("rxx", lbl, single.rxx(ab, th)),
                      ("rzz", lbl, single.rzz(ab, th)), ("swap", lbl, single.swap(ab)),
                      ("i_swap", lbl, single.i_swap(ab)), ("sqrt_swap", lbl, single.sqrt_swap(ab))]
        c, s = math.cos(th), math.sin(th)
        cases += [("u1", n - 3, op.SingleOp(op.K_U1, 1 << (n - 3), matrix=[c, -s, s, c])),
                  ("u2", "2,n-1", op.SingleOp(op.K_U2, 1 << 2, 1 << (n - 1), matrix=[
                      c, 0, 0, -s, 0, c, -s, 0, 0, s, c, 0, s, 0, 0, c]))]
        cases += [("x.c1", n - 1, single.x(1 << (n - 1)).c(1 << 3)),
                  ("rx.c2", 7, single.rx(1 << 7, th).c((1 << 2) | (1 << (n - 2)))),
                  ("rz.c1 (qft)", n - 1, single.rz(1 << (n - 1), th).c(1 << 0))]
        for name, tgt, sop in cases:
            ms, bytes_ = time_op(reg, sop, args.reps)
            gbs = bytes_ / ms / 1e6
            rows.append((n, name, str(tgt), ms, bytes_ / 1e9, gbs, gbs / pk))
        # reductions
        reg.stats_reset()
        reg.set_option("profile", 1)
        for _ in range(args.reps):
            reg.get_absolute()
        st = reg.stats()
        reg.set_option("profile", 0)
        ms = st["ms"][2] / max(1, st["launches"][2]) * (st["launches"][2] / args.reps)
        gb = st["alg_bytes"][2] / args.reps / 1e9
        rows.append((n, "norm_sqr (get_absolute)", "-", ms, gb, gb / ms * 1e3, gb / ms * 1e3 / pk))
        reg.close()
    print(f"| n | gate | target | ms | alg GB | GB/s | of measured {pk:.0f} |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.3f} | {r[4]:.2f} | {r[5]:.0f} | {r[6]:.2f} |")
    if args.csv:
        with open(args.csv, "w") as f:
            f.write("n;gate;target;ms;alg_gb;gbs;frac_of_measured\n")
            for r in rows:
                f.write(";".join(str(x) for x in r) + "\n")
    fr = [r[6] for r in rows]
    print(f"\nmin {min(fr):.2f}  median {sorted(fr)[len(fr) // 2]:.2f}  max {max(fr):.2f} of the measured copy peak "
          f"({len(rows)} sweeps)")


if __name__ == "__main__":
    main()
