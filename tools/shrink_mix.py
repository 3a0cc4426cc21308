#!/usr/bin/env python
"""Delta-debug a parity failure of the PTX op loop: shrink workloads.fast_mix(n, layers, seed) to a
minimal op list on which option ptx_ops=1 and ptx_ops=0 (the C++ loop, the specification) disagree,
then print the ops and the encoded plan of the failing case."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import plan, workloads  # noqa: E402
from qvnt_b200.op import MultiOp  # noqa: E402
from qvnt_b200.register import QReg  # noqa: E402


def run(n, ops, v, ptx):
    g = QReg.new(n)
    g.set_option("ptx_ops", ptx)
    g.write_amplitudes(v)
    g.apply(MultiOp(ops))
    return g.amplitudes()


def differs(n, ops, v):
    if len(ops) < 2:
        return False
    return np.abs(run(n, ops, v, 1) - run(n, ops, v, 0)).max() > 1e-9


def main():
    n, layers, seed = (int(x) for x in (sys.argv[1:4] or (12, 40, 3)))
    ops = list(workloads.fast_mix(n, layers, seed))
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    assert differs(n, ops, v), "no difference to shrink"
    changed = True
    while changed:
        changed = False
        k = 0
        while k < len(ops):
            trial = ops[:k] + ops[k + 1:]
            if differs(n, trial, v):
                ops = trial
                changed = True
            else:
                k += 1
    print(f"minimal: {len(ops)} ops")
    for s in ops:
        print("  ", s)
    print(plan.describe_text(n, MultiOp(ops)) if hasattr(plan, "describe_text") else "")
    for p in plan.describe(n, MultiOp(ops)):
        print("pass T", getattr(p, "T", None), "direct", p.direct, "full", p.full)
        for si, st in enumerate(p.stages):
            print("  stage", si, "r_lpos", getattr(st, "r_lpos", None), "t_lpos", getattr(st, "t_lpos", None))
            for m in st.mops:
                print(f"    code={m.code} flags={m.flags:#x} okmask={m.okmask:#06x} ctrl_thr={m.ctrl_thr:#x} a_thr={m.a_thr:#x} "
                      f"a_reg={m.a_reg} idx={m.idx} c={m.c} alt={m.alt} ctrl_base={m.ctrl_base:#x} a_base={m.a_base:#x}")
    a1, a0 = run(n, ops, v, 1), run(n, ops, v, 0)
    bad = np.nonzero(np.abs(a1 - a0) > 1e-9)[0]
    print("wrong amplitudes:", len(bad), "first", bad[:16], "index bits set in all / none:",
          bin(int(np.bitwise_and.reduce(bad))), bin(int(np.bitwise_or.reduce(bad))))


if __name__ == "__main__":
    main()
