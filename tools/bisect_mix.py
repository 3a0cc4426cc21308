#!/usr/bin/env python
"""Max |amplitude error| of workloads.fast_mix against the CPU oracle under several option sets
(which arm family a parity failure of the fast interpreter comes from)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import workloads  # noqa: E402
from qvnt_b200.register import QReg  # noqa: E402
from oracle import oracle  # noqa: E402

SETS = [{}, {"butterfly": 0}, {"single_ctrl": 0}, {"butterfly": 0, "single_ctrl": 0}, {"ptx_ops": 0},
        {"fuse": 0}]


def main():
    for n, layers, seed in ((12, 500, 3), (12, 120, 3), (12, 40, 3), (16, 400, 11)):
        circ = workloads.fast_mix(n, layers, seed)
        rng = np.random.default_rng(seed)
        v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        v /= np.linalg.norm(v)
        o = oracle.OracleReg.new(n, threads=oracle.max_threads())
        o.write_amplitudes(v)
        o.apply(circ)
        want = o.amplitudes()
        for opts in SETS:
            try:
                g = QReg.new(n)
                for k, val in opts.items():
                    g.set_option(k, val)
                g.write_amplitudes(v)
                g.apply(circ)
                err = np.abs(g.amplitudes() - want).max()
            except Exception as ex:      # unknown option in an older library
                err = repr(ex)[:60]
            print(f"n={n} layers={layers} seed={seed} {opts}: {err}", flush=True)


if __name__ == "__main__":
    main()
