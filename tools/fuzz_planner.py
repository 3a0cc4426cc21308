#!/usr/bin/env python
"""CPU fuzz of the planner's ENCODED output: random circuits (every fast kind with random controls,
the configs[1] generator, QFT pieces, all-kinds mixes), random register / shard / tile geometry,
both settings of the planner options; the numpy emulator of the fast stage interpreter
(tests/tile_emulator.py) executes the micro-ops and the result must match the CPU oracle to 1e-12.
No GPU.  Usage: fuzz_planner.py [seconds=300] [first_seed=0]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import op, workloads  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.test_tile_emulator import _emulate, _oracle_apply, _state  # noqa: E402


def circuit(rng, n):
    parts = []
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(0, 5))
        s = int(rng.integers(1, 1 << 30))
        if k == 0:
            parts.append(workloads.fast_mix(n, int(rng.integers(20, 160)), seed=s))
        elif k == 1:
            parts.append(workloads.random_layered(n, int(rng.integers(1, 6)), seed=s))
        elif k == 2:
            parts.append(workloads.mixed_all_kinds(n, int(rng.integers(10, 80)), seed=s))
        elif k == 3:
            m = int(rng.integers(1, 1 << n))
            parts.append(op.qft(m) if bin(m).count("1") > 1 else op.h(m))
        else:
            parts.append(op.h(int(rng.integers(1, 1 << n))))
    c = parts[0]
    for p in parts[1:]:
        c = c * p
    return c


def one_case(seed):
    """-> ("ok" | "skip" | "fail", description)"""
    rng = np.random.default_rng(seed)
    world = int(rng.choice([1, 1, 2, 4, 8]))
    n = int(rng.integers(max(7, 6 + world.bit_length() - 1), 14))
    n_local = n - (world.bit_length() - 1)
    kw = {}
    if rng.integers(0, 2):
        T = int(rng.integers(6, min(12, n_local) + 1))
        kw["tile_bits"] = T
        kw["chunk_bits"] = int(rng.integers(2, T + 1))
    if world > 1 and rng.integers(0, 3) == 0:
        kw["remap"] = False
    if rng.integers(0, 2):
        kw["lower_two_bit"] = True
    circ = circuit(rng, n)
    v = _state(n, seed)
    what = f"seed={seed} n={n} world={world} kw={kw} ops={len(circ)}"
    try:
        got = _emulate(oracle, n, circ, v.copy(), world=world, **kw)[0]
        err = float(np.abs(got - _oracle_apply(oracle, n, v, circ)).max())
        return ("ok" if err <= 1e-12 else "fail"), f"{what} err={err}"
    except AssertionError as ex:
        if "full-interpreter remap passes" in str(ex):      # the emulator has no full interpreter: GPU tests cover these
            return "skip", what
        return "fail", f"{what} {ex!r}"[:300]
    except Exception as ex:          # planner error
        return "fail", f"{what} {ex!r}"[:300]


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    done = skipped = 0
    while time.time() - t0 < budget:
        st, what = one_case(seed)
        if st == "fail":
            print("FAIL", what, flush=True)
        done += st != "skip"
        skipped += st == "skip"
        seed += 1
    print(f"{done} cases checked, {skipped} skipped (full-interpreter remap pass) in {time.time() - t0:.0f} s, last seed {seed - 1}",
          flush=True)


if __name__ == "__main__":
    main()
