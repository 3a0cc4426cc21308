#!/usr/bin/env python
"""Micro-op composition of a workload's schedule (host only, no GPU): passes, stages, and how many
ops run as which arm of the fast stage interpreter -- the instruction-cost model behind
DESIGN.md section 9 (pair / diagonal arms: 64 FP64 instr per thread + ~16 of fetch / dispatch; x with a control in a register
slot: 48 moves; lazy x: a few integer instr; merged diagonal run members: ~14 instr)."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import plan, workloads  # noqa: E402

NAMES = {0: "pair_real", 4: "pair_cross", 8: "diag_slot", 12: "diag_thread", 13: "diag_generic", 14: "lazy_x_thread",
         15: "lazy_x_slot", 16: "diag_run_header", 40: "swap_x(ctrl in slot)"}
COST = {"pair_real": 80, "pair_cross": 80, "diag_thread": 78, "diag_slot": 80,
        "diag_generic": 120, "swap_x(ctrl in slot)": 70,      # (48 moves where the control holds)
        "butterfly_h": 56, "single-control diag_slot": 50, "single-control diag_thread": 55, "single-control diag_run_header": 60,
        "lazy_x_thread": 16, "lazy_x_slot": 10, "diag_run_header": 80, "diag_run_member": 14}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="random28", choices=["random28", "qft", "qft_h"])
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--depth", type=int, default=100)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--chunk-bits", type=int, default=0)
    a = ap.parse_args()
    n = a.qubits
    circ = (workloads.random_layered(n, a.depth) if a.workload == "random28"
            else workloads.qft_full(n) if a.workload == "qft" else workloads.qft_plus_h(n))
    ps = plan.describe(n, circ, world=a.world, peers=a.world > 1, tile_bits=a.tile_bits, chunk_bits=a.chunk_bits)
    cnt = collections.Counter()
    full = direct = 0
    for p in ps:
        if p.direct:
            direct += 1
            continue
        if p.full:
            full += 1
        for st in p.stages:
            skip = 0
            for m in st.mops:
                if skip:
                    cnt["diag_run_member"] += 1
                    skip -= 1
                    continue
                raw = m.code
                if 132 <= raw < 156:          # single-control arms (engine.h FC_DS1 / FC_DU1 / FC_DM1)
                    cnt["single-control " + ("diag_slot" if raw < 148 else "diag_thread" if raw < 152 else "diag_run_header")] += 1
                    if raw >= 152:
                        skip = m.a_reg
                    continue
                if 156 <= raw < 160:          # butterfly h (FC_HB)
                    cnt["butterfly_h"] += 1
                    continue
                c = raw % 44
                c = c - 20 if 20 <= c < 40 else c
                base = max(k for k in NAMES if k <= c)
                cnt[NAMES[base]] += 1
                if base == 16:
                    skip = m.a_reg
    s = plan.summary(ps)
    print(f"{len(circ)} SingleOps -> {s['passes']} passes ({direct} direct, {full} full-interpreter, "
          f"{s['peer_passes']} peer), {s['stages']} stages, {s['ops_per_pass']:.1f} ops/pass")
    tot = sum(COST.get(k, 110) * v for k, v in cnt.items())
    for k, v in cnt.most_common():
        print(f"  {k:18s} {v:6d}   ~{COST.get(k, 110) * v / max(tot, 1) * 100:4.1f} % of the per-thread instruction work")


if __name__ == "__main__":
    main()
