#!/usr/bin/env python
"""Runs every kernel of the library once or twice at a realistic size (for `ncu --set full` captures of
the kernels the headline bench does not launch): direct sweeps incl. the quad class, the reductions,
measurement (block weights, locate, collapse), zero/scale/set_basis, tensor product, sample_all,
combine / linear_composition."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qvnt_b200 import QReg, op  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    a = ap.parse_args()
    n = a.qubits
    r = QReg.new(n)
    r.set_option("fuse", 0)
    r.apply(op.h((1 << n) - 1))                                   # k_direct_quad<H2> x n/2
    r.apply(op.rx(0.3, 1 << 5).c(1 << 1))                         # k_direct_pair, control on a low bit
    r.apply(op.rz(0.3, 1 << (n - 1)) * op.t(0b101))               # k_direct_diag
    r.apply(op.swap((1 << 3) | (1 << (n - 2))))                   # k_direct_pair (odd-parity subspace)
    r.set_option("fuse", 1)
    print("norm", r.get_absolute())                               # k_norm_partial + k_sum_partials
    print("measure", r.measure_mask_full((1 << (n - 1)) | 0b1011, 0.37))   # k_block_weights/sums, k_total, k_locate, k_zero_where
    r.reset_by_mask(0b110)                                        # k_zero_where<1> + k_norm + k_scale
    h = r.sample_all(1 << 20, seed=3)                             # k_sample_noise_sum, k_sample_counts
    print("sample_all", int(h.sum()))
    r.close()
    m = n // 2
    x, y = QReg.with_state(m, 3), QReg.with_state(n - m, 5)
    x.apply(op.h((1 << m) - 1))
    y.apply(op.h((1 << (n - m)) - 1))
    z = x * y                                                     # k_tensor_prod
    print("tensor", z.q_num)
    z.close()
    x.close()
    y.close()
    p, q = QReg.new(n - 1), QReg.with_state(n - 1, 1)
    c = QReg.combine_with_unitary(p, q, [0.6, 0.8j, 0.8j, 0.6])   # k_combine_unitary
    p.linear_composition(q, (0.6, 0.8j))                          # k_linear_composition
    print("combine", c.q_num, c.get_absolute())
    for t in (p, q, c):
        t.close()


if __name__ == "__main__":
    main()
