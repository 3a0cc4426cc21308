"""The planner's ENCODED micro-ops, executed by a numpy emulator of the kernel's fast stage
interpreter (tests/tile_emulator.py), must reproduce the oracle's state: coefficient forms,
slot / thread / tile splits of controls and diagonal masks, lazy x, merged diagonal runs, tile
geometry -- and, for sharded registers, tile ownership and rank-dependent flags.  CPU only."""
import os
import numpy as np
import pytest

from qvnt_b200 import op, plan, workloads
from qvnt_b200.op import MultiOp
from tests import tile_emulator as emu
from tests.test_planner import planned_sequence


def _oracle_apply(oracle, n, psi, mop):
    r = oracle.OracleReg.new(n)
    r.write_amplitudes(psi)
    r.apply(mop)
    out = r.amplitudes().copy()
    r.close()
    return out


def _emulate(oracle, n, circ, psi, world=1, **kw):
    """psi is kept indexed by PHYSICAL index bits; `perm` (qubit -> index bit) follows the remap
    passes; the result is returned indexed by qubits."""
    plans = [plan.describe(n, circ, rank=r, world=world, peers=world > 1, **kw) for r in range(world)]
    n_local = n - (world.bit_length() - 1)
    perm = list(range(n))
    n_fast = n_lazy = n_runs = n_remap = 0
    for k, p0 in enumerate(plans[0]):
        if p0.direct or p0.full:
            # direct sweeps and full-interpreter passes: their ops, in scheduled order, via the oracle
            assert not (not p0.direct and p0.remap), "full-interpreter remap passes are covered on the GPU"
            lg = _oracle_apply(oracle, n, emu.to_logical(psi, perm), planned_sequence([p0], circ))
            psi = emu.to_physical(lg, perm)
            continue
        n_fast += 1
        snap = psi.copy() if p0.remap else None
        for r in range(world):
            p = plans[r][k]
            assert not p.direct and not p.full and p.gpos == p0.gpos
            assert (p.remap, p.rg, p.rb) == (p0.remap, p0.rg, p0.rb)
            emu.run_tile_pass(psi, p, rank=r, n_local=n_local, src=snap)
        if p0.remap:
            n_remap += 1
            qa, qb = perm.index(p0.rg), perm.index(p0.rb)
            perm[qa], perm[qb] = p0.rb, p0.rg
        for st in p0.stages:
            gen = [emu.generic_code(m) % emu.FC_TOTAL for m in st.mops]     # (single-control arms -> their generic code)
            n_lazy += sum(1 for c in gen if c in (emu.FC_LX, emu.FC_LI))
            n_runs += sum(1 for c in gen if c in (emu.FC_DM, emu.FC_DM + emu.FC_MASKED))
    _emulate.last_remaps = n_remap
    return emu.to_logical(psi, perm), n_fast, n_lazy, n_runs


def _state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


_fast_mix = workloads.fast_mix


@pytest.mark.parametrize("n,circ_fn", [
    (12, lambda n: workloads.random_layered(n, 8)),
    (13, lambda n: op.qft((1 << n) - 1) * op.h((1 << n) - 1)),
    (12, lambda n: _fast_mix(n, 250, seed=3)),
    (14, lambda n: _fast_mix(n, 150, seed=4) * workloads.random_layered(n, 3)),
])
@pytest.mark.parametrize("tile_bits,chunk_bits", [(0, 0), (8, 3), (12, 7)])
def test_encoded_plan_reproduces_oracle(oracle, n, circ_fn, tile_bits, chunk_bits):
    circ = circ_fn(n)
    v = _state(n, n)
    want = _oracle_apply(oracle, n, v, circ)
    got, n_fast, n_lazy, n_runs = _emulate(oracle, n, circ, v.copy(), tile_bits=tile_bits, chunk_bits=chunk_bits)
    assert n_fast >= 1
    assert np.abs(got - want).max() <= 1e-12


def test_emulator_covers_lazy_x_and_merged_runs(oracle):
    n = 13
    circ = workloads.random_layered(n, 10) * op.qft((1 << n) - 1)
    v = _state(n, 7)
    got, n_fast, n_lazy, n_runs = _emulate(oracle, n, circ, v.copy())
    assert n_lazy >= 5 and n_runs >= 3
    assert np.abs(got - _oracle_apply(oracle, n, v, circ)).max() <= 1e-12


@pytest.mark.parametrize("world", [2, 4, 8])
def test_encoded_sharded_plan_reproduces_oracle(oracle, world):
    """Every rank's tiles of every pass, emulated on the global state: ownership, peer chunks and
    the flags derived from the rank bits."""
    n = 13
    top = n - 1
    circ = workloads.random_layered(n, 6) * op.qft((1 << n) - 1) * _fast_mix(n, 80, seed=world)
    circ *= op.x(1 << top).c(1 << 0) * op.rz(0.3, 1 << top).c(1 << (top - 1)) * op.h(1 << top) * op.x(1 << 2).c(1 << top)
    v = _state(n, 11 + world)
    want = _oracle_apply(oracle, n, v, circ)
    got, n_fast, _, _ = _emulate(oracle, n, circ, v.copy(), world=world)
    assert n_fast >= 2 and _emulate.last_remaps >= 2          # global qubits are swapped in, not exchanged twice
    assert np.abs(got - want).max() <= 1e-12
    got, _, _, _ = _emulate(oracle, n, circ, v.copy(), world=world, remap=False)
    assert _emulate.last_remaps == 0
    assert np.abs(got - want).max() <= 1e-12


# ---------------------------------------------------------------- the op loop's code numbering
def _engine_codes():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "qvnt_b200", "csrc", "engine.h")).read()
    return {k: int(v) for k, v in re.findall(r"\b(FC_[A-Z0-9_]+|MOP_END|MOP_NOP|MOP_NOP_RUN)\s*=\s*(\d+)", txt)}


def test_code_numbering_is_consistent():
    """engine.h is the definition; the PTX generator and the emulator restate it.  Every code whose arm
    works on one register slot must keep code & 3 == slot in every variant (the loop's prologue reads
    the slot's inversion byte from the code's low bits): the bases are multiples of 4."""
    import importlib.util
    c = _engine_codes()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_fastops", os.path.join(root, "qvnt_b200", "csrc", "gen_fastops.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    for name in ("FC_PR", "FC_PX", "FC_DS", "FC_DU", "FC_DG", "FC_LX", "FC_LI", "FC_DM", "FC_MASKED", "FC_SW", "FC_TOTAL",
                 "FC_DS1", "FC_DU1", "FC_DM1", "FC_HB"):
        assert getattr(gen, name) == c[name], name
        if hasattr(emu, name):
            assert getattr(emu, name) == c[name], name
    assert emu.FC_SPECIAL_END == c["FC_SPECIAL_END"]
    for base in ("FC_PR", "FC_PX", "FC_DS", "FC_MASKED", "FC_SW", "FC_TOTAL", "FC_DS1", "FC_HB"):
        assert c[base] % 4 == 0, base
    assert 3 * c["FC_TOTAL"] <= c["FC_DS1"] and c["FC_SPECIAL_END"] <= c["MOP_NOP_RUN"] < c["MOP_NOP"] < c["MOP_END"] == 255


def test_generated_ptx_is_current(tmp_path):
    """fastops_ptx.inc is committed (the build does not need python); it must be what the generator emits."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "qvnt_b200", "csrc")
    out = tmp_path / "fastops_ptx.inc"
    subprocess.run([sys.executable, os.path.join(src, "gen_fastops.py"), str(out)], check=True, capture_output=True)
    assert out.read_text() == open(os.path.join(src, "fastops_ptx.inc")).read()


# ---------------------------------------------------------------- butterfly h: the scale travels
def _count_codes(n, circ, lo, hi, **kw):
    return sum(1 for p in plan.describe(n, circ, **kw) if not p.direct for st in p.stages for m in st.mops if lo <= m.code < hi)


@pytest.mark.parametrize("n,circ_fn", [
    # passes made of butterflies only: the last one turns back into a pair op that carries the product
    (13, lambda n: op.h((1 << n) - 1)),
    # single h1 gates (scale 1/sqrt2 each) around lazy inversions of their slots and a carrier of the crossed form
    (12, lambda n: MultiOp([op.single.h1(1 << q) for q in range(n)]) * op.x(0b101).c(0b10) *
         MultiOp([op.single.h1(1 << q) for q in range(n)]) * op.rx(0.3, 1 << 5) * op.x(1 << 7) * op.h(0b11 << 6)),
    # controlled h is NOT a butterfly (its factor is not on every amplitude); it may be the carrier only if unconditional
    (12, lambda n: op.h(0b1111).c(1 << 9) * op.h((1 << n) - 1) * op.rz(0.7, 1 << 3).c(1 << 4) * op.h(0b110000)),
])
@pytest.mark.parametrize("tile_bits,chunk_bits", [(0, 0), (8, 3)])
def test_butterfly_h_scale_is_folded(oracle, n, circ_fn, tile_bits, chunk_bits):
    circ = circ_fn(n)
    assert _count_codes(n, circ, emu.FC_HB, emu.FC_HB + 4, tile_bits=tile_bits, chunk_bits=chunk_bits) >= 4
    v = _state(n, 3)
    got, n_fast, _, _ = _emulate(oracle, n, circ, v.copy(), tile_bits=tile_bits, chunk_bits=chunk_bits)
    assert n_fast >= 1
    assert np.abs(got - _oracle_apply(oracle, n, v, circ)).max() <= 1e-12


def test_planner_fuzz_smoke():
    """A fixed slice of tools/fuzz_planner.py (random circuits x shard counts x tile geometries x planner
    options): the emulated encoded plan must match the oracle.  (2771 + 709 cases ran clean at the end of round 2.)"""
    from tools import fuzz_planner
    res = [fuzz_planner.one_case(seed) for seed in range(40)]
    bad = [w for st, w in res if st == "fail"]
    assert not bad, bad
    assert sum(1 for st, _ in res if st == "ok") >= 25
