"""Host logic of qvnt_reg_apply's scheduler (planner.cu), checked on the CPU through
`qvnt_plan_describe`: structural invariants of passes / stages, tile coverage of sharded
registers, and -- with the oracle -- that the scheduled ORDER of ops gives the same state as
the reference's list order (only commuting ops may be exchanged)."""
import numpy as np
import pytest

from qvnt_b200 import op, plan, workloads
from qvnt_b200.op import MultiOp, SingleOp, single

DIAG_KINDS = {op.K_Z, op.K_S, op.K_T, op.K_RZ, op.K_RZZ}
QUAD_KINDS = {op.K_H2, op.K_U2}


def _pmix(o):
    """mix bits as index bits at scheduling time"""
    import copy
    q = copy.copy(o)
    q.a, q.b = o.pa, o.pb
    return _mix(q)


def _mix(o):
    if o.kind in DIAG_KINDS:
        return 0
    return (o.a | o.b) if o.kind in QUAD_KINDS else o.a


def check_structure(passes, q_num, world=1, rank=0):
    wb = world.bit_length() - 1
    n_local = q_num - wb
    for p in passes:
        if p.direct:
            assert _pmix(p.op) >> n_local == 0         # global-qubit gates never run as direct sweeps
            continue
        assert 4 <= p.T <= 12 and p.L <= p.T and p.T - p.L <= 8
        assert p.gpos == sorted(set(p.gpos)) and len(p.gpos) == p.T
        assert p.gpos[:p.L] == list(range(p.L))
        tile_mask = sum(1 << g for g in p.gpos)
        kg = sum(1 for g in p.gpos if g >= n_local)
        assert p.peer == (1 if kg else 0)
        assert len(p.fx_pos) == (p.T - kg) + kg and p.n_tiles == 1 << (n_local - len(p.fx_pos))
        assert p.base_or & tile_mask == 0
        for st in p.stages:
            bits = st.r_lpos + st.t_lpos
            assert sorted(bits) == list(range(p.T))                  # a permutation of the tile-local bits
            regs = [p.gpos[l] for l in st.r_lpos]
            for o in st.ops:
                m = _pmix(o) if o.form != 5 else 0
                assert m & ~sum(1 << g for g in regs) == 0, "partner bit outside the register bits"
                if o.form == 5:                                    # lazy x: a permutation between threads
                    assert o.kind == op.K_X and o.pa & tile_mask == o.pa
                    assert (o.pa | o.pctrl) & sum(1 << g for g in regs) == 0
                elif o.form == 7:                                  # lazy x on a register slot: slot marked inverted
                    assert o.kind == op.K_X and 1 << regs[o.ra] == o.pa
                    assert o.pctrl & sum(1 << g for g in regs) == 0
                elif o.form == 1:
                    assert 1 << regs[o.ra] == o.pa
                elif o.form in (2, 3):
                    assert (1 << regs[o.ra]) | (1 << regs[o.rb]) == o.pa and regs[o.ra] < regs[o.rb]
                elif o.form == 4:
                    assert 1 << regs[o.ra] == o.pa and 1 << regs[o.rb] == o.pb
                else:
                    assert o.form == 0 and o.kind in DIAG_KINDS


SPLIT_KINDS = {op.K_X, op.K_Y, op.K_Z, op.K_S, op.K_T}     # multi-bit masks are scheduled bit by bit
LOWERED_KINDS = {op.K_SWAP, op.K_ISWAP, op.K_RXX, op.K_RYY}  # two-bit kinds scheduled as products of fast kinds


def planned_sequence(passes, circ):
    """The ops in scheduled order, rebuilt from the caller's SingleOps (the planner's exact
    factorisations honoured: x/y/z/s/t(mask) = product over bits, h2(a, b) = h1(a) h1(b))."""
    src = list(circ)
    out = []
    for p in passes:
        for o in p.all_ops():
            s = src[o.src].clone()
            if s.kind in LOWERED_KINDS and o.kind != s.kind:
                # swap / i_swap / rxx / ryy of an op list are scheduled as products of fast kinds
                # (planner.cu lower_ops): rebuild the factor from what the planner reports
                assert o.a & ~s.a_mask == 0 and o.ctrl & ~(s.ctrl | s.a_mask) == 0
                if o.kind == op.K_RZZ:
                    f = SingleOp(op.K_RZZ, o.a, phase=s.phase)
                else:
                    f = SingleOp(o.kind, o.a, dagger=bool(o.dagger))
                    assert o.kind in (op.K_X, op.K_Z, op.K_S, op.K_H1)
                out.append(f.c(o.ctrl) if o.ctrl else f)
                continue
            assert s.ctrl == o.ctrl
            if s.kind == op.K_H2 and o.kind == op.K_H1:
                assert o.a in (s.a_mask, s.b_mask)
                # the half exactly as the planner scheduled it: a butterfly scaled by o.scale (1 and 0.5,
                # not 1/sqrt2 twice), so replaying one half through the oracle and emulating the other agree
                f = o.scale
                s = SingleOp(op.K_U1, o.a, matrix=[complex(f), complex(f), complex(f), complex(-f)])   # (not unitary alone)
                if o.ctrl:
                    s = s.c(o.ctrl)
            else:
                assert s.kind == o.kind
                if s.a_mask != o.a:
                    assert s.kind in SPLIT_KINDS and o.a & ~s.a_mask == 0
                    s.a_mask = o.a
            out.append(s)
    return MultiOp(out)


def every_op_scheduled_once(passes, circ):
    seen = {}
    for p in passes:
        for o in p.all_ops():
            seen.setdefault(o.src, 0)
            if circ[o.src].kind in LOWERED_KINDS and o.kind != circ[o.src].kind:
                seen[o.src] = -1                       # (the product's factors: checked by the replay)
                continue
            if o.kind in SPLIT_KINDS or (o.kind == op.K_H1 and circ[o.src].kind == op.K_H2):
                assert seen[o.src] & o.a == 0
                seen[o.src] |= o.a
            else:
                assert seen[o.src] == 0
                seen[o.src] = -1
    for i, s in enumerate(circ):
        if s.kind == op.K_ID:
            continue
        if s.kind in SPLIT_KINDS:
            assert seen.get(i, 0) == s.a_mask or (s.a_mask == 0 and i not in seen)
        elif s.kind == op.K_H2:
            assert seen.get(i) in (-1, s.a_mask | s.b_mask)
        else:
            assert seen.get(i) == -1, (i, s)


@pytest.mark.parametrize("n,circ_fn", [
    (10, lambda n: workloads.mixed_all_kinds(n, 300, seed=1)),
    (13, lambda n: workloads.mixed_all_kinds(n, 200, seed=2)),
    (14, lambda n: workloads.random_layered(n, 12)),
    (12, lambda n: op.qft((1 << n) - 1) * op.h((1 << n) - 1)),
])
@pytest.mark.parametrize("tile_bits,chunk_bits", [(0, 0), (8, 4), (6, 3)])
@pytest.mark.parametrize("lower_two_bit", [False, True])
def test_schedule_preserves_result(oracle, n, circ_fn, tile_bits, chunk_bits, lower_two_bit):
    circ = circ_fn(n)
    if lower_two_bit and not any(s.kind in LOWERED_KINDS for s in circ):
        pytest.skip("no swap / i_swap / rxx / ryy in this circuit")
    passes = plan.describe(n, circ, tile_bits=tile_bits, chunk_bits=chunk_bits, lower_two_bit=lower_two_bit)
    if lower_two_bit:
        assert any(o.kind != circ[o.src].kind and circ[o.src].kind in LOWERED_KINDS for p in passes for o in p.all_ops())
    check_structure(passes, n)
    every_op_scheduled_once(passes, circ)
    seq = planned_sequence(passes, circ)
    rng = np.random.default_rng(n)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    a, b = oracle.OracleReg.new(n), oracle.OracleReg.new(n)
    a.write_amplitudes(v)
    b.write_amplitudes(v)
    a.apply(circ)
    b.apply(seq)
    assert np.abs(a.amplitudes() - b.amplitudes()).max() <= 1e-12


def test_unfused_schedule_is_list_order():
    n = 12
    circ = workloads.mixed_all_kinds(n, 100, seed=3)
    passes = plan.describe(n, circ, fuse=False)
    assert all(p.direct for p in passes)
    kept = [i for i, s in enumerate(circ) if not (s.kind in (op.K_X, op.K_Y, op.K_Z, op.K_S, op.K_T) and s.a_mask == 0)]
    assert [p.op.src for p in passes] == kept


def test_fusion_depth_of_headline_workloads():
    # configs[1]: 4150 SingleOps; configs[2]: 528 + 16
    s = plan.summary(plan.describe(28, workloads.random_layered(28, 100)))
    assert s["ops"] == 4150 and s["passes"] <= 260
    s = plan.summary(plan.describe(32, workloads.qft_plus_h(32)))
    assert s["ops"] == 528 + 32 and s["passes"] <= 8     # the 16 h2 run as 32 h1


def _tile_indices(p, n_local, tile_i):
    base = tile_i
    for pos in p.fx_pos:
        base = ((base >> pos) << (pos + 1)) | (base & ((1 << pos) - 1))
    base |= p.fx_val | p.base_or
    idx = np.full(1, base, dtype=np.uint64)
    for g in p.gpos:
        idx = np.concatenate([idx, idx | np.uint64(1 << g)])
    return idx


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_tiles_cover_every_amplitude_once(world):
    n = 12
    wb = world.bit_length() - 1
    n_local = n - wb
    circ = workloads.mixed_all_kinds(n, 120, seed=world) * op.h((1 << n) - 1) * op.qft((1 << n) - 1)
    plans = [plan.describe(n, circ, rank=r, world=world, peers=True, tile_bits=8, chunk_bits=3) for r in range(world)]
    n_pass = len(plans[0])
    assert all(len(p) == n_pass for p in plans)
    peer_passes = 0
    for k in range(n_pass):
        p0 = plans[0][k]
        for r in range(world):
            pr = plans[r][k]
            assert pr.direct == p0.direct
            if not pr.direct:
                assert pr.gpos == p0.gpos and [(o.src, o.a) for o in pr.all_ops()] == [(o.src, o.a) for o in p0.all_ops()]
        if p0.direct:
            continue
        check_structure([plans[r][k] for r in range(world)], n, world)
        peer_passes += p0.peer
        count = np.zeros(1 << n, dtype=np.int32)
        for r in range(world):
            pr = plans[r][k]
            for t in range(pr.n_tiles):
                idx = _tile_indices(pr, n_local, t)
                count[idx.astype(np.int64)] += 1
                if not pr.peer:
                    assert np.all(idx >> np.uint64(n_local) == r)
        assert np.all(count == 1)
    assert peer_passes >= 1


def test_global_gate_without_peers_is_refused():
    from qvnt_b200 import QvntError
    with pytest.raises(QvntError) as e:
        plan.describe(10, op.h(1 << 9), rank=0, world=2, peers=False)
    assert e.value.status == 5      # QVNT_ERR_COMM
    # diagonal gates and controls on global qubits need no peers
    passes = plan.describe(10, op.rz(0.3, 1 << 9) * op.x(1).c(1 << 9) * op.z(0x3FF), rank=1, world=2, peers=False)
    assert sum(len(p.all_ops()) for p in passes) == 2 + 10      # z(mask) is scheduled bit by bit


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_preserves_result(oracle, world):
    """The order every RANK of a sharded register schedules (peer tile passes included) is a
    reordering of commuting ops only: replayed through the oracle it gives the list-order state."""
    n = 12
    circ = workloads.mixed_all_kinds(n, 150, seed=9) * workloads.random_layered(n, 4)
    rng = np.random.default_rng(3)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    a = oracle.OracleReg.new(n)
    a.write_amplitudes(v)
    a.apply(circ)
    want = a.amplitudes().copy()
    for rank in range(world):
        passes = plan.describe(n, circ, rank=rank, world=world, peers=True)
        check_structure(passes, n, world, rank)
        every_op_scheduled_once(passes, circ)
        b = oracle.OracleReg.new(n)
        b.write_amplitudes(v)
        b.apply(planned_sequence(passes, circ))
        assert np.abs(b.amplitudes() - want).max() <= 1e-12


@pytest.mark.parametrize("tile_bits,cap", [(12, 1024), (11, 2048)])
def test_long_pass_programs_are_split(tile_bits, cap):
    """A pass program must fit the CTA's shared memory next to the tile: very long runs of ops that
    would all fit ONE pass (diagonal gates never constrain the tile) are cut into several passes."""
    n = 16
    circ = MultiOp()
    for k in range(5000):
        circ *= op.rz(0.001 * (k + 1), 1 << (k % n)).c(1 << ((k + 5) % n))
    passes = plan.describe(n, circ, tile_bits=tile_bits)
    assert sum(len(p.all_ops()) for p in passes) == 5000
    assert len(passes) >= 3
    for p in passes:
        assert len(p.all_ops()) <= cap and len(p.stages) <= (32 if tile_bits >= 12 else 64)
