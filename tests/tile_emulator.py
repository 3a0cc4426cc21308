"""A numpy emulator of the fused tile pass's FAST stage interpreter (qvnt_b200/csrc/tile.cu),
driven by the planner's ENCODED output (`qvnt_plan_describe` -> plan.PlanMop: codes, okmask,
thread masks, coefficients, per-tile base masks) exactly as the kernel reads it.

It checks on the CPU what the oracle replay of the scheduled ORDER (tests/test_planner.py) cannot:
that the micro-op encodings -- coefficient forms, register-slot / thread / tile splits of controls
and diagonal masks, lazy x index permutations, merged diagonal runs, tile geometry and (for sharded
registers) tile ownership and rank-dependent flags -- mean what the op list says.  Timing hazards
are out of its reach (see TStage::sync_after_load); arithmetic is complex128 without FMA, so
results agree with the oracle to ~1e-15, not bit for bit.

FCode values mirror engine.h.
"""
from __future__ import annotations

import numpy as np

FC_PR, FC_PX, FC_DS, FC_DU, FC_DG, FC_LX, FC_LI, FC_DM, FC_MASKED, FC_SW = 0, 4, 8, 12, 13, 14, 15, 16, 20, 40
FC_TOTAL = 44
MOP_SKIP0, MOP_COND, MOP_PARB, MOP_CONDB, MOP_ATHR, MOP_STATIC = 0x02, 0x04, 0x08, 0x10, 0x20, 0x40
NV = 16
FC_DS1, FC_DU1, FC_DM1, FC_HB, FC_SPECIAL_END = 132, 148, 152, 156, 160


def generic_code(m) -> int:
    """The single-control arms (engine.h FC_DS1 / FC_DU1 / FC_DM1: one control in register slot c, no
    other control) compute what the generic masked arm computes under okmask = 'slot bit c set'; the
    butterfly h (FC_HB + j) is FC_PR + j with the coefficients (1, 1, 1, -1) its descriptor carries."""
    c = m.code
    if not (FC_DS1 <= c < FC_SPECIAL_END):
        return c
    assert not (m.flags & (MOP_COND | MOP_CONDB)) and m.ctrl_thr == 0 and m.ctrl_base == 0
    if c >= FC_HB:
        assert m.okmask == 0xFFFF and tuple(m.c) == (1.0, 1.0, 1.0, -1.0) and tuple(m.alt) == (-1.0, 1.0, 1.0, 1.0)
        assert (c & 3) == c - FC_HB, "the op loop takes the slot's inversion byte from code & 3"
        return FC_PR + (c - FC_HB)
    if c < FC_DU1:
        ctl, j = (c - FC_DS1) >> 2, (c - FC_DS1) & 3
        assert j != ctl
        assert (c & 3) == j, "the op loop takes the target slot's inversion byte from code & 3"
        g = FC_MASKED + FC_DS + j
    elif c < FC_DM1:
        ctl = c - FC_DU1
        g = FC_MASKED + FC_DU
    else:
        ctl = c - FC_DM1
        g = FC_MASKED + FC_DM
    assert m.okmask == sum(1 << K for K in range(NV) if K & (1 << ctl)), "single-control arm with another okmask"
    return g


U64 = (1 << 64) - 1


def _popc(x):
    return np.bitwise_count(np.asarray(x, dtype=np.uint64)).astype(np.int64)


def _pc(x: int) -> int:
    return bin(x & U64).count("1")


def run_stage(tile: np.ndarray, st, T: int, gbase: int):
    """One stage on one tile (in place).  `gbase`: global index of the tile's first amplitude
    (tile bits clear, rank bits included) -- what the kernel's prepare() derives the flags from.

    Registers are modelled the way the kernel holds them: regs[g, K] is PHYSICAL register K of
    thread g; after lazy inversions (FC_LI) it holds the amplitude of slot pattern K ^ ib[g]."""
    G = 1 << (T - 4)
    grp = np.arange(G, dtype=np.int64)
    jl = np.zeros(G, dtype=np.int64)
    for k, lp in enumerate(st.t_lpos):
        jl |= ((grp >> k) & 1) << lp
    kbits = np.zeros(NV, dtype=np.int64)
    for K in range(NV):
        for s in range(4):
            if K >> s & 1:
                kbits[K] |= 1 << st.r_lpos[s]
    regs = tile[jl[:, None] | kbits[None, :]].copy()          # G x 16
    vgrp = grp.copy()
    jl_cur = jl.copy()
    ib = np.zeros(G, dtype=np.int64)                          # inverted slots per thread
    mops = st.mops
    i = 0

    def cond(m):
        # the kernel tests controls only for ops flagged MOP_COND / MOP_CONDB: the flags must be right
        act = np.ones(G, dtype=bool)
        if m.ctrl_thr or m.ctrl_base:
            assert m.flags & MOP_COND, "op with thread / outer controls must carry MOP_COND"
        if m.ctrl_base:
            assert m.flags & MOP_CONDB
        if m.flags & MOP_COND:
            act &= (~vgrp & m.ctrl_thr) == 0
            if m.flags & MOP_CONDB:
                act &= (~gbase & m.ctrl_base & U64) == 0
        return act

    def dpar(m):
        if m.a_base:
            assert m.flags & MOP_PARB, "diagonal op with target bits outside the tile must carry MOP_PARB"
        if m.a_thr:
            assert m.flags & MOP_ATHR, "diagonal op with target bits on thread bits must carry MOP_ATHR"
        c = _popc(vgrp & m.a_thr)
        if m.flags & MOP_PARB:
            c = c + (_pc(gbase & m.a_base) & 7)
        return c & 1

    def okmask(m, masked):
        if not masked:
            assert m.okmask == 0xFFFF, "unmasked arms ignore okmask: only without slot controls"
            return np.full((G, NV), True)
        # register K holds slot pattern K ^ ib: its predicate is okmask bit K ^ ib (perm_ok)
        Ks = np.arange(NV)[None, :] ^ ib[:, None]
        return ((m.okmask >> Ks) & 1).astype(bool)

    while i < len(mops):
        m = mops[i]
        raw = generic_code(m)
        code = raw % FC_TOTAL                    # the byte also carries the control class (engine.h)
        assert raw // FC_TOTAL == (2 if m.flags & MOP_CONDB else 1 if m.flags & MOP_COND else 0)
        masked = False
        if FC_MASKED <= code < FC_SW:
            code -= FC_MASKED
            masked = True
        elif code >= FC_SW:
            masked = True
        act = cond(m)
        if code == FC_DM:
            cnt = m.a_reg
            okm = okmask(m, masked)
            if m.flags & MOP_STATIC:
                # tabulated run: the kernel takes the thread-bit members' factor from a table built with
                # vgrp == thread index, and the outside members' factor from a per-tile table
                assert not any(q.code < 3 * FC_TOTAL and q.code % FC_TOTAL == FC_LX for q in mops), \
                    "a tabulated run needs vgrp == thread index"
                assert all((e.a_thr != 0) != (e.a_base != 0) for e in mops[i + 1:i + 1 + cnt])
                assert bool(m.flags & MOP_PARB) == any(e.a_base != 0 for e in mops[i + 1:i + 1 + cnt])
                assert 0 <= m.a_thr < 12
            acc = np.ones(G, dtype=np.complex128)
            for k in range(1, cnt + 1):
                e = mops[i + k]
                assert e.code < 3 * FC_TOTAL and e.code % FC_TOTAL in (FC_DU, FC_MASKED + FC_DU) and e.okmask == m.okmask and e.ctrl_thr == m.ctrl_thr \
                    and e.ctrl_base == m.ctrl_base
                par = dpar(e)
                f = np.where(par == 1, complex(e.c[2], e.c[3]), complex(e.c[0], e.c[1]))
                if e.flags & MOP_SKIP0:
                    f = np.where(par == 1, f, 1.0)
                acc = acc * f
            for K in range(NV):
                sel = act & okm[:, K]
                regs[sel, K] *= acc[sel]
            i += cnt + 1
            continue
        if code == FC_LX:
            assert not masked and m.okmask == 0xFFFF
            vgrp[act] ^= m.a_thr
            jl_cur[act] ^= 1 << (m.a_reg & 0xFF)
            i += 1
            continue
        if code == FC_LI:
            assert not masked and m.okmask == 0xFFFF
            ib[act] ^= 1 << (m.a_reg & 3)
            i += 1
            continue
        okm = okmask(m, masked)
        if code < FC_DS:                                       # pair forms on slot s
            s = code & 3
            assert (m.code & 3) == s, "engine.h invariant: the op loop takes the slot's inversion byte from code & 3"
            form = code - s
            bit = 1 << s
            inv = (ib >> s) & 1
            c0 = np.where(inv == 1, m.alt[0], m.c[0])
            c1 = np.where(inv == 1, m.alt[1], m.c[1])
            c2 = np.where(inv == 1, m.alt[2], m.c[2])
            c3 = np.where(inv == 1, m.alt[3], m.c[3])
            for K in range(NV):
                if K & bit:
                    continue
                sel = act & okm[:, K]
                assert np.array_equal(okm[:, K], okm[:, K | bit])
                p0, p1 = regs[sel, K].copy(), regs[sel, K | bit].copy()
                if form == FC_PR:
                    n0, n1 = c0[sel] * p0 + c1[sel] * p1, c2[sel] * p0 + c3[sel] * p1
                else:
                    assert form == FC_PX
                    n0, n1 = c0[sel] * p0 - 1j * c1[sel] * p1, -1j * c2[sel] * p0 + c3[sel] * p1
                regs[sel, K], regs[sel, K | bit] = n0, n1
        elif code >= FC_SW:
            bit = 1 << (code - FC_SW)
            for K in range(NV):
                if K & bit:
                    continue
                sel = act & okm[:, K]
                assert np.array_equal(okm[:, K], okm[:, K | bit])
                p0, p1 = regs[sel, K].copy(), regs[sel, K | bit].copy()
                regs[sel, K], regs[sel, K | bit] = p1, p0
        elif code == FC_DU:
            par = dpar(m)
            f = np.where(par == 1, complex(m.c[2], m.c[3]), complex(m.c[0], m.c[1]))
            sel0 = act.copy()
            if m.flags & MOP_SKIP0:
                sel0 &= par == 1
            for K in range(NV):
                sel = sel0 & okm[:, K]
                regs[sel, K] *= f[sel]
        elif FC_DS <= code < FC_DU:
            s = code - FC_DS
            assert (m.code & 3) == s, "engine.h invariant: the op loop takes the slot's inversion byte from code & 3"
            bit = 1 << s
            par = dpar(m)
            blk = ((ib >> s) & 1) ^ par                         # 1: the alt block (roles exchanged)
            f0 = np.where(blk == 1, complex(m.alt[0], m.alt[1]), complex(m.c[0], m.c[1]))
            f1 = np.where(blk == 1, complex(m.alt[2], m.alt[3]), complex(m.c[2], m.c[3]))
            skip0 = bool(m.flags & MOP_SKIP0)
            for K in range(NV):
                if K & bit:
                    sel = act & okm[:, K] & ((not skip0) | (blk == 0))
                    regs[sel, K] *= f1[sel]
                else:
                    sel = act & okm[:, K] & ((not skip0) | (blk == 1))
                    regs[sel, K] *= f0[sel]
        elif code == FC_DG:
            a_reg = m.a_reg & 0xF
            par = (dpar(m) + _popc(ib & a_reg)) & 1
            for K in range(NV):
                sel = act & okm[:, K]
                pk = (par + _pc(K & a_reg)) & 1
                f = np.where(pk == 1, complex(m.c[2], m.c[3]), complex(m.c[0], m.c[1]))
                regs[sel, K] *= f[sel]
        else:
            raise AssertionError(f"unknown fast code {m.code}")
        i += 1
    kk = np.arange(NV)[None, :] ^ ib[:, None]                   # register K holds slot pattern K ^ ib
    dst = jl_cur[:, None] | kbits[kk]
    assert np.unique(dst).size == dst.size, "a stage's stores must cover every tile slot exactly once"
    tile[dst] = regs


def run_tile_pass(psi: np.ndarray, p, rank: int = 0, n_local: int = 64, src: np.ndarray = None):
    """All tiles of one rank's pass `p` (plan.PlanPass, fast interpreter) on the GLOBAL state (indexed
    by PHYSICAL index bits).  Tiles are read from `src` (default: psi itself) and written to psi.
    A remap pass (p.remap) writes both halves of every tile into the rank's own shard, the value of
    the rank bit rg going to local bit rb -- the caller must then read every rank's tiles from a
    snapshot taken before the pass (on the GPU a per-tile handshake orders the two)."""
    T = p.T
    src = psi if src is None else src
    jj = np.arange(1 << T, dtype=np.int64)
    scat = np.zeros(1 << T, dtype=np.int64)
    scat_st = np.zeros(1 << T, dtype=np.int64)
    for l, g in enumerate(p.gpos):
        scat |= ((jj >> l) & 1) << g
        scat_st |= ((jj >> l) & 1) << (p.rb if (p.remap and g == p.rg) else g)
    for t in range(p.n_tiles):
        base = t
        for pos in p.fx_pos:
            base = ((base >> pos) << (pos + 1)) | (base & ((1 << pos) - 1))
        gbase = base | p.fx_val | p.base_or
        idx = gbase | scat
        tile = src[idx].copy()
        for st in p.stages:
            run_stage(tile, st, T, gbase)
        if p.remap:
            sbase = ((base | p.fx_val) & ~(1 << p.rb)) | (rank << n_local)
            psi[sbase | scat_st] = tile
        else:
            psi[idx] = tile


def to_logical(psi: np.ndarray, perm) -> np.ndarray:
    """State indexed by index bits (qubit q at bit perm[q]) -> indexed by qubits."""
    n = len(perm)
    x = np.arange(1 << n, dtype=np.int64)
    y = np.zeros(1 << n, dtype=np.int64)
    for q in range(n):
        y |= ((x >> q) & 1) << perm[q]
    return psi[y]


def to_physical(psi_l: np.ndarray, perm) -> np.ndarray:
    n = len(perm)
    x = np.arange(1 << n, dtype=np.int64)
    y = np.zeros(1 << n, dtype=np.int64)
    for q in range(n):
        y |= ((x >> q) & 1) << perm[q]
    out = np.empty_like(psi_l)
    out[y] = psi_l
    return out
