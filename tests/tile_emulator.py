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

FC_PR, FC_PX, FC_SW, FC_DU, FC_DS, FC_DG, FC_PA, FC_LX, FC_DM, FC_ALL = 0, 4, 8, 12, 13, 17, 18, 22, 23, 24
MOP_SKIP0 = 0x02
NV = 16
U64 = (1 << 64) - 1


def _popc(x):
    return np.bitwise_count(np.asarray(x, dtype=np.uint64)).astype(np.int64)


def _pc(x: int) -> int:
    return bin(x & U64).count("1")


def run_stage(tile: np.ndarray, st, T: int, gbase: int):
    """One stage on one tile (in place).  `gbase`: global index of the tile's first amplitude
    (tile bits clear, rank bits included) -- what the kernel's prepare() derives the flags from."""
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
    mops = st.mops
    i = 0
    while i < len(mops):
        m = mops[i]
        code = m.code
        ok = m.okmask
        if code >= FC_ALL:
            code -= FC_ALL
            assert m.okmask == 0xFFFF, "ALL arms ignore okmask: the planner must only use them without slot controls"
            ok = 0xFFFF
        okb = (~gbase & m.ctrl_base & U64) == 0
        par_base = _pc(gbase & m.a_base) & 7
        act = np.full(G, okb) & ((~vgrp & m.ctrl_thr) == 0)
        c0, c1, c2, c3 = m.c
        if code == FC_DM:
            cnt = m.a_reg
            acc = np.ones(G, dtype=np.complex128)
            for k in range(1, cnt + 1):
                e = mops[i + k]
                assert e.code in (FC_DU, FC_ALL + FC_DU) and e.okmask == m.okmask and e.ctrl_thr == m.ctrl_thr \
                    and e.ctrl_base == m.ctrl_base
                par = (_popc(vgrp & e.a_thr) + (_pc(gbase & e.a_base) & 7)) & 1
                f = np.where(par == 1, complex(e.c[2], e.c[3]), complex(e.c[0], e.c[1]))
                if e.flags & MOP_SKIP0:
                    f = np.where(par == 1, f, 1.0)
                acc = acc * f
            for K in range(NV):
                if ok >> K & 1:
                    regs[act, K] *= acc[act]
            i += cnt + 1
            continue
        if code == FC_LX:
            vgrp[act] ^= m.a_thr
            jl_cur[act] ^= 1 << (m.a_reg & 0xFF)
            i += 1
            continue
        par = (_popc(vgrp & m.a_thr) + par_base) & 1
        if FC_PR <= code < FC_DU or FC_PA <= code < FC_LX:
            s = code & 3 if code < FC_DU else (code - FC_PA)
            form = code - s
            bit = 1 << s
            for K in range(NV):
                if K & bit or not (ok >> K & 1):
                    continue
                p0, p1 = regs[act, K].copy(), regs[act, K | bit].copy()
                if form == FC_PR:
                    n0, n1 = c0 * p0 + c1 * p1, c2 * p0 + c3 * p1
                elif form == FC_PX:
                    n0, n1 = c0 * p0 - 1j * c1 * p1, -1j * c2 * p0 + c3 * p1
                elif form == FC_PA:
                    n0, n1 = (p0 + p1) * c0, (p0 - p1) * c0
                else:
                    assert form == FC_SW
                    n0, n1 = p1, p0
                regs[act, K], regs[act, K | bit] = n0, n1
        elif code == FC_DU:
            f = np.where(par == 1, complex(c2, c3), complex(c0, c1))
            sel = act.copy()
            if m.flags & MOP_SKIP0:
                sel &= par == 1
            for K in range(NV):
                if ok >> K & 1:
                    regs[sel, K] *= f[sel]
        elif FC_DS <= code < FC_DG:
            bit = 1 << (code - FC_DS)
            f0 = np.where(par == 1, complex(c2, c3), complex(c0, c1))      # roles swap with the outer parity
            f1 = np.where(par == 1, complex(c0, c1), complex(c2, c3))
            skip0 = bool(m.flags & MOP_SKIP0) & (par == 0)
            for K in range(NV):
                if not (ok >> K & 1):
                    continue
                if K & bit:
                    regs[act, K] *= f1[act]
                else:
                    sel = act & ~skip0
                    regs[sel, K] *= f0[sel]
        elif code == FC_DG:
            a_reg = m.a_reg & 0xF
            for K in range(NV):
                if not (ok >> K & 1):
                    continue
                pk = (par + _pc(K & a_reg)) & 1
                f = np.where(pk == 1, complex(c2, c3), complex(c0, c1))
                regs[act, K] *= f[act]
        else:
            raise AssertionError(f"unknown fast code {m.code}")
        i += 1
    dst = jl_cur[:, None] | kbits[None, :]
    assert np.unique(dst).size == dst.size, "a stage's stores must cover every tile slot exactly once"
    tile[dst] = regs


def run_tile_pass(psi: np.ndarray, p):
    """All tiles of one rank's pass `p` (plan.PlanPass, fast interpreter) on the GLOBAL state."""
    T = p.T
    jj = np.arange(1 << T, dtype=np.int64)
    scat = np.zeros(1 << T, dtype=np.int64)
    for l, g in enumerate(p.gpos):
        scat |= ((jj >> l) & 1) << g
    for t in range(p.n_tiles):
        base = t
        for pos in p.fx_pos:
            base = ((base >> pos) << (pos + 1)) | (base & ((1 << pos) - 1))
        gbase = base | p.fx_val | p.base_or
        idx = gbase | scat
        tile = psi[idx].copy()
        for st in p.stages:
            run_stage(tile, st, T, gbase)
        psi[idx] = tile
