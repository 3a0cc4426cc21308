#!/usr/bin/env python3
"""Generates fastops_ptx.inc: the op loop of the FAST stage interpreter (tile.cu) as ONE inline-PTX
block -- header fetch, a single `brx.idx` jump table and every arm -- in two flavours: QV_FASTOPS_PTX
(lean) and QV_FASTOPS_PTX_SC (plus the single-control arms; see gen()).

Why PTX: the C++ switch compiles to a compare tree plus divergence bookkeeping (BSSY / BSYNC / BREAK),
~40 instructions and several dependent branches per op around 64 FP64 instructions of work.  Written
by hand, an unconditional op costs: the (prefetched) header, two coefficient loads, the jump-table
load and the indirect branch -- ~17 instructions -- and the FP64 body.  Everything an op only
SOMETIMES needs is kept off that path:
  * controls on thread bits: the op's code carries a class (code = arm + 44 * cls), class 1 jumps to a
    stub that tests the controls and re-dispatches (or skips the op);
  * controls outside the tile (class 2) are the same for the whole tile: before a tile's first stage
    the kernel rewrites those code bytes into the class 0 / 1 code or a skip code (tile.cu, patch_codes),
    so the loop never sees them;
  * controls in register slots (okmask): the masked arms are reached through a stub that builds the
    predicate mask (and permutes it by the inverted slots);
  * diagonal ops with target bits outside the register slots, merged runs, lazy x: the arm re-reads the
    header words it needs.
INVARIANT the prologue relies on: a code whose arm works on ONE register slot j has code & 3 == j in every
variant (masked, class 1, single-control, butterfly) -- `blk`, the inversion byte of that slot, picks the
coefficient block.  tests/test_tile_emulator.py checks the numbering against engine.h.

Operands of the asm statement (tile.cu, stage_ops_fast_ptx):
  %0..%31  the 16 register-resident amplitudes, v[K].x = %(2K), v[K].y = %(2K+1)      ("+d")
  %32 vgrp  %33 inv  %34 jl  %35 mine_o ("+r")   %36 goff ("+l")
  %37 first op (shared-window address; a sentinel descriptor ends the stage)  %38 unused  %39 this tile's flag bytes
  %40 this thread's column of the tabulated-run table  %41 this tile's outside-the-tile factors
  %42 byte stride between runs in the table ("r")

Descriptor layout: engine.h MOp (80 bytes): w0 = code | flags << 8 | okmask << 16, ctrl_thr, a_thr,
w3 = a_reg | idx << 16, c0..c3 at +16, alt block at +48.  Codes: engine.h FCode.
"""
import sys

NV = 16
FC_PR, FC_PX, FC_DS, FC_DU, FC_DG, FC_LX, FC_LI, FC_DM, FC_MASKED, FC_SW, FC_TOTAL = 0, 4, 8, 12, 13, 14, 15, 16, 20, 40, 44
FC_DS1, FC_DU1, FC_DM1 = 132, 148, 152      # one control in register slot c, no other control (engine.h)
FC_HB = 156                                 # butterfly h on slot j (engine.h)
SKIP0, COND, PARB, CONDB, ATHR, STATIC = 0x02 << 8, 0x04 << 8, 0x08 << 8, 0x10 << 8, 0x20 << 8, 0x40 << 8
MOP = 80

# Code layout: the arms a random circuit keeps jumping between (pairs, unmasked diagonals, lazy x, the
# controlled swaps) are emitted first and stay together behind the prologue (the SM's instruction
# cache holds 32 KB); then the arms of qft-like circuits; the generic predicated arms come last.
SECT = {"hot": [], "warm": [], "cold": []}
CUR = ["hot"]


def e(s=""):
    SECT[CUR[0]].append(s)


def sect(name):
    CUR[0] = name


def X(K):
    return f"%{2 * K}"


def Y(K):
    return f"%{2 * K + 1}"


def pred(K, masked):
    """emit the predicate of slot pattern K for masked arms; returns the guard prefix"""
    if not masked:
        return ""
    e(f"and.b32 u, ok, {1 << K};")
    e("setp.ne.u32 pk, u, 0;")
    return "@pk "


def pair_real(K0, K1, g):
    x0, y0, x1, y1 = X(K0), Y(K0), X(K1), Y(K1)
    e(f"{g}mul.rn.f64 ta, c0, {x0};")
    e(f"{g}mul.rn.f64 tb, c2, {x0};")
    e(f"{g}mul.rn.f64 tc, c0, {y0};")
    e(f"{g}mul.rn.f64 td, c2, {y0};")
    e(f"{g}fma.rn.f64 {x0}, c1, {x1}, ta;")
    e(f"{g}fma.rn.f64 {y0}, c1, {y1}, tc;")
    e(f"{g}fma.rn.f64 {x1}, c3, {x1}, tb;")
    e(f"{g}fma.rn.f64 {y1}, c3, {y1}, td;")


def pair_cross(K0, K1, g):
    # new0 = c0*p0 - i*c1*p1 ; new1 = -i*c2*p0 + c3*p1   (n1 = -c1, n2 = -c2)
    x0, y0, x1, y1 = X(K0), Y(K0), X(K1), Y(K1)
    e(f"{g}mul.rn.f64 ta, c0, {x0};")
    e(f"{g}mul.rn.f64 tc, c0, {y0};")
    e(f"{g}mul.rn.f64 tb, c2, {y0};")
    e(f"{g}mul.rn.f64 td, n2, {x0};")
    e(f"{g}fma.rn.f64 {x0}, c1, {y1}, ta;")
    e(f"{g}fma.rn.f64 {y0}, n1, {x1}, tc;")
    e(f"{g}fma.rn.f64 {x1}, c3, {x1}, tb;")
    e(f"{g}fma.rn.f64 {y1}, c3, {y1}, td;")


def cmul(K, fr, fi, nfi, g):
    x, y = X(K), Y(K)
    e(f"{g}mul.rn.f64 ta, {nfi}, {y};")
    e(f"{g}mul.rn.f64 tb, {fi}, {x};")
    e(f"{g}fma.rn.f64 {x}, {fr}, {x}, ta;")
    e(f"{g}fma.rn.f64 {y}, {fr}, {y}, tb;")


def swap(K0, K1, g):
    x0, y0, x1, y1 = X(K0), Y(K0), X(K1), Y(K1)
    e(f"{g}mov.f64 ta, {x0};")
    e(f"{g}mov.f64 tb, {y0};")
    e(f"{g}mov.f64 {x0}, {x1};")
    e(f"{g}mov.f64 {y0}, {y1};")
    e(f"{g}mov.f64 {x1}, ta;")
    e(f"{g}mov.f64 {y1}, tb;")


def reload(at="p", off=-MOP):
    """the full header of the op at `at` + off into w0..w3 (arms that need more than w0 / w1)"""
    e(f"ld.shared.v4.u32 {{{{w0, w1, w2, w3}}}}, [{at}+{off}];".replace("{{", "{").replace("}}", "}"))


def dpar():
    """par = (popc(vgrp & a_thr) + (PARB ? flag byte : 0)) & 1 from the header in w0, w2, w3"""
    e("and.b32 t, %32, w2;")
    e("popc.b32 par, t;")
    e(f"and.b32 u, w0, {PARB};")
    e("setp.ne.u32 pk, u, 0;")
    e("@pk shr.u32 u, w3, 16;")
    e("@pk add.u32 u, u, %39;")
    e("@pk ld.shared.u8 fl, [u];")
    e("@pk add.u32 par, par, fl;")
    e("and.b32 par, par, 1;")


def gen(sc):
    """sc: with the single-control arms (FC_DS1 / FC_DU1 / FC_DM1).  The kernel is compiled in both
    flavours (tile.cu, template parameter SC) and a pass without such ops runs the lean one: the arms are
    never executed there, but one asm block is ONE register-allocation problem for ptxas, and with them
    in it the arms of a random circuit came out 3 % slower (A/B on one box, profiles/r02s_ab.txt)."""
    arm = ["$BAD"] * FC_TOTAL            # the arms proper (a code nobody emits traps)
    for j in range(4):
        arm[FC_PR + j] = f"$PR{j}"
        arm[FC_PX + j] = f"$PX{j}"
        arm[FC_DS + j] = f"$DS{j}"
        arm[FC_MASKED + FC_PR + j] = f"$MPR{j}"
        arm[FC_MASKED + FC_PX + j] = f"$MPX{j}"
        arm[FC_MASKED + FC_DS + j] = f"$MDS{j}"
        arm[FC_SW + j] = f"$SW{j}"
    arm[FC_DU] = "$DU"
    arm[FC_MASKED + FC_DU] = "$MDU"
    arm[FC_MASKED + FC_DG] = "$DG"
    arm[FC_LX] = "$LX"
    arm[FC_LI] = "$LI"
    arm[FC_DM] = "$DM"
    arm[FC_MASKED + FC_DM] = "$MDM"
    # first-level table: masked arms go through $MPRE; class 1 (code + 44) through its stub
    t1 = [("$MPRE" if (FC_MASKED <= c and arm[c] != "$BAD") else arm[c]) for c in range(FC_TOTAL)]
    # class 2 (controls outside the tile) never reaches the loop: the kernel rewrites those code bytes
    # for every tile (tile.cu, patch_codes) into the class 0 / 1 code or one of the two skip codes
    tbl = t1 + ["$C1"] * FC_TOTAL + ["$BAD"] * FC_TOTAL
    tbl += ["$BAD"] * (256 - len(tbl))
    for j in range(4):
        tbl[FC_HB + j] = f"$HB{j}"
    for c in range(4):
        for j in range(4):
            if j != c:
                tbl[FC_DS1 + 4 * c + j] = f"$DS{j}c{c}" if sc else "$BAD"      # (code & 3 == j: the prologue's blk)
        tbl[FC_DU1 + c] = f"$DUc{c}" if sc else "$BAD"
        tbl[FC_DM1 + c] = f"$DMc{c}" if sc else "$BAD"
    tbl[253] = "$SKIPRUN"                # engine.h MOP_NOP_RUN: a switched-off run header takes its members with it
    tbl[254] = "$TOP"                    # engine.h MOP_NOP: an op this tile's outside controls switch off
    tbl[255] = "$END"                    # the sentinel descriptor behind every stage (engine.h MOP_END)
    # the lazy x forms are a handful of instructions: they test their thread controls themselves
    # (ctrl_thr is 0 for an unconditional one) instead of paying a second indirect jump
    tbl[FC_TOTAL + FC_LX] = "$LX"
    tbl[FC_TOTAL + FC_LI] = "$LI"

    e("{")
    e(".reg .pred pq, pb, pk, pz, pa;")
    e(".reg .u32 p, w0, w1, w2, w3, h0, h1, code, t, u, blk, ca, ok, fl, cnt, par, lo, hi, areg, ib" + (", dmv" if sc else "") + ";")
    e(".reg .f64 c0, c1, c2, c3, ta, tb, tc, td, n1, n2, ar, ai, fr, fi;")
    e(".reg .u64 g64;")
    e("$TBL: .branchtargets " + ", ".join(tbl) + ";")
    e("$TBM: .branchtargets " + ", ".join(arm) + ";")
    e("mov.u32 p, %37;")
    e("mov.u32 ok, 65535;")
    # The header of op k+1 is fetched while op k runs (h*).  (The fetch behind the stage's last op
    # reads the next descriptor or the table after the op array -- inside the CTA's shared memory.)
    e("ld.shared.v2.u32 {h0, h1}, [p];")
    # ---- per-op prologue ----
    e("$TOP:")
    e("mov.u32 w0, h0;")
    e("mov.u32 w1, h1;")
    # (ptxas sinks this add behind the loads whatever the order here, and the add then waits until the
    # loads have read their address register: 2.4 % of the kernel's samples; a run-time step was no better)
    e(f"add.u32 p, p, {MOP};")
    e("and.b32 code, w0, 255;")
    e("and.b32 t, w0, 3;")                 # the op's register slot is code & 3: its inversion byte (0 / 32)
    e("or.b32 t, t, 0x4440;")              # picks the coefficient block (the alt block while inverted)
    e("prmt.b32 blk, %33, 0, t;")
    e("ld.shared.v2.u32 {h0, h1}, [p];")
    e("add.u32 ca, p, blk;")
    e(f"ld.shared.v2.f64 {{c0, c1}}, [ca+{16 - MOP}];")
    e(f"ld.shared.v2.f64 {{c2, c3}}, [ca+{32 - MOP}];")
    e("brx.idx.uni code, $TBL;")
    # ---- class 1: controls on thread bits ----
    e("$C1:")
    e(f"sub.u32 code, code, {FC_TOTAL};")
    e("not.b32 t, %32;")
    e("and.b32 t, t, w1;")
    e("setp.ne.u32 pk, t, 0;")
    e("@pk bra $SKIP;")
    e("brx.idx.uni code, $TBL;")
    # a skipped run header takes its members with it
    e("$SKIP:")
    e(f"setp.eq.u32 pk, code, {FC_DM};")
    e(f"setp.eq.u32 pb, code, {FC_DM + FC_MASKED};")
    e("or.pred pk, pk, pb;")
    e("@!pk bra $TOP;")
    e("$SKIPRUN:")
    reload()
    e("and.b32 cnt, w3, 65535;")
    e(f"mad.lo.u32 p, cnt, {MOP}, p;")
    e("ld.shared.v2.u32 {h0, h1}, [p];")
    e("bra $TOP;")
    # ---- controls in register slots: okmask, permuted by the inverted slots (register K holds pattern K ^ ib)
    e("$MPRE:")
    e("shr.u32 ok, w0, 16;")
    e("setp.eq.u32 pz, %33, 0;")
    e("@pz bra $MGO;")
    for j, (m, sh) in enumerate([(0x5555, 1), (0x3333, 2), (0x0F0F, 4), (0x00FF, 8)]):
        e(f"and.b32 t, %33, {0xFF << (8 * j)};")
        e("setp.ne.u32 pk, t, 0;")
        e(f"and.b32 u, ok, {m};")
        e(f"shl.b32 u, u, {sh};")
        e(f"shr.u32 t, ok, {sh};")
        e(f"and.b32 t, t, {m};")
        e("or.b32 u, u, t;")
        e("@pk mov.u32 ok, u;")
    e("$MGO:")
    e("brx.idx.uni code, $TBM;")

    # ---- pair arms ----
    for masked in (False, True):
        pre = "M" if masked else ""
        sect("cold" if masked else "hot")
        for j in range(4):
            b = 1 << j
            e(f"${pre}PR{j}:")
            for K in range(NV):
                if K & b:
                    continue
                g = pred(K, masked)
                pair_real(K, K | b, g)
            e("bra.uni $TOP;")
            e(f"${pre}PX{j}:")
            e("neg.f64 n1, c1;")
            e("neg.f64 n2, c2;")
            for K in range(NV):
                if K & b:
                    continue
                g = pred(K, masked)
                pair_cross(K, K | b, g)
            e("bra.uni $TOP;")
    sect("hot")
    for j in range(4):
        b = 1 << j
        e(f"$SW{j}:")
        for K in range(NV):
            if K & b:
                continue
            g = pred(K, True)
            swap(K, K | b, g)
        e("bra.uni $TOP;")

    # ---- butterfly h on slot j: (p0 + p1, p0 - p1); with the slot inverted (blk != 0) register K holds p1 ----
    # Two FP64 instructions per component and no move: sum = p0 + p1 in place, then the difference as
    # sum - 2 * p1 in p1's register (one extra rounding of the sum, ~1e-16 relative).
    sect("hot")
    for j in range(4):
        b = 1 << j
        e(f"$HB{j}:")
        e("mov.f64 ta, 0dC000000000000000;")          # -2.0
        e("setp.ne.u32 pk, blk, 0;")
        e(f"@pk bra $HB{j}i;")
        for inv in (False, True):
            if inv:
                e(f"$HB{j}i:")
            for K in range(NV):
                if K & b:
                    continue
                lo_, hi_ = (K | b, K) if inv else (K, K | b)        # lo_ holds p0, hi_ holds p1
                for comp in (X, Y):
                    e(f"add.rn.f64 {comp(lo_)}, {comp(lo_)}, {comp(hi_)};")
                    e(f"fma.rn.f64 {comp(hi_)}, {comp(hi_)}, ta, {comp(lo_)};")
            e("bra.uni $TOP;")

    # ---- diagonal, one target bit in register slot j ----
    def ds_arm(lab, j, masked, only=None):
        """only = c: the single-control form -- slot patterns with bit c set, no predicates"""
        b = 1 << j
        keep = (lambda K: True) if only is None else (lambda K: bool(K & (1 << only)))
        e(f"${lab}:")
        if only is not None:
            # the control's slot inverted by a lazy x: the generic masked arm permutes the predicates
            e(f"and.b32 t, %33, {0xFF << (8 * only)};")
            e("setp.ne.u32 pk, t, 0;")
            e(f"@pk bra $FBDS{j};")
        # target bits outside the register slots (rzz's second bit, bits outside the tile): their
        # parity exchanges the roles of f0 and f1 -- the other coefficient block
        e(f"and.b32 u, w0, {PARB | ATHR};")
        e("setp.eq.u32 pk, u, 0;")
        e(f"@pk bra.uni ${lab}g;")
        reload()
        dpar()
        e("setp.eq.u32 pk, par, 0;")
        e(f"@pk bra ${lab}g;")
        e("xor.b32 blk, blk, 32;")
        e(f"sub.u32 ca, p, {MOP};")
        e("add.u32 ca, ca, blk;")
        e("ld.shared.v2.f64 {c0, c1}, [ca+16];")
        e("ld.shared.v2.f64 {c2, c3}, [ca+32];")
        e(f"${lab}g:")
        # MOP_SKIP0: the ORIGINAL f0 is 1 -- with the roles exchanged it sits in (c2, c3)
        e(f"and.b32 u, w0, {SKIP0};")
        e("setp.ne.u32 pk, u, 0;")
        e("setp.eq.u32 pz, blk, 0;")
        e("and.pred pa, pk, pz;")
        e(f"@pa bra ${lab}b;")
        e("neg.f64 n1, c1;")
        for K in range(NV):
            if (K & b) or not keep(K):
                continue
            g = pred(K, masked)
            cmul(K, "c0", "c1", "n1", g)
        e(f"${lab}b:")
        e(f"and.b32 u, w0, {SKIP0};")
        e("setp.ne.u32 pk, u, 0;")
        e("setp.ne.u32 pz, blk, 0;")
        e("and.pred pa, pk, pz;")
        e("@pa bra $TOP;")
        e("neg.f64 n2, c3;")
        for K in range(NV):
            if not (K & b) or not keep(K):
                continue
            g = pred(K, masked)
            cmul(K, "c2", "c3", "n2", g)
        e("bra.uni $TOP;")

    for masked in (False, True):
        sect("cold" if masked else "hot")
        for j in range(4):
            ds_arm(("M" if masked else "") + f"DS{j}", j, masked)
    sect("warm")
    for j in range(4 if sc else 0):
        for c in range(4):
            if c != j:
                ds_arm(f"DS{j}c{c}", j, False, only=c)
        e(f"$FBDS{j}:")
        e(f"mov.u32 code, {FC_MASKED + FC_DS + j};")
        e("bra $MPRE;")

    # ---- diagonal, no target bit in a register slot: one factor per thread ----
    def du_arm(lab, masked, only=None):
        e(f"${lab}:")
        if only is not None:
            e(f"and.b32 t, %33, {0xFF << (8 * only)};")
            e("setp.ne.u32 pk, t, 0;")
            e("@pk bra $FBDU;")
        reload()
        dpar()
        e(f"and.b32 u, w0, {SKIP0};")
        e("setp.ne.u32 pk, u, 0;")
        e("setp.eq.u32 pz, par, 0;")
        e("and.pred pa, pk, pz;")
        e("@pa bra $TOP;")
        e("shl.b32 t, par, 4;")
        e(f"sub.u32 ca, p, {MOP - 16};")
        e("add.u32 ca, ca, t;")
        e("ld.shared.v2.f64 {fr, fi}, [ca];")
        e("neg.f64 n1, fi;")
        for K in range(NV):
            if only is not None and not (K & (1 << only)):
                continue
            g = pred(K, masked)
            cmul(K, "fr", "fi", "n1", g)
        e("bra.uni $TOP;")

    sect("hot")
    du_arm("DU", False)
    sect("cold")
    du_arm("MDU", True)
    sect("warm")
    for c in range(4 if sc else 0):
        du_arm(f"DUc{c}", False, only=c)
    if sc:
        e("$FBDU:")
        e(f"mov.u32 code, {FC_MASKED + FC_DU};")
        e("bra $MPRE;")

    # ---- diagonal, any set of target bits in register slots (always dispatched as masked) ----
    sect("cold")
    e("$DG:")
    reload()
    dpar()
    e("and.b32 areg, w3, 15;")
    e("shr.u32 ib, %33, 5;")
    e("and.b32 ib, ib, 1;")
    for j, sh in ((1, 12), (2, 19), (3, 26)):
        e(f"shr.u32 t, %33, {sh};")
        e(f"and.b32 t, t, {1 << j};")
        e("or.b32 ib, ib, t;")
    e("and.b32 t, ib, areg;")          # inverted target slots add their parity
    e("popc.b32 t, t;")
    e("add.u32 par, par, t;")
    for K in range(NV):
        e(f"and.b32 t, areg, {K};")
        e("popc.b32 t, t;")
        e("add.u32 t, t, par;")
        e("and.b32 t, t, 1;")
        e("setp.ne.u32 pz, t, 0;")
        e("selp.f64 fr, c2, c0, pz;")
        e("selp.f64 fi, c3, c1, pz;")
        e("neg.f64 n1, fi;")
        g = pred(K, True)
        cmul(K, "fr", "fi", "n1", g)
    e("bra.uni $TOP;")

    sect("hot")
    # ---- lazy x: target on a thread bit (c0 = smem delta | bit << 32, c1 = shard-offset delta,
    #      c2 = the target's thread bit) ----
    e("$LX:")
    e("not.b32 t, %32;")
    e("and.b32 t, t, w1;")
    e("setp.eq.u32 pa, t, 0;")             # controls on thread bits satisfied
    e("mov.b64 {lo, hi}, c2;")
    e("@pa xor.b32 %32, %32, lo;")
    e("mov.b64 {lo, hi}, c0;")
    e("and.b32 t, %34, hi;")
    e("setp.ne.u32 pk, t, 0;")
    e("and.pred pz, pa, pk;")
    e("not.pred pk, pk;")
    e("and.pred pk, pa, pk;")
    e("@pz sub.u32 %35, %35, lo;")
    e("@pk add.u32 %35, %35, lo;")
    e("@pa xor.b32 %34, %34, hi;")
    e("mov.b64 g64, c1;")
    e("@pa xor.b64 %36, %36, g64;")
    e("bra.uni $TOP;")
    # ---- lazy x: target in register slot (c0's low word = the slot's inversion byte, 32 << 8 * slot) ----
    e("$LI:")
    e("not.b32 t, %32;")
    e("and.b32 t, t, w1;")
    e("setp.eq.u32 pa, t, 0;")
    e("mov.b64 {lo, hi}, c0;")
    e("@pa xor.b32 %33, %33, lo;")
    e("bra.uni $TOP;")

    # ---- merged diagonal run: header + cnt members (FC_DU forms sharing the header's controls) ----
    if not sc:
        # lean flavour: one copy per arm, no variant register
        for masked in (False, True):
            pre = "M" if masked else ""
            sect("cold" if masked else "warm")
            e(f"${pre}DM:")
            reload()
            e("and.b32 cnt, w3, 65535;")
            # tabulated run (MOP_STATIC): one load for the thread-bit members, one for the members outside the tile
            e(f"and.b32 u, w0, {STATIC};")
            e("setp.eq.u32 pk, u, 0;")
            e(f"@pk bra.uni ${pre}DMd;")
            e("mad.lo.u32 ca, w2, %42, %40;")
            e("ld.shared.v2.f64 {ar, ai}, [ca];")
            e(f"mad.lo.u32 p, cnt, {MOP}, p;")
            e(f"and.b32 u, w0, {PARB};")
            e("setp.eq.u32 pk, u, 0;")
            e(f"@pk bra.uni ${pre}DMa;")
            e("shl.b32 t, w2, 4;")
            e("add.u32 ca, t, %41;")
            e("ld.shared.v2.f64 {fr, fi}, [ca];")
            e("mul.rn.f64 ta, ai, fi;")
            e("mul.rn.f64 tb, ai, fr;")
            e("neg.f64 ta, ta;")
            e("fma.rn.f64 ta, ar, fr, ta;")
            e("fma.rn.f64 ai, ar, fi, tb;")
            e("mov.f64 ar, ta;")
            e(f"bra.uni ${pre}DMa;")
            e(f"${pre}DMd:")
            e("mov.f64 ar, 0d3FF0000000000000;")
            e("mov.f64 ai, 0d0000000000000000;")
            e(f"${pre}DMl:")
            e("setp.eq.u32 pq, cnt, 0;")
            e(f"@pq bra.uni ${pre}DMa;")
            e("ld.shared.v4.u32 {w0, w1, w2, w3}, [p];")
            dpar()
            e(f"and.b32 u, w0, {SKIP0};")
            e("setp.ne.u32 pk, u, 0;")
            e("setp.eq.u32 pz, par, 0;")
            e("and.pred pa, pk, pz;")
            e(f"@pa bra ${pre}DMn;")
            e("shl.b32 t, par, 4;")
            e("add.u32 ca, p, t;")
            e("ld.shared.v2.f64 {fr, fi}, [ca+16];")
            e("mul.rn.f64 ta, ai, fi;")            # (ar, ai) *= (fr, fi)
            e("mul.rn.f64 tb, ai, fr;")
            e("neg.f64 ta, ta;")
            e("fma.rn.f64 ta, ar, fr, ta;")
            e("fma.rn.f64 ai, ar, fi, tb;")
            e("mov.f64 ar, ta;")
            e(f"${pre}DMn:")
            e(f"add.u32 p, p, {MOP};")
            e("sub.u32 cnt, cnt, 1;")
            e(f"bra.uni ${pre}DMl;")
            e(f"${pre}DMa:")
            e("ld.shared.v2.u32 {h0, h1}, [p];")
            e("neg.f64 n1, ai;")
            for K in range(NV):
                g = pred(K, masked)
                cmul(K, "ar", "ai", "n1", g)
            e("bra.uni $TOP;")
    else:
        sect("warm")
        # One copy of the factor code for every variant; dmv says how the factor is applied at the end:
        # 0 every slot pattern, 1 under the okmask predicates, 2 + c the patterns with slot bit c set.
        e("$DM:")
        e("mov.u32 dmv, 0;")
        e("bra.uni $DMgo;")
        for c in range(4):
            e(f"$DMc{c}:")
            e(f"and.b32 t, %33, {0xFF << (8 * c)};")
            e("setp.ne.u32 pk, t, 0;")
            e("@pk bra $FBDM;")
            e(f"mov.u32 dmv, {2 + c};")
            e("bra.uni $DMgo;")
        e("$FBDM:")
        e(f"mov.u32 code, {FC_MASKED + FC_DM};")
        e("bra $MPRE;")
        e("$MDM:")
        e("mov.u32 dmv, 1;")
        e("$DMgo:")
        reload()
        e("and.b32 cnt, w3, 65535;")
        # tabulated run (MOP_STATIC): one load for the thread-bit members, one for the members outside the tile
        e(f"and.b32 u, w0, {STATIC};")
        e("setp.eq.u32 pk, u, 0;")
        e("@pk bra.uni $DMd;")
        e("mad.lo.u32 ca, w2, %42, %40;")
        e("ld.shared.v2.f64 {ar, ai}, [ca];")
        e(f"mad.lo.u32 p, cnt, {MOP}, p;")
        e(f"and.b32 u, w0, {PARB};")
        e("setp.eq.u32 pk, u, 0;")
        e("@pk bra.uni $DMa;")
        e("shl.b32 t, w2, 4;")
        e("add.u32 ca, t, %41;")
        e("ld.shared.v2.f64 {fr, fi}, [ca];")
        e("mul.rn.f64 ta, ai, fi;")
        e("mul.rn.f64 tb, ai, fr;")
        e("neg.f64 ta, ta;")
        e("fma.rn.f64 ta, ar, fr, ta;")
        e("fma.rn.f64 ai, ar, fi, tb;")
        e("mov.f64 ar, ta;")
        e("bra.uni $DMa;")
        e("$DMd:")
        e("mov.f64 ar, 0d3FF0000000000000;")
        e("mov.f64 ai, 0d0000000000000000;")
        e("$DMl:")
        e("setp.eq.u32 pq, cnt, 0;")
        e("@pq bra.uni $DMa;")
        e("ld.shared.v4.u32 {w0, w1, w2, w3}, [p];")
        dpar()
        e(f"and.b32 u, w0, {SKIP0};")
        e("setp.ne.u32 pk, u, 0;")
        e("setp.eq.u32 pz, par, 0;")
        e("and.pred pa, pk, pz;")
        e("@pa bra $DMn;")
        e("shl.b32 t, par, 4;")
        e("add.u32 ca, p, t;")
        e("ld.shared.v2.f64 {fr, fi}, [ca+16];")
        e("mul.rn.f64 ta, ai, fi;")            # (ar, ai) *= (fr, fi)
        e("mul.rn.f64 tb, ai, fr;")
        e("neg.f64 ta, ta;")
        e("fma.rn.f64 ta, ar, fr, ta;")
        e("fma.rn.f64 ai, ar, fi, tb;")
        e("mov.f64 ar, ta;")
        e("$DMn:")
        e(f"add.u32 p, p, {MOP};")
        e("sub.u32 cnt, cnt, 1;")
        e("bra.uni $DMl;")
        e("$DMa:")
        e("ld.shared.v2.u32 {h0, h1}, [p];")
        e("neg.f64 n1, ai;")
        e("setp.eq.u32 pk, dmv, 0;")
        e("@pk bra $DMA0;")
        e("setp.eq.u32 pk, dmv, 1;")
        e("@pk bra $DMA1;")
        for c in range(3):
            e(f"setp.eq.u32 pk, dmv, {2 + c};")
            e(f"@pk bra $DMAc{c};")
        for c in (3, 2, 1, 0):
            e(f"$DMAc{c}:")
            for K in range(NV):
                if K & (1 << c):
                    cmul(K, "ar", "ai", "n1", "")
            e("bra.uni $TOP;")
        e("$DMA0:")
        for K in range(NV):
            cmul(K, "ar", "ai", "n1", "")
        e("bra.uni $TOP;")
        sect("cold")
        e("$DMA1:")
        for K in range(NV):
            g = pred(K, True)
            cmul(K, "ar", "ai", "n1", g)
        e("bra.uni $TOP;")

    sect("cold")
    e("$BAD:")                      # a code without an arm (a single-control code in the lean flavour, an unpatched
    e("trap;")                      # class-2 code): a planner / launch-selection bug -- fail loudly
    e("$END:")
    e("}")
    out = SECT["hot"] + SECT["warm"] + SECT["cold"]
    for k in SECT:
        SECT[k] = []
    CUR[0] = "hot"
    return out


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "fastops_ptx.inc"
    with open(out, "w") as f:
        f.write("// GENERATED by gen_fastops.py -- do not edit; the op loop of the FAST stage interpreter.\n")
        for name, sc in (("QV_FASTOPS_PTX", False), ("QV_FASTOPS_PTX_SC", True)):
            lines = gen(sc)
            f.write(f"#define {name} \\\n")
            for i, s in enumerate(lines):
                s = s.replace("$", "QF_")
                f.write(f'    "{s}\\n\\t"' + (" \\\n" if i + 1 < len(lines) else "\n"))
            print(f"{out}: {name} {len(lines)} PTX lines")


if __name__ == "__main__":
    main()
