"""Host-side view of the scheduler: which passes / stages qvnt_reg_apply would run
(`qvnt_plan_describe`, include/qvnt_b200.h).  Needs no GPU."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_size_t
from dataclasses import dataclass, field
from typing import List, Optional

from . import _ffi
from .op import MultiOp, SingleOp


@dataclass
class PlanOp:
    src: int
    kind: int
    dagger: int
    a: int
    b: int
    ctrl: int
    form: int
    ra: int
    rb: int
    scale: float = 0.0      # h1: the butterfly factor (1/sqrt2; halves of a split h2: 1 and 0.5)
    pa: int = 0             # the same masks as index bits at scheduling time (differ after a remap pass)
    pb: int = 0
    pctrl: int = 0


@dataclass
class PlanMop:
    """One encoded micro-op of a stage, as the kernel reads it (engine.h: MOp + MBase)."""
    code: int
    flags: int
    okmask: int
    ctrl_thr: int
    a_thr: int
    a_reg: int
    c: List[float]
    ctrl_base: int
    a_base: int
    src: int = -1
    alt: List[float] = field(default_factory=list)
    idx: int = 0


@dataclass
class PlanStage:
    r_lpos: List[int]
    t_lpos: List[int]
    ops: List[PlanOp] = field(default_factory=list)
    mops: List[PlanMop] = field(default_factory=list)
    sync: int = 0


@dataclass
class PlanPass:
    direct: bool
    T: int = 0
    L: int = 0
    n_tiles: int = 0
    base_or: int = 0
    peer: int = 0
    full: int = 0
    fx_val: int = 0
    remap: int = 0          # the pass leaves rank bit rg and local bit rb exchanged (planner.cu)
    rg: int = 0
    rb: int = 0
    gpos: List[int] = field(default_factory=list)
    fx_pos: List[int] = field(default_factory=list)
    stages: List[PlanStage] = field(default_factory=list)
    op: Optional[PlanOp] = None

    def all_ops(self) -> List[PlanOp]:
        return [self.op] if self.direct else [o for s in self.stages for o in s.ops]


def _ints(s: str) -> List[int]:
    return [int(x) for x in s.split(",") if x != ""]


def describe(q_num: int, circ, rank: int = 0, world: int = 1, peers: bool = False, fuse: bool = True,
             tile_bits: int = 0, chunk_bits: int = 0, remap: bool = True,
             lower_two_bit: bool = False) -> List[PlanPass]:
    if isinstance(circ, SingleOp):
        circ = MultiOp([circ])
    arr, n = circ.to_c_array()
    need = c_size_t(0)
    lib = _ffi.lib()
    args = (q_num, rank, world, (1 if remap else 3) if peers else 0, (3 if lower_two_bit else 1) if fuse else 0,
            tile_bits, chunk_bits, arr, n)
    _ffi.check(lib.qvnt_plan_describe(*args, None, 0, byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    _ffi.check(lib.qvnt_plan_describe(*args, buf, need.value, byref(need)))
    passes: List[PlanPass] = []
    last_src = -1
    for line in buf.value.decode().splitlines():
        tok = line.split()
        kv = dict(t.split("=", 1) for t in tok if "=" in t)
        if tok[0] == "pass":
            if tok[1] == "direct":
                passes.append(PlanPass(direct=True))
            else:
                passes.append(PlanPass(direct=False, T=int(kv["T"]), L=int(kv["L"]), n_tiles=int(kv["n_tiles"]),
                                       base_or=int(kv["base_or"]), peer=int(kv["peer"]), full=int(kv.get("full", 0)),
                                       fx_val=int(kv["fx_val"]), remap=int(kv.get("remap", 0)),
                                       rg=int(kv.get("rg", 0)), rb=int(kv.get("rb", 0)),
                                       gpos=_ints(kv["gpos"]), fx_pos=_ints(kv["fx_pos"])))
        elif tok[0] == "stage":
            passes[-1].stages.append(PlanStage(r_lpos=_ints(kv["r"]), t_lpos=_ints(kv["t"]),
                                               sync=int(kv.get("sync", 0))))
        elif tok[0] == "mop":
            passes[-1].stages[-1].mops.append(PlanMop(
                code=int(kv["code"]), flags=int(kv["flags"]), okmask=int(kv["okmask"]), ctrl_thr=int(kv["ctrl_thr"]),
                a_thr=int(kv["a_thr"]), a_reg=int(kv["a_reg"]),
                c=[float.fromhex(kv[k]) for k in ("c0", "c1", "c2", "c3")],
                ctrl_base=int(kv["ctrl_base"]), a_base=int(kv["a_base"]), src=last_src,
                alt=[float.fromhex(kv[k]) for k in ("a0", "a1", "a2", "a3")], idx=int(kv["idx"])))
        elif tok[0] == "op":
            o = PlanOp(**{k: (float.fromhex(v) if k == "scale" else int(v)) for k, v in kv.items()})
            last_src = o.src
            if o.form == 6:          # header of a merged diagonal run: not an op of its own
                continue
            if passes[-1].direct:
                passes[-1].op = o
            else:
                passes[-1].stages[-1].ops.append(o)
    return passes


def summary(passes: List[PlanPass]) -> dict:
    tile = [p for p in passes if not p.direct]
    n_ops = sum(len(p.all_ops()) for p in passes)
    return {"passes": len(passes), "tile_passes": len(tile), "direct_passes": len(passes) - len(tile),
            "ops": n_ops, "stages": sum(len(p.stages) for p in tile),
            "ops_per_pass": n_ops / max(1, len(passes)),
            "peer_passes": sum(1 for p in tile if p.peer)}
