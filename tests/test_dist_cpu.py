"""Host logic of the N > 1 path on the CPU: two `gloo` ranks (one process each, as bench.py is
launched under torchrun) plan the SAME op list for their own shard, exchange what they planned
with `all_gather_object` (the transport bench.py uses for the IPC blobs) and check the SPMD
contract of qvnt_reg_apply on a sharded register:
  - every rank schedules the same passes in the same order (same barriers -> no deadlock),
  - in every tile pass the ranks' tiles partition the full index space: each amplitude of the
    whole register is owned by exactly one tile of one rank,
  - passes that touch a peer shard are flagged on every rank.
No CUDA involved (qvnt_plan_describe is host-only)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist           # noqa: E402
import torch.multiprocessing as mp         # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tile_indices(p, tile_i):
    base = tile_i
    for pos in p.fx_pos:
        base = ((base >> pos) << (pos + 1)) | (base & ((1 << pos) - 1))
    base |= p.fx_val | p.base_or
    idx = np.full(1, base, dtype=np.uint64)
    for g in p.gpos:
        idx = np.concatenate([idx, idx | np.uint64(1 << g)])
    return idx


def _worker(rank, world, port, n, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qvnt_b200 import op, plan, workloads
        circ = workloads.random_layered(n, 5) * op.qft((1 << n) - 1) * op.swap((1 << (n - 1)) | 1)
        passes = plan.describe(n, circ, rank=rank, world=world, peers=True)
        mine = []
        for p in passes:
            if p.direct:
                mine.append(("direct", p.op.src, None))
                continue
            owned = np.concatenate([_tile_indices(p, t) for t in range(p.n_tiles)])
            mine.append(("tile", p.peer, (tuple(p.gpos), [o.src for o in p.all_ops()], np.sort(owned))))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        ok = True
        msg = ""
        if rank == 0:
            n_pass = {len(g) for g in gathered}
            ok = len(n_pass) == 1
            peer_passes = 0
            for k in range(len(mine)) if ok else []:
                kinds = {g[k][0] for g in gathered}
                if len(kinds) != 1:
                    ok, msg = False, f"pass {k}: ranks disagree on the kind"
                    break
                if mine[k][0] == "direct":
                    ok = ok and len({g[k][1] for g in gathered}) == 1
                    continue
                ok = ok and len({g[k][1] for g in gathered}) == 1                  # peer flag
                ok = ok and len({(g[k][2][0], tuple(g[k][2][1])) for g in gathered}) == 1   # geometry, ops
                peer_passes += mine[k][1]
                count = np.zeros(1 << n, dtype=np.int32)
                for g in gathered:
                    np.add.at(count, g[k][2][2].astype(np.int64), 1)
                if not np.all(count == 1):
                    ok, msg = False, f"pass {k}: tiles do not partition the register"
                    break
            ok = ok and peer_passes >= 1
        flag = torch.tensor([1 if ok else 0])
        dist.broadcast(flag, src=0)
        result[rank] = (int(flag.item()), msg)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_spmd_plans_partition_the_register(world):
    n = 12
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, result), nprocs=world, join=True)
    assert all(result[r][0] == 1 for r in range(world)), dict(result)
