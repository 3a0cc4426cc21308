"""One rank of the multi-PROCESS sharded-register parity check (tests/test_multi_process.py,
bench.py's parity leg use the same routine).  Launched by torch.distributed.run, one process per
shard; the shards reach each other through CUDA IPC handles exactly as under `bench.py --gpus N`.
With fewer devices than ranks several shards share a device (still separate processes, still
cudaIpcOpenMemHandle).  Rank 0 gathers the shards, runs the CPU oracle on the same circuit and
state, and prints one JSON line.

Bars (BASELINE.json north_star): amplitudes <= 1e-10 absolute, sampled index bit-exact for the
injected uniform variate, state after collapse <= 1e-10.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_circuit(name, n):
    from qvnt_b200 import op, workloads
    if name == "layered+mixed":
        return workloads.random_layered(n, 6) * workloads.mixed_all_kinds(n, 60, seed=5)
    if name == "qft+mixed":
        return op.qft((1 << n) - 1) * workloads.mixed_all_kinds(n, 40, seed=11)
    if name == "layered":
        return workloads.random_layered(n, 8)
    if name == "fast_mix":       # controlled forms on register slots that lazy x left inverted
        return workloads.fast_mix(n, 300, seed=9) * workloads.random_layered(n, 3)
    raise SystemExit(f"unknown circuit {name}")


def sharded_parity(dist, rank, world, device, n, circ_name, state=0, u=0.4321, mask=None, options=()):
    """Runs `circ_name` on an n-qubit register sharded over the process group, compares with the
    oracle on rank 0.  `dist` is an initialised torch.distributed module (any backend whose
    collectives take CPU tensors, or NCCL with one device per rank).  Returns a dict on rank 0."""
    import torch
    from qvnt_b200 import QReg
    circ = build_circuit(circ_name, n)
    n_local = n - (world.bit_length() - 1)
    mask = ((1 << (n - 1)) | 0b1011) if mask is None else mask
    reg = QReg.sharded(n, state, rank, world, device=device)
    for k, v in options:
        reg.set_option(k, v)
    blobs = [None] * world
    dist.all_gather_object(blobs, reg.export_ipc())
    reg.attach_peers(blobs)
    dist.barrier()
    reg.apply(circ)
    reg.sync()
    dist.barrier()
    mine = reg.amplitudes()
    sampled = reg.measure_mask_full(mask, u)
    mine2 = reg.amplitudes()
    st = reg.stats()
    dist.barrier()
    reg.close()

    use_cuda = dist.get_backend() == "nccl"

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.float64))
        if use_cuda:
            t = t.cuda()
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        if rank != 0:
            return None
        return np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])

    full, full2 = gather(mine), gather(mine2)
    all_sampled = [None] * world
    dist.all_gather_object(all_sampled, tuple(int(x) for x in sampled))
    if rank != 0:
        return None
    from oracle import oracle as orc
    o = orc.OracleReg.with_state(n, state, threads=os.cpu_count() or 1)
    o.apply(circ)
    want = o.amplitudes().copy()
    m = tuple(int(x) for x in o.measure_mask_full(mask, u))
    want2 = o.amplitudes().copy()
    o.close()
    err = float(np.abs(full - want).max())
    err2 = float(np.abs(full2 - want2).max())
    same = all(tuple(s) == m for s in all_sampled)
    return {"circuit": circ_name, "qubits": n, "world": world, "n_local": n_local, "single_ops": len(circ),
            "max_abs_err": err, "max_abs_err_after_collapse": err2, "sampled": list(all_sampled[0]),
            "oracle_sampled": list(m), "sampled_match": bool(same), "peer_bytes": int(st["peer_bytes"]),
            "passes": int(st["passes"]), "ok": bool(err <= 1e-10 and err2 <= 1e-10 and same)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=22)
    ap.add_argument("--circuits", nargs="*", default=["layered+mixed", "qft+mixed"])
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--opt", nargs="*", default=[])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ndev = torch.cuda.device_count()
    device = local % max(ndev, 1)
    torch.cuda.set_device(device)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo")
    opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in a.opt]
    results = []
    for c in a.circuits:
        r = sharded_parity(dist, rank, world, device, a.qubits, c, options=opts)
        if rank == 0:
            results.append(r)
    if rank == 0:
        print("MP_PARITY " + json.dumps({"world": world, "devices": ndev, "results": results,
                                         "ok": all(r["ok"] for r in results)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
