#!/usr/bin/env python
"""configs[3] / configs[4] at full size: a register sharded over N GPUs (one process per GPU,
launched with torch.distributed.run), driven SPMD through the C ABI.

  --workload random   configs[3]: the random layered circuit over ALL n qubits (global qubits are
                      hit by h/rx/ry/rz and the controlled-x bricks), then measure_mask
  --workload qasm     configs[4]: generated OpenQASM 2.0 (h, ccx/cccx chains, rzz, i_swap layers,
                      `measure q -> c`) lowered by qvnt_b200.qasm and executed like Sym::finish

Prints one JSON line on rank 0: wall/device time, gates/s, HBM passes, bytes moved over NVLink and
the NVLink GB/s they imply, and the size-independent checks that stand in for the oracle at sizes no
CPU can hold: sum |a|^2 = 1 after the unitary part (global reduction), every rank agrees on the
sampled index, the collapsed state's norm equals the probability of the outcome's subspace
(<= 1), and the measured bits are consistent with the mask.  Parity itself is tested on the
scaled-down twins (tests/test_multi_gpu.py, tests/test_qasm.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="qasm", choices=["qasm", "random"])
    ap.add_argument("--qubits", type=int, default=36)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--u", type=float, default=0.6180339887)
    ap.add_argument("--opt", nargs="*", default=[], help="library options key=value (fuse, tile_bits, chunk_bits, tile_ctas, tma)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    from qvnt_b200 import QReg, qasm, workloads
    from qvnt_b200.op import MultiOp

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.qubits
    t0 = time.perf_counter()
    reg = QReg.sharded(n, 0, rank, world, device=local_rank)
    if world > 1:
        blobs = [None] * world
        dist.all_gather_object(blobs, reg.export_ipc())
        reg.attach_peers(blobs)
    else:
        reg.attach_peers([reg.export_ipc()])
    for kv in args.opt:
        k, v = kv.split("=")
        reg.set_option(k, int(v))
    reg.sync()
    t_alloc = time.perf_counter() - t0

    if args.workload == "qasm":
        prog = qasm.Int(workloads.qasm_config5(n, args.layers))
        segs = [(m, s) for m, s in prog.q_ops.segs] + [(prog.q_ops.tail, qasm.Sep("Nop"))]
        name = f"configs[4]: {n}-qubit QASM circuit (h, ccx/cccx, rzz, i_swap x {args.layers} layers, measure q -> c)"
    else:
        circ = workloads.random_layered(n, args.layers)
        segs = [(circ, qasm.Sep("Measure", (1 << n) - 1, (1 << n) - 1))]
        name = f"configs[3]: {n}-qubit random layered circuit over all qubits, depth {args.layers}, then measure"
    n_ops = sum(len(m) for m, _ in segs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reg.stats_reset()
    barrier()
    t0 = time.perf_counter()
    reg.mark(0)
    norm_before = None
    outcome = None
    norm_after = None
    for mop, sep in segs:
        if len(mop):
            arr, cnt = mop.to_c_array()
            reg.apply_raw(arr, cnt)
        if sep.kind == "Measure":
            reg.mark(1)
            norm_before = reg.get_absolute()              # global sum |a|^2 (collective)
            outcome = reg.measure_mask_full(sep.a, args.u)
            norm_after = reg.get_absolute()
    reg.mark(2)
    reg.sync()
    barrier()
    wall = time.perf_counter() - t0
    ms_apply = reg.elapsed_ms(0, 1) if norm_before is not None else reg.elapsed_ms(0, 2)
    st = reg.stats()
    vals = torch.tensor([ms_apply, float(st["peer_bytes"]), float(outcome[1] if outcome else 0)],
                        device="cuda", dtype=torch.float64)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = vals.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = mn = sm = vals
    if rank == 0:
        ms = float(mx[0])
        peer_total = float(sm[1])
        mask = (1 << n) - 1
        checks = {
            "norm_sqr_before_measure": norm_before,
            "norm_ok": norm_before is not None and abs(norm_before - 1.0) < 1e-9,
            "all_ranks_same_sample": float(mx[2]) == float(mn[2]),
            "outcome": outcome[0] if outcome else None,
            "outcome_in_mask": outcome is not None and (outcome[0] & ~mask) == 0,
            "norm_sqr_after_collapse": norm_after,
            "collapse_norm_le_1": norm_after is not None and 0.0 < norm_after <= 1.0 + 1e-9,
        }
        print(json.dumps({
            "workload": name, "n_gpus": world, "qubits": n, "local_qubits": n - (world.bit_length() - 1),
            "shard_gib": (16 << (n - (world.bit_length() - 1))) / 2 ** 30, "single_ops": n_ops,
            "apply_ms": ms, "gates_per_s": n_ops / (ms * 1e-3), "wall_s_incl_measure": wall,
            "alloc_init_s": t_alloc, "hbm_passes": int(st["passes"]),
            "amplitude_gbs_equivalent": n_ops * 32.0 * (1 << n) / (ms * 1e-3) / 1e9,
            "nvlink_bytes_total": peer_total,
            "nvlink_gbs_per_gpu_both_directions": peer_total / max(world, 1) / (ms * 1e-3) / 1e9,
            "launches": st["launches"], "checks": checks,
        }))
    reg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
