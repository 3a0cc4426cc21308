#!/usr/bin/env python
"""bench.py -- headline benchmark of the gate-application hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload random28|qft|qft_h] [--qubits n] [--depth d]

One "step" = one pass of the hot path over the workload circuit: `QReg::apply(&MultiOp)` of the
whole circuit on the device-resident register (BASELINE.json configs[1] by default: the 28-qubit
random layered circuit, depth 100, 4150 SingleOps, 4 GiB state).

  value  gates/s (reference SingleOps applied per second), state resident in HBM, CUDA-event time
         on the register's stream, max over ranks.
  e2e    the same metric through the public API a user of the reference calls per run:
         QReg reset -> apply(host op list) -> measure_mask -> outcome on the host, wall clock,
         host->device copy of the op descriptors and device->host read of the result inside.
  roofline  dominant kernel class: algorithmic bytes (16 B read + 16 B written per amplitude a
         launch can change) / CUDA-event duration of those launches, vs MEASURED_PEAKS.json.
  cpu_baseline  the oracle (C/OpenMP restatement of the reference; Rust cannot be built here)
         timed on this box's host cores on a bounded sample of the same circuit.

`--impl reference` times that CPU restatement as the reference arm (same metric/config/unit).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLASS_NAMES = ["direct_sweep", "tile_pass", "reduce_measure", "init_collapse_scale", "xgpu_barrier"]


def build_workload(args):
    from qvnt_b200 import workloads
    if args.workload == "random28":
        n = args.qubits or 28
        depth = args.depth or 100
        circ = workloads.random_layered(n, depth)
        name = f"configs[1]: {n}-qubit random layered circuit (h/rx/ry/rz + controlled x), depth {depth}"
    elif args.workload == "qft":
        n = args.qubits or 30
        circ = workloads.qft_full(n)
        name = f"{n}-qubit full QFT (op::qft)"
    elif args.workload == "qft_h":
        n = args.qubits or 32
        circ = workloads.qft_plus_h(n)
        name = f"configs[2]: {n}-qubit full QFT + Hadamard transform"
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    return n, circ, name


def measured_traffic(kernel, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture (profiles/traffic.json: bytes moved per amplitude)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        per_amp = json.load(open(p))[kernel]["dram_bytes_per_amplitude"]
        return per_amp * float(1 << n_local)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(pw))
        return out


def cpu_sample(n, circ, want_seconds=15.0, threads=None):
    """Time the CPU oracle on a bounded sample (a prefix of the same circuit, same state size)."""
    from oracle import oracle as orc
    from qvnt_b200.op import MultiOp
    threads = threads or orc.max_threads()
    ops = list(circ)
    reg = orc.OracleReg.new(n, threads=threads)
    # calibrate on 2 ops, then size the sample for ~want_seconds
    t0 = time.perf_counter()
    reg.apply(MultiOp(ops[:2]))
    per_op = max((time.perf_counter() - t0) / 2, 1e-6)
    k = int(max(2, min(len(ops), want_seconds / per_op)))
    reg.reset(0)
    arr, cnt = MultiOp(ops[:k]).to_c_array()
    t0 = time.perf_counter()
    reg.apply_raw(arr, cnt)
    dt = time.perf_counter() - t0
    reg.close()
    return {"value": k / dt, "unit": "gates/s", "cores": threads, "kind": "port",
            "sample": f"first {k} of {len(ops)} SingleOps of the same circuit on a {n}-qubit register "
                      f"({dt:.2f} s, out-of-place sweep per SingleOp like the reference, "
                      f"OpenMP {threads} threads)",
            "seconds": dt, "ops": k}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (C/OpenMP restatement; the Rust crate
    cannot be built in this image) on the box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from qvnt_b200.op import MultiOp
    n, circ, name = build_workload(args)
    threads = orc.max_threads()
    ops = list(circ)
    reg = orc.OracleReg.new(n, threads=threads)
    t0 = time.perf_counter()
    reg.apply(MultiOp(ops[:2]))
    per_op = max((time.perf_counter() - t0) / 2, 1e-6)
    total = args.steps + args.warmup
    k = int(max(1, min(len(ops), (150.0 / total) / per_op)))      # whole run within a few minutes
    arr, cnt = MultiOp(ops[:k]).to_c_array()
    times = []
    for s in range(total):
        reg.reset(0)
        t0 = time.perf_counter()
        reg.apply_raw(arr, cnt)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    reg.close()
    tt = sum(times)
    v = k * len(times) / tt
    sample = (f"each step = first {k} of {len(ops)} SingleOps of the circuit on a {n}-qubit register, "
              f"C/OpenMP restatement of the reference's out-of-place sweeps, {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": "gates/s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tt / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": name, "qubits": n, "single_ops": len(ops),
                                        "state_bytes": 16 << n, "ops_per_step": k},
        "cpu_baseline": {"value": v, "unit": "gates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "amplitude_gbs": v * 32 * (1 << n) / 1e9,
    }))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    import ctypes
    from qvnt_b200 import QReg, _ffi

    n, circ, name = build_workload(args)
    arr, n_ops = circ.to_c_array()
    op_bytes = n_ops * ctypes.sizeof(_ffi.QvntOp)

    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        reg = QReg.sharded(n, 0, rank, world, device=local_rank)
        blob = reg.export_ipc()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        reg.attach_peers(blobs)
    else:
        reg = QReg.new(n)
    if args.tile_bits:
        reg.set_option("tile_bits", args.tile_bits)
    if args.chunk_bits:
        reg.set_option("chunk_bits", args.chunk_bits)
    if args.no_fuse:
        reg.set_option("fuse", 0)
    if args.nbuf:
        reg.set_option("tile_nbuf", args.nbuf)
    if args.stagger >= 0:
        reg.set_option("tile_stagger", args.stagger)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    measure_mask = 0b100

    # ---- warm-up ----------------------------------------------------------------------------
    for _ in range(args.warmup):
        reg.reset(0)
        reg.apply_raw(arr, n_ops)
    reg.sync()

    # ---- timed: device-resident state, CUDA events on the register's stream -----------------
    reg.reset(0)
    reg.sync()
    reg.stats_reset()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    reg.mark(0)
    for _ in range(args.steps):
        reg.apply_raw(arr, n_ops)
    reg.mark(1)
    reg.sync()
    barrier()
    ms = reg.elapsed_ms(0, 1)
    clocks = sampler.stop() if sampler else None
    st = reg.stats()
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(sum(st["launches"]))
    value = n_ops * args.steps / (ms * 1e-3)
    timed_alg_gbs = [b / (ms * 1e-3) / 1e9 for b in st["alg_bytes"]]      # per class, over the timed region

    # ---- e2e: public API with host buffers, wall clock, copies inside --------------------------
    reg.stats_reset()
    barrier()
    t0 = time.perf_counter()
    outcome = None
    for _ in range(args.steps):
        reg.reset(0)
        reg.apply_raw(arr, n_ops)                       # op descriptors: host -> device every step
        outcome = reg.measure_mask(measure_mask, 0.5).get()   # result: device -> host every step
    reg.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    st2 = reg.stats()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": n_ops * args.steps / e2e_s, "unit": "gates/s",
           "h2d_bytes_per_step": int(st2["h2d_bytes"] // args.steps),
           "d2h_bytes_per_step": int(st2["d2h_bytes"] // args.steps),
           "ms_per_step": 1e3 * e2e_s / args.steps, "outcome": outcome,
           "what": "reset + apply(host op list) + measure_mask -> host, wall clock"}

    # ---- roofline: one extra instrumented step (CUDA events around every launch) --------------
    reg.reset(0)
    reg.sync()
    reg.stats_reset()
    reg.set_option("profile", 1)
    reg.apply_raw(arr, n_ops)
    reg.sync()
    sp = reg.stats()
    reg.set_option("profile", 0)
    peak, peak_src = measured_peaks()
    dom = max(range(len(CLASS_NAMES)), key=lambda c: sp["ms"][c])
    roofline = None
    if sp["launches"][dom]:
        achieved = sp["alg_bytes"][dom] / (sp["ms"][dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": CLASS_NAMES[dom], "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak,
                    # same kernel class over the TIMED region itself (its launches are the whole step:
                    # share_of_step): algorithmic bytes of those launches / CUDA-event time of the region
                    "achieved_timed_region": timed_alg_gbs[dom],
                    "traffic": measured_traffic(CLASS_NAMES[dom], n - (world.bit_length() - 1)),
                    "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel, per amplitude)",
                    "peak_source": peak_src,
                    "launches_per_step": int(sp["launches"][dom]),
                    "avg_launch_ms": sp["ms"][dom] / sp["launches"][dom],
                    "alg_bytes_per_launch": sp["alg_bytes"][dom] / sp["launches"][dom],
                    "share_of_step": sp["ms"][dom] / max(sum(sp["ms"]), 1e-12),
                    "passes_per_step": int(sp["passes"]), "gates_per_pass": n_ops / max(1, sp["passes"])}

    out = {
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "qubits": n, "single_ops": n_ops, "state_bytes": 16 << n,
                   "l2": "state (>= 4 GiB) is far larger than the 126 MB L2; no flush needed",
                   "sharding": f"top {world.bit_length() - 1} qubits across {world} GPU(s)",
                   "fuse": not args.no_fuse, "tile_bits": args.tile_bits or 11, "chunk_bits": args.chunk_bits or 4},
        "amplitude_gbs": value * 32 * (1 << n) / 1e9,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            try:
                out["cpu_baseline"] = cpu_sample(n, circ, want_seconds=args.cpu_seconds)
            except Exception as ex:                       # pragma: no cover
                out["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(out))
    reg.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="random28")
    ap.add_argument("--qubits", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--nbuf", type=int, default=0, help="tile buffers per CTA for T <= 11 (0 = auto)")
    ap.add_argument("--chunk-bits", type=int, default=0)
    ap.add_argument("--stagger", type=int, default=-1, help="start offset (cycles) between the CTAs of an SM")
    ap.add_argument("--no-fuse", action="store_true", help="one in-place sweep per SingleOp")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
