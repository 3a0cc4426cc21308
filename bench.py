#!/usr/bin/env python
"""bench.py -- headline benchmark of the gate-application hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload random|qft20|qft|qft_h] [--qubits n] [--depth d]

One "step" = one pass of the hot path over the workload circuit: `QReg::apply(&MultiOp)` of the
whole circuit on the device-resident register.  Default workload: BASELINE.json configs[1]'s
generator (random layered circuit h/rx/ry/rz + controlled x, depth 100) at the **30 qubits** the
metric is quoted on for one GPU (4450 SingleOps, 16 GiB state).  `--gpus N` shards the SAME circuit
by its top log2(N) qubits (strong scaling).

  value  gates/s (reference SingleOps applied per second), state resident in HBM, CUDA-event time
         on the register's stream, max over ranks.
  e2e    the same metric through the public API a user of the reference calls per run:
         QReg reset -> apply(host op list) -> measure_mask -> outcome on the host, wall clock,
         host->device copy of the op descriptors and device->host read of the result inside.
  roofline  dominant kernel class: algorithmic bytes (16 B read + 16 B written per amplitude a
         launch can change) / CUDA-event duration of those launches, vs MEASURED_PEAKS.json.
  cpu_baseline  the oracle (C/OpenMP restatement of the reference; Rust cannot be built here)
         timed on this box's host cores on a bounded sample of the same circuit.
  parity_check  (outside the timed region) a scaled-down twin of the workload on the same kind of
         register (same sharding, same process group), gathered and compared with the oracle.

`--impl reference` times the CPU restatement as the reference arm on the SAME config: every step
times k consecutive SingleOp sweeps of the circuit (out of place, all host cores, buffers allocated
and touched beforehand); the per-`apply` cost the reference pays once per MultiOp (two state-sized
allocations + the single-threaded `to_vec` copy, multi/mod.rs:96-114) is measured once and
amortised over the whole op list, as it is when the reference applies the whole circuit.
"""
from __future__ import annotations

import argparse
import ctypes
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
    from qvnt_b200 import op, workloads
    w = args.workload
    if w in ("random", "random28", "random30"):
        n = args.qubits or 30
        depth = args.depth or 100
        circ = workloads.random_layered(n, depth)
        name = f"configs[1] generator: {n}-qubit random layered circuit (h/rx/ry/rz + controlled x), depth {depth}"
    elif w == "qft20":
        n = args.qubits or 20
        circ = op.qft((1 << n) - 1)
        name = f"configs[0]: {n}-qubit QFT (op::qft(all)) from QReg::with_state, then measure_mask"
    elif w == "qft":
        n = args.qubits or 30
        circ = workloads.qft_full(n)
        name = f"{n}-qubit full QFT (op::qft)"
    elif w == "qft_h":
        n = args.qubits or 32
        circ = workloads.qft_plus_h(n)
        name = f"configs[2]: {n}-qubit full QFT + Hadamard transform"
    else:
        raise SystemExit(f"unknown workload {w}")
    return n, circ, name


def config_of(args, n, n_ops, name, world):
    """The SAME dict for both arms (the driver compares them)."""
    state = 16 << n
    l2 = ("state is far larger than the 126 MB L2; no flush needed" if state >= (1 << 30) else
          "state fits the L2: every step re-initialises the register (a full write) before the timed apply")
    return {"workload": name, "qubits": n, "single_ops": n_ops, "state_bytes": state, "l2": l2,
            "sharding": f"top {world.bit_length() - 1} qubits across {world} GPU(s)"}


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


# ---- the reference's CPU path (oracle port), timed ------------------------------------------------
class CpuSweeper:
    """Out-of-place sweeps of consecutive SingleOps between two pre-allocated, pre-touched buffers
    (dispatch.rs:50-67 on all host cores), plus the per-apply overhead of MultiOp::apply
    (multi/mod.rs:96-114: allocate the output buffer, `to_vec` the input) measured on its own."""

    def __init__(self, n, circ, threads):
        import numpy as np
        from oracle import oracle as orc
        self.orc, self.np = orc, np
        self.lib = orc.lib()
        self.n = n
        self.threads = threads
        self.ops = list(circ)
        from qvnt_b200.op import MultiOp
        self.arr, self.cnt = MultiOp(self.ops).to_c_array()
        self.op_size = int(self.lib.qo_sizeof_op())
        self.len = max(1 << n, 8)
        self.a = np.zeros(self.len, dtype=np.complex128)
        self.b = np.zeros(self.len, dtype=np.complex128)
        self.a[0] = 1.0
        self.b[:] = 0.0                                  # first touch outside every timed region
        self.pos = 0

    def sweeps(self, k):
        """Times k consecutive sweeps (cycling through the circuit); returns seconds."""
        base = ctypes.addressof(self.arr)
        t0 = time.perf_counter()
        for _ in range(k):
            self.lib.qo_sweep(ctypes.c_void_p(base + self.op_size * self.pos), ctypes.c_void_p(self.a.ctypes.data),
                              ctypes.c_void_p(self.b.ctypes.data), self.len, self.threads)
            self.a, self.b = self.b, self.a
            self.pos = (self.pos + 1) % self.cnt
        return time.perf_counter() - t0

    def apply_overhead(self):
        """What MultiOp::apply pays once per call besides its sweeps: a fresh output buffer
        (allocated, first touched by the first sweep) and the single-threaded copy of the input."""
        np = self.np
        t0 = time.perf_counter()
        out = np.empty(self.len, dtype=np.complex128)    # Vec::with_capacity + set_len
        inp = self.a.copy()                              # psi_i.to_vec()
        out[::256] = 0.0                                 # page-fault every 4 KiB page, as the first sweep does
        dt = time.perf_counter() - t0
        del out, inp
        return dt


def cpu_reference_rate(n, circ, steps, warmup, budget_s, threads=None):
    """gates/s of the reference's CPU algorithm on the whole op list: n_ops sweeps + one apply overhead."""
    threads = threads or (os.cpu_count() or 1)
    sw = CpuSweeper(n, circ, threads)
    n_ops = sw.cnt
    t1 = sw.sweeps(1)                                    # calibration (also warms the OpenMP pool)
    total = steps + warmup
    k = int(max(1, min(n_ops, (budget_s / max(total, 1)) / max(t1, 1e-7))))
    k = min(k, 64)
    overhead = sw.apply_overhead()
    times = []
    for s in range(total):
        dt = sw.sweeps(k)
        if s >= warmup:
            times.append(dt)
    per_sweep = sum(times) / (k * len(times))
    rate = n_ops / (n_ops * per_sweep + overhead)
    return {"rate": rate, "per_sweep_s": per_sweep, "apply_overhead_s": overhead, "k": k, "threads": threads,
            "ms_per_step": 1e3 * sum(times) / len(times), "n_ops": n_ops,
            "sample": (f"{len(times)} steps x {k} consecutive SingleOp sweeps of the circuit on a {n}-qubit state "
                       f"({per_sweep * 1e3:.1f} ms per out-of-place sweep, OpenMP {threads} threads, buffers "
                       f"pre-allocated) + the once-per-apply allocation and to_vec copy ({overhead:.2f} s) "
                       f"amortised over all {n_ops} SingleOps; C restatement of the reference (kind: port)")}


def cpu_qft20_rate(n, circ, steps, warmup, threads=None):
    """configs[0] end to end on the CPU port: with_state + apply(qft) + measure_mask per step."""
    from oracle import oracle as orc
    threads = threads or (os.cpu_count() or 1)
    arr, cnt = circ.to_c_array()
    times = []
    for s in range(steps + warmup):
        t0 = time.perf_counter()
        reg = orc.OracleReg.with_state(n, 0x5A5A5 & ((1 << n) - 1), threads=threads)
        reg.apply_raw(arr, cnt)
        reg.measure_mask_full(0b100, 0.5)
        dt = time.perf_counter() - t0
        reg.close()
        if s >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    return {"rate": cnt / per, "ms_per_step": 1e3 * per, "threads": threads, "n_ops": cnt,
            "sample": (f"{len(times)} steps of with_state + apply({cnt} SingleOps, one out-of-place sweep each) + "
                       f"measure_mask on a {n}-qubit register, OpenMP {threads} threads; C restatement (kind: port)")}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (C/OpenMP restatement; the Rust crate
    cannot be built in this image) on all host cores of the box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, circ, name = build_workload(args)
    threads = os.cpu_count() or 1                        # not OMP_NUM_THREADS: torchrun sets that to 1
    if args.workload == "qft20":
        r = cpu_qft20_rate(n, circ, args.steps, args.warmup, threads)
    else:
        r = cpu_reference_rate(n, circ, args.steps, args.warmup, budget_s=args.ref_seconds, threads=threads)
    v = r["rate"]
    print(json.dumps({
        "impl": "reference", "metric": "gates/s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(args, n, r["n_ops"], name, args.gpus),
        "cpu_baseline": {"value": v, "unit": "gates/s", "cores": threads, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "equivalent_unfused_gbs": v * 32 * (1 << n) / 1e9,
    }))


# ---- parity leg (outside the timed region) ---------------------------------------------------------
def parity_check(dist, rank, world, local_rank, qubits):
    if dist is None:
        import numpy as np
        from oracle import oracle as orc
        from qvnt_b200 import QReg
        from tests.mp_worker import build_circuit
        circ = build_circuit("layered+mixed", qubits)
        g = QReg.new(qubits)
        g.apply(circ)
        got = g.amplitudes()
        mask = (1 << (qubits - 1)) | 0b1011
        sg = tuple(int(x) for x in g.measure_mask_full(mask, 0.4321))
        got2 = g.amplitudes()
        g.close()
        o = orc.OracleReg.new(qubits, threads=os.cpu_count() or 1)
        o.apply(circ)
        err = float(np.abs(got - o.amplitudes()).max())
        so = tuple(int(x) for x in o.measure_mask_full(mask, 0.4321))
        err2 = float(np.abs(got2 - o.amplitudes()).max())
        o.close()
        return {"circuit": "layered+mixed", "qubits": qubits, "world": 1, "single_ops": len(circ),
                "max_abs_err": err, "max_abs_err_after_collapse": err2, "sampled_match": sg == so,
                "ok": bool(err <= 1e-10 and err2 <= 1e-10 and sg == so)}
    from tests.mp_worker import sharded_parity
    return sharded_parity(dist, rank, world, local_rank, qubits, "layered+mixed")


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    from qvnt_b200 import QReg

    n, circ, name = build_workload(args)
    arr, n_ops = circ.to_c_array()
    init_state = (0x5A5A5 & ((1 << n) - 1)) if args.workload == "qft20" else 0

    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        reg = QReg.sharded(n, init_state, rank, world, device=local_rank)
        blob = reg.export_ipc()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        reg.attach_peers(blobs)
    else:
        reg = QReg.with_state(n, init_state)
    if args.tile_bits:
        reg.set_option("tile_bits", args.tile_bits)
    if args.chunk_bits:
        reg.set_option("chunk_bits", args.chunk_bits)
    if args.no_fuse:
        reg.set_option("fuse", 0)
    if args.ctas:
        reg.set_option("tile_ctas", args.ctas)
    if args.tma >= -1:
        reg.set_option("tma", args.tma)
    for kv in args.opt:
        k, v = kv.split("=")
        reg.set_option(k, int(v))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    measure_mask = 0b100

    # ---- warm-up ----------------------------------------------------------------------------
    for _ in range(args.warmup):
        reg.reset(init_state)
        reg.apply_raw(arr, n_ops)
    reg.sync()

    # ---- timed: device-resident state, CUDA events on the register's stream -----------------
    small = (16 << n) < (1 << 30)          # L2-resident state: re-initialise (full write) between steps
    reg.reset(init_state)
    reg.sync()
    reg.stats_reset()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if small:
        ms = 0.0
        for _ in range(args.steps):
            reg.reset(init_state)
            reg.mark(0)
            reg.apply_raw(arr, n_ops)
            reg.mark(1)
            ms += reg.elapsed_ms(0, 1)
    else:
        reg.mark(0)
        for _ in range(args.steps):
            reg.apply_raw(arr, n_ops)
        reg.mark(1)
        reg.sync()
        ms = reg.elapsed_ms(0, 1)
    barrier()
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
    peer_gbs = st["peer_bytes"] / (ms * 1e-3) / 1e9

    # ---- e2e: public API with host buffers, wall clock, copies inside --------------------------
    reg.stats_reset()
    barrier()
    t0 = time.perf_counter()
    outcome = None
    for _ in range(args.steps):
        reg.reset(init_state)
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
    reg.reset(init_state)
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
    reg.close()

    # ---- parity leg: scaled-down twin through the same kind of register, vs the oracle ---------
    parity = None
    if not args.no_check:
        try:
            parity = parity_check(dist, rank, world, local_rank, args.check_qubits)
        except Exception as ex:                           # pragma: no cover
            parity = {"ok": False, "error": repr(ex)}

    cfg = config_of(args, n, n_ops, name, world)
    out = {
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "engine": {"fuse": not args.no_fuse, "tile_bits": args.tile_bits or 11, "chunk_bits": args.chunk_bits or 4,
                   "tile_loads": ("cp.async.bulk + mbarrier" if args.tma == 1 else "cp.async 16 B" if args.tma == 0 else
                                  "cp.async 16 B (local passes), cp.async.bulk + mbarrier (passes that read a peer shard)")},
        "equivalent_unfused_gbs": value * 32 * (1 << n) / 1e9,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
        "parity_check": parity,
    }
    if world > 1:
        out["nvlink_gbs_per_gpu"] = peer_gbs          # bytes this GPU read + wrote in peer HBM / step time
    if rank == 0:
        if world == 1 and not args.no_cpu:
            try:
                if args.workload == "qft20":
                    r = cpu_qft20_rate(n, circ, 3, 1)
                else:
                    r = cpu_reference_rate(n, circ, steps=3, warmup=1, budget_s=args.cpu_seconds)
                out["cpu_baseline"] = {"value": r["rate"], "unit": "gates/s", "cores": r["threads"], "kind": "port",
                                       "sample": r["sample"]}
            except Exception as ex:                       # pragma: no cover
                out["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="random")
    ap.add_argument("--qubits", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the parity leg")
    ap.add_argument("--check-qubits", type=int, default=24)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="reference arm: budget for all timed sweeps")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0, help="CTAs per SM of the tile pass for T <= 11 (0 = auto, 3..5)")
    ap.add_argument("--tma", type=int, default=-2, help="tile loads: 1 cp.async.bulk, 0 16-byte cp.async, -1/default auto (bulk for peer passes)")
    ap.add_argument("--chunk-bits", type=int, default=0)
    ap.add_argument("--opt", nargs="*", default=[], help="library options key=value")
    ap.add_argument("--no-fuse", action="store_true", help="one in-place sweep per SingleOp")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
