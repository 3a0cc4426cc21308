"""Sharded-register parity: the register split by its top log2(P) qubits into P
shards, each shard driven by its own host thread through the C ABI, exactly as one process per
GPU would drive it.  Global-qubit gates run inside the tile kernel through peer-mapped memory.

On a 1-GPU box all P shards live on device 0 (same-process attach path of
qvnt_reg_attach_peers); with >= 2 devices shard k lives on device k % n_dev (peer access over
NVLink).  The cross-process CUDA-IPC attach path is exercised by `bench.py --gpus N` under
torchrun.  Bars: amplitudes <= 1e-10 vs the oracle on the full register, sampled index bit-exact.
"""
import threading

import numpy as np
import pytest

from qvnt_b200 import QReg, op, workloads

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-10


def n_devices():
    from qvnt_b200 import _ffi
    return _ffi.device_count()


def run_sharded(n, world, state, circ, measure=None, fuse=1, init=None):
    """Returns (full amplitude vector after circ, measure result per rank, amplitudes after measure)."""
    ndev = n_devices()
    n_local = n - (world.bit_length() - 1)
    blobs = [None] * world
    amps = [None] * world
    amps2 = [None] * world
    meas = [None] * world
    errs = []
    bar = threading.Barrier(world)

    def worker(rank):
        try:
            reg = QReg.sharded(n, state, rank, world, device=rank % ndev)
            reg.set_option("fuse", fuse)
            blobs[rank] = reg.export_ipc()
            bar.wait()
            reg.attach_peers(blobs)
            if init is not None:
                reg.write_amplitudes(init[rank << n_local:(rank + 1) << n_local])
                reg.sync()
            bar.wait()
            reg.apply(circ)
            reg.sync()
            bar.wait()
            amps[rank] = reg.amplitudes()
            if measure is not None:
                meas[rank] = reg.measure_mask_full(*measure)
                amps2[rank] = reg.amplitudes()
            bar.wait()
            reg.close()
        except Exception as ex:       # pragma: no cover
            errs.append((rank, repr(ex)))
            bar.abort()

    ts = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "sharded run hung"
    assert not errs, errs
    full = np.concatenate(amps)
    full2 = np.concatenate(amps2) if measure is not None else None
    return full, meas, full2


def oracle_run(oracle, n, state, circ, measure=None, init=None):
    o = oracle.OracleReg.with_state(n, state, threads=oracle.max_threads())
    if init is not None:
        o.write_amplitudes(init)
    o.apply(circ)
    a = o.amplitudes().copy()
    m, a2 = None, None
    if measure is not None:
        m = o.measure_mask_full(*measure)
        a2 = o.amplitudes().copy()
    o.close()
    return a, m, a2


def global_heavy(n, world):
    """Circuit that keeps hitting the sharded (top) qubits with every gate class."""
    top = n - 1
    wb = world.bit_length() - 1
    lo_g = n - wb
    c = op.h((1 << n) - 1)
    c *= op.rx(0.37, 1 << top) * op.ry(1.1, 1 << lo_g) * op.rz(0.77, 1 << top)
    c *= op.x(1 << top).c(1 << 0) * op.x(1 << 1).c(1 << top)
    c *= op.swap((1 << top) | 1) * op.i_swap((1 << lo_g) | (1 << 2))
    c *= op.rzz(0.3, (1 << top) | (1 << 3)) * op.rxx(0.9, (1 << top) | (1 << 4))
    c *= op.ryy(0.5, (1 << lo_g) | (1 << 5)) * op.sqrt_swap((1 << top) | (1 << 2))
    c *= op.y((1 << top) | 3) * op.z((1 << top) | 4) * op.s(1 << top) * op.t(1 << lo_g)
    c *= op.h((1 << top) | (1 << 1)) * op.x((1 << 2) | (1 << 3)).c((1 << top) | (1 << 0))
    c *= op.sqrt_i_swap((1 << top) | (1 << 6)).dgr()
    c *= op.u3(0.3, 0.4, 0.5, 1 << top)
    return c


@pytest.mark.parametrize("world,fuse", [(2, 0), (2, 1), (4, 1), (8, 1)])
def test_sharded_global_gates(oracle, world, fuse):
    n = 13
    rng = np.random.default_rng(5 + world)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    circ = global_heavy(n, world)
    got, _, _ = run_sharded(n, world, 0, circ, fuse=fuse, init=v)
    want, _, _ = oracle_run(oracle, n, 0, circ, init=v)
    assert np.abs(got - want).max() <= AMP_TOL


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_config4_twin(oracle, world):
    """configs[3] scaled down: the random layered circuit over ALL
    qubits, so the top log2(P) qubits are hit by h/rx/ry/rz and by the controlled-x bricks."""
    n = 16
    circ = workloads.random_layered(n, 6)
    mask = (1 << (n - 1)) | 0b1011
    got, meas, got2 = run_sharded(n, world, 0, circ, measure=(mask, 0.4321))
    want, m, want2 = oracle_run(oracle, n, 0, circ, measure=(mask, 0.4321))
    assert np.abs(got - want).max() <= AMP_TOL
    assert all(tuple(x) == tuple(m) for x in meas), (meas, m)
    assert np.abs(got2 - want2).max() <= AMP_TOL


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_qft_and_mixed(oracle, world):
    n = 14
    circ = op.qft((1 << n) - 1) * workloads.mixed_all_kinds(n, 40, seed=11 + world)
    got, meas, got2 = run_sharded(n, world, 0x1234, circ, measure=((1 << n) - 1, 0.77))
    want, m, want2 = oracle_run(oracle, n, 0x1234, circ, measure=((1 << n) - 1, 0.77))
    assert np.abs(got - want).max() <= AMP_TOL
    assert all(tuple(x) == tuple(m) for x in meas), (meas, m)
    assert np.abs(got2 - want2).max() <= AMP_TOL


def test_p_invariance():
    """The same circuit on P = 1, 2, 4 shards gives the same state."""
    n = 18
    circ = workloads.random_layered(n, 4) * op.qft((1 << n) - 1)
    ref = QReg.new(n)
    ref.apply(circ)
    want = ref.amplitudes()
    ref.close()
    for world in (2, 4):
        got, _, _ = run_sharded(n, world, 0, circ)
        assert np.abs(got - want).max() <= AMP_TOL


def test_sharded_lazy_x_large_tiles_regression():
    """Regression: a lazy x makes threads store into each other's slots of the tile buffer; without
    a barrier between a stage's loads and stores a fast warp could overwrite amplitudes a slow warp
    had not loaded yet.  Only showed with >= 2^26-amplitude shards and two kernels sharing the GPU
    (found by the norm check of tools/run_sharded.py at 34 qubits on 8 GPUs)."""
    n = 27
    circ = workloads.random_layered(n, 20)
    ref = QReg.new(n)
    ref.apply(circ)
    want = ref.amplitudes()
    ref.close()
    got, _, _ = run_sharded(n, 2, 0, circ)
    assert abs(float(np.vdot(got, got).real) - 1.0) < 1e-12
    assert np.abs(got - want).max() <= AMP_TOL


# ---- one handle, one process, several GPUs (qvnt_reg_create_multi / QReg::num_threads) -------------
@pytest.fixture
def share_devices(monkeypatch):
    """On a box with fewer GPUs than shards the shards share devices (test facility of the library)."""
    monkeypatch.setenv("QVNT_MULTI_SHARE_DEVICES", "1")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_handle_multi_gpu(oracle, share_devices, world):
    n = 16
    circ = workloads.random_layered(n, 6) * global_heavy(n, world) * op.qft((1 << n) - 1)
    mask = (1 << (n - 1)) | 0b1011
    reg = QReg.multi(n, 0x4321, world)
    reg.apply(circ)
    got = reg.amplitudes()
    assert got.size == 1 << n
    m = reg.measure_mask_full(mask, 0.4321)
    got2 = reg.amplitudes()
    p = reg.get_probabilities()
    st = reg.stats()
    reg.close()
    want, mo, want2 = oracle_run(oracle, n, 0x4321, circ, measure=(mask, 0.4321))
    assert np.abs(got - want).max() <= AMP_TOL
    assert tuple(m) == tuple(mo)
    assert np.abs(got2 - want2).max() <= AMP_TOL
    assert abs(p.sum() - 1.0) < 1e-12
    assert st["peer_bytes"] > 0


def test_num_threads_moves_the_register_to_more_gpus(oracle, share_devices):
    """QReg::num_threads (quant.rs:186-200) with GPUs for threads: the state survives the move."""
    n = 15
    c1 = workloads.random_layered(n, 3)
    c2 = workloads.mixed_all_kinds(n, 40, seed=3) * op.h(1 << (n - 1))
    reg = QReg.with_state(n, 5)
    reg.apply(c1)
    reg = reg.num_threads(4)
    assert reg is not None and reg.world == 4
    reg.apply(c2)
    reg = reg.num_threads(1)
    assert reg is not None and reg.world == 1
    got = reg.amplitudes()
    reg.close()
    assert QReg.new(4).num_threads(0) is None
    want, _, _ = oracle_run(oracle, n, 5, c1 * c2)
    assert np.abs(got - want).max() <= AMP_TOL


def test_remap_keeps_api_in_qubit_order(oracle, share_devices):
    """After remap passes the qubits sit at other index bits; every index-addressed call
    (read, probabilities, collapse, reset_by_mask, measure) must still see qubit order."""
    n = 14
    circ = workloads.random_layered(n, 5)
    for world in (2, 4):
        reg = QReg.multi(n, 0, world)
        o = oracle.OracleReg.new(n, threads=oracle.max_threads())
        reg.apply(circ)
        o.apply(circ)
        reg.collapse_mask(0b101 << (n - 3), 0b111 << (n - 3))
        o.collapse_mask(0b101 << (n - 3), 0b111 << (n - 3))
        reg.apply(circ)                     # a second apply continues from the restored layout
        o.apply(circ)
        reg.reset_by_mask(1 << (n - 1))
        o.reset_by_mask(1 << (n - 1))
        assert np.abs(reg.amplitudes() - o.amplitudes()).max() <= AMP_TOL
        assert np.abs(reg.get_probabilities() - o.get_probabilities()).max() <= 1e-12
        reg.close()
        o.close()
