"""Oracle parity on the path SCALE times: one PROCESS per shard under torch.distributed.run, shards
attached through CUDA IPC handles (qvnt_reg_export_ipc / qvnt_reg_attach_peers), global-qubit gates
through peer-mapped memory.  With one visible GPU the ranks share it (the IPC path is the same);
with >= 2 GPUs rank k runs on device k and the peer traffic crosses NVLink.

Bars: amplitudes <= 1e-10 vs the oracle on the gathered register, sampled index bit-exact."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_workers(world, qubits, circuits, backend="gloo", opts=(), timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mp_worker.py"), "--qubits", str(qubits), "--backend", backend,
           "--circuits", *circuits]
    if opts:
        cmd += ["--opt", *opts]
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)       # torchrun would pin the oracle to one thread
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MP_PARITY ")]
    assert p.returncode == 0 and lines, f"rc={p.returncode}\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return json.loads(lines[-1][len("MP_PARITY "):])


def _worlds():
    from qvnt_b200 import _ffi
    ndev = _ffi.device_count()
    return [2, 4] if ndev < 8 else [2, 4, 8]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_parity(world):
    if world not in _worlds():
        pytest.skip("8 ranks only on an 8-GPU box")
    res = run_workers(world, 22, ["layered+mixed", "qft+mixed", "fast_mix"])
    for r in res["results"]:
        assert r["max_abs_err"] <= 1e-10, r
        assert r["max_abs_err_after_collapse"] <= 1e-10, r
        assert r["sampled_match"], r
        assert r["peer_bytes"] > 0, "the circuit must exercise the peer path"
    assert res["ok"]


def test_multi_process_parity_cp_async_path():
    res = run_workers(2, 20, ["layered"], opts=["tma=0"])
    assert res["ok"], res
