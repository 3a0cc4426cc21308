"""The drop-in boundary: libqvnt_b200.so loads and exports every symbol include/qvnt_b200.h
declares; the POD op descriptor has the documented layout; without a CUDA device the product
path fails loudly (no CPU fallback).  No compute calls, no GPU needed."""
import ctypes
import os
import re

import pytest

from qvnt_b200 import _ffi
from qvnt_b200.optypes import QvntOp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qvnt_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qvnt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qvnt_b200.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_ffi.PROTOTYPES) == _declared_symbols()


def test_rust_bindings_cover_header():
    """rust/qvnt-b200-sys/src/lib.rs (uncompiled here: no Rust toolchain) declares every entry
    point of the header, with the POD layouts the C side uses."""
    rs = open(os.path.join(ROOT, "rust", "qvnt-b200-sys", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (qvnt_[a-z0-9_]+)\(", rs)))
    assert declared == _declared_symbols()
    assert "pub matrix: [f64; 32]" in rs and "#[repr(C)]" in rs


def test_op_layout():
    assert ctypes.sizeof(QvntOp) == 304
    assert QvntOp.a_mask.offset == 8 and QvntOp.ctrl.offset == 24
    assert QvntOp.phase_re.offset == 32 and QvntOp.matrix.offset == 48


def test_version_and_error_string():
    lib = _ffi.lib()
    assert lib.qvnt_version() >= 100
    assert lib.qvnt_last_error() is not None


def test_product_does_not_link_oracle():
    # the product library must not depend on anything under oracle/
    import subprocess
    out = subprocess.run(["ldd", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qvnt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "qvnt_oracle" not in txt, f


@pytest.mark.skipif(os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"),
                    reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from qvnt_b200 import QReg, QvntError
    with pytest.raises(QvntError) as e:
        QReg.new(4)
    assert e.value.status == 4          # QVNT_ERR_CUDA
