"""The C++ host mirror of `qvnt::prelude` (include/qvnt.hpp) over the C ABI: compiled here with
g++, its host-only checks (operator algebra, lowering, names) run on the CPU; the device checks
(the reference's `quantum_reg` golden state, Bell pair, matrix_repr, tensor product) run on a GPU."""
import os
import subprocess

import pytest

from qvnt_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_test.bin")


def build():
    libdir = os.path.dirname(_ffi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_test.cpp"), "-o", EXE,
           "-L", libdir, "-lqvnt_b200", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return EXE


def test_cpp_host_builds_and_host_checks_pass():
    exe = build()
    out = subprocess.run([exe, "--host-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp host mirror ok" in out.stdout


@pytest.mark.gpu
def test_cpp_host_device_checks():
    exe = build()
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp host mirror ok" in out.stdout
