"""The EXE logic of the product (kanzi-go_b200/csrc/kz_exe_core.cuh, shared by host and device code) compiled with g++ and run on the
CPU against the oracle; the kernels of kz_exe.cu call the same functions."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as po
from test_oracle_pre import _synth_arm64, _synth_x86

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "exe_core_host.cpp")
SO = os.path.join(ROOT, "tests", "host", "_build", "libexe_core_host.so")


@pytest.fixture(scope="module")
def xc():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_exe_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.xc_forward.restype = C.c_int64
    lib.xc_forward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int)]
    lib.xc_inverse.restype = C.c_int64
    lib.xc_inverse.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return lib


def _check(xc, x, dt_in=0):
    x = np.ascontiguousarray(x, np.uint8)
    n = len(x)
    cap = int(po.lib().kzo_transform_max_encoded_len(po.T_EXE, n)) + 64
    dst = np.zeros(cap, np.uint8)
    dt = C.c_int(0)
    try:
        want, wdt = po.transform_forward(po.T_EXE, x, dt_in)
    except po.OracleError:  # the reference panics (malformed executable header): a block error, not a skip
        assert xc.xc_forward(x.ctypes.data, n, dst.ctypes.data, cap, dt_in, C.byref(dt)) == -2
        return "panic"
    r = xc.xc_forward(x.ctypes.data, n, dst.ctypes.data, cap, dt_in, C.byref(dt))
    assert dt.value == wdt, (dt.value, wdt)
    if want is None:
        assert r == -1
        return None
    assert r == len(want) and np.array_equal(dst[:r], want), (r, len(want))
    back = np.zeros(n + 64, np.uint8)
    m = xc.xc_inverse(want.ctypes.data, len(want), back.ctypes.data, len(back))
    assert m == n and np.array_equal(back[:m], x)
    return want


@pytest.mark.parametrize("n", [4000, 4096, 100003, 1 << 20])
def test_exe_core_matches_oracle(xc, n):
    _check(xc, _synth_x86(n, n))
    _check(xc, _synth_arm64(n, n + 1))
    _check(xc, np.random.default_rng(n).integers(0, 256, n).astype(np.uint8))
    for dt in (1, 3, 7):
        _check(xc, _synth_x86(n, n + 2), dt)


def test_exe_core_real_binaries_and_headers(xc):
    exe = np.fromfile(sys.executable, dtype=np.uint8)
    if len(exe) > (1 << 16) and bytes(exe[:4]) == b"\x7fELF":
        assert _check(xc, exe[: 4 << 20]) is not None  # ELF header: the section table gives the code range
        _check(xc, exe[1 << 20: 3 << 20])              # no header: heuristics only
        bad = exe[: 1 << 20].copy()
        bad[0x28:0x30] = 0xFF                          # section table offset out of range: the reference indexes out of bounds
        _check(xc, bad)
        neg = exe[: 1 << 20].copy()
        neg[0x28:0x30] = [0x9C, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF]  # section table at offset -100: the reference panics
        assert _check(xc, neg) == "panic"
        bad2 = exe[: 1 << 20].copy()
        bad2[0x3C:0x3E] = [0xFF, 0x7F]                 # absurd section count
        _check(xc, bad2)
    mz = _synth_x86(1 << 16, 5).copy()
    mz[:2] = [0x4D, 0x5A]
    mz[60:64] = [0x80, 0, 0, 0]
    mz[0x80:0x84] = [0x50, 0x45, 0, 0]                  # "PE\\0\\0"
    mz[0x84:0x86] = [0x64, 0x86]                        # AMD64
    mz[0x80 + 28: 0x80 + 32] = np.frombuffer(np.uint32(40000).tobytes(), np.uint8)  # size of code
    mz[0x80 + 44: 0x80 + 48] = np.frombuffer(np.uint32(4096).tobytes(), np.uint8)   # base of code
    assert _check(xc, mz) is not None
    macho = _synth_x86(1 << 16, 6).copy()
    macho[:4] = [0xCF, 0xFA, 0xED, 0xFE]
    _check(xc, macho)                                   # random load commands: skip or panic, same as the oracle
