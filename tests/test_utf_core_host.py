"""The UTF logic of the product (kanzi-go_b200/csrc/kz_utf_core.cuh, shared by host and device code) compiled with g++ and run on the
CPU against the oracle; the kernels of kz_utf.cu call the same functions."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

synth = importlib.import_module("kanzi-go_b200.synth")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "utf_core_host.cpp")
SO = os.path.join(ROOT, "tests", "host", "_build", "libutf_core_host.so")


@pytest.fixture(scope="module")
def uc():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_utf_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.uc_forward.restype = C.c_int64
    lib.uc_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.uc_inverse.restype = C.c_int64
    lib.uc_inverse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    return lib


def utf_text(n_bytes, seed, n_symbols=0):
    r = np.random.default_rng(seed)
    words = ["héllo", "wörld", "naïve", "日本語", "テキスト", "данные", "текст", "abc", "xyz", " ", "\n", "😀", "€", "ß"]
    if n_symbols:  # many distinct code points: two-byte aliases (more than 128 symbols)
        cps = [chr(c) for c in r.choice(np.arange(0x400, 0x3000), n_symbols, replace=False)]
        words = words + ["".join(r.choice(cps, 3)) for _ in range(400)]
    out = []
    size = 0
    while size < n_bytes + 8:
        w = words[int(r.integers(0, len(words)))]
        out.append(w)
        size += len(w.encode("utf-8"))
    return np.frombuffer("".join(out).encode("utf-8")[:n_bytes], np.uint8)  # the cut may split a sequence: the last 4 bytes are copied raw


def _check(uc, x, dt_in=0):
    x = np.ascontiguousarray(x, np.uint8)
    want, wdt = po.transform_forward(po.T_UTF, x, dt_in)
    dst = np.zeros(len(x) + 8192 + 64, np.uint8)
    dt = C.c_int(0)
    n = uc.uc_forward(x.ctypes.data, len(x), dst.ctypes.data, dt_in, C.byref(dt))
    assert dt.value == wdt, (dt.value, wdt)
    if want is None:
        assert n == -1
        return None
    assert n == len(want) and np.array_equal(dst[:n], want), (n, len(want))
    back = np.zeros(len(x) + 64, np.uint8)
    m = uc.uc_inverse(want.ctypes.data, len(want), back.ctypes.data, len(back))
    assert m == len(x) and np.array_equal(back[:m], x)
    return want


@pytest.mark.parametrize("n", [1024, 1027, 4096, 50001, 1 << 20])
def test_utf_core_matches_oracle(uc, n):
    assert _check(uc, utf_text(n, n)) is not None
    assert _check(uc, utf_text(n, n + 1, n_symbols=min(2000, n // 20))) is not None or n < 8192
    bom = np.concatenate([np.frombuffer(b"x\xef\xbb\xbf", np.uint8), utf_text(n - 4, n + 2)])
    _check(uc, bom)
    _check(uc, utf_text(n, n + 3)[1:])  # starts inside a sequence: up to 3 leading bytes are copied raw
    _check(uc, synth.markov_text(n, seed=n))  # ASCII only: too few multi-byte sequences
    _check(uc, np.random.default_rng(n).integers(0, 256, n).astype(np.uint8))
    bad = utf_text(n, n + 4).copy()
    bad[n // 2] = 0xC0
    _check(uc, bad)
    bad2 = utf_text(n, n + 5).copy()
    bad2[n - 6] = 0xF8  # forbidden byte in the last partial window: not seen by the periodic check of validateUTF, caught by the walk
    _check(uc, bad2)
    for dt in (1, 7, 8):
        _check(uc, utf_text(n, n + 6), dt)
    _check(uc, np.random.default_rng(n + 7).integers(0, 256, n).astype(np.uint8), 8)  # declared UTF-8 but is not: the walk must decline
