"""The TEXT state machine of the product (kanzi-go_b200/csrc/kz_text_core.cuh, shared by host and device code) compiled with g++ and run
on the CPU against the oracle. The kernels of kz_text.cu call the same functions from one thread per block; this test confirms the
logic without a GPU (the GPU parity tests confirm the kernels around it)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

synth = importlib.import_module("kanzi-go_b200.synth")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "text_core_host.cpp")
INC = os.path.join(ROOT, "kanzi-go_b200", "csrc", "_gen", "kz_text_dict.inc")
SO = os.path.join(ROOT, "tests", "host", "_build", "libtext_core_host.so")


@pytest.fixture(scope="module")
def tc():
    gen = importlib.import_module("kanzi-go_b200.gen_text_dict")
    if not gen.main() or not po.lib().kzo_text_available():
        pytest.skip("static dictionary not available (no reference tree and no generated file)")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC, INC, os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_text_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.tc_forward.restype = C.c_int64
    lib.tc_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.tc_inverse.restype = C.c_int64
    lib.tc_inverse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_uint64]
    return lib


def _words(n_bytes, n_distinct, seed, crlf=False):
    """text with many distinct words (dictionary growth, wrap-around at 2^19 entries when n_distinct is large)"""
    r = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", np.uint8)
    lens = r.integers(3, 12, n_distinct)
    vocab = [bytes(r.choice(alphabet, int(l))) for l in lens]
    common = [b"the", b"and", b"that", b"have", b"with", b"people", b"because", b"The", b"And", b"HTTP", b"don't", b"x", b"ab"]
    out = bytearray()
    seps = [b" ", b" ", b" ", b", ", b". ", b"\r\n" if crlf else b"\n", b"; ", b" (", b") ", b"_", b"\t"]
    while len(out) < n_bytes:
        w = vocab[int(r.integers(0, n_distinct))] if r.random() < 0.7 else common[int(r.integers(0, len(common)))]
        if r.random() < 0.1:
            w = w.capitalize()
        out += w + seps[int(r.integers(0, len(seps)))]
        if r.random() < 0.01:
            out += bytes([0x0F, 0xC3, 0xA9]) + b" "  # escape token and a two-byte UTF-8 letter
    return np.frombuffer(bytes(out[:n_bytes]), np.uint8)


def _check(tc, x, bs, dt_in=0):
    x = np.ascontiguousarray(x, np.uint8)
    want, wdt = po.transform_forward(po.T_TEXT, x, dt_in, block_size=bs)
    dst = np.zeros(len(x) + 64, np.uint8)
    dt = C.c_int(0)
    n = tc.tc_forward(x.ctypes.data, len(x), dst.ctypes.data, bs, dt_in, 0, C.byref(dt))
    assert dt.value == wdt
    if want is None:
        assert n == -1
        return None
    assert n == len(want) and np.array_equal(dst[:n], want), (n, len(want))
    back = np.zeros(len(x) + 4096, np.uint8)
    m = tc.tc_inverse(want.ctypes.data, len(want), back.ctypes.data, len(back), bs)
    assert m == len(x) and np.array_equal(back[:m], x)
    return want


@pytest.mark.parametrize("n,bs", [(1024, 1024), (5000, 4096), (100000, 1 << 16), (1 << 20, 1 << 20), (3 << 20, 4 << 20)])
def test_text_core_matches_oracle(tc, n, bs):
    assert _check(tc, synth.markov_text(n, seed=n), bs) is not None
    assert _check(tc, _words(n, 50, n + 1), bs) is not None
    assert _check(tc, _words(n, 5000, n + 2, crlf=True), bs) is not None
    _check(tc, np.random.default_rng(n).integers(0, 256, n).astype(np.uint8), bs)  # not text: both decline with the same data type
    _check(tc, np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(n + 1).integers(0, 4, n)], bs)
    for dt in (1, 2, 6, 7):
        _check(tc, synth.markov_text(n, seed=n + 3), bs, dt)


def test_text_core_dictionary_wraps(tc):
    # more than 2^19 distinct words: the entry index wraps back to the end of the static dictionary (:1316-1321) and entries are reused
    x = _words(12 << 20, 700000, 7)
    assert _check(tc, x, 16 << 20) is not None


def test_text_core_xml_flag(tc):
    body = (b"<item id=\"1\">the &amp; and &lt;that&gt; have &quot;with&quot;</item>\n" * 3000)
    y = _check(tc, np.frombuffer(body, np.uint8), 1 << 20)
    assert y is not None and (y[0] & 0x20)
