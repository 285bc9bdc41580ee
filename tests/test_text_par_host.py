"""CPU check of the parallel TEXT forward (kanzi-go_b200/csrc/kz_text_par_core.cuh, what the kernels of kz_text_par.cu run): candidate words in
parallel, the dictionary either as one pass over the candidates or as data-parallel rounds to a fixed point, emission through prefix
sums — all reproduce the one-pass walk (kz_text_core.cuh forward_walk, itself checked against the oracle) byte for byte."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from test_text_core_host import _words

synth = importlib.import_module("kanzi-go_b200.synth")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "text_par_host.cpp")
SO = os.path.join(ROOT, "tests", "host", "_build", "libtext_par_host.so")


@pytest.fixture(scope="module")
def tp():
    gen = importlib.import_module("kanzi-go_b200.gen_text_dict")
    if not gen.main():
        pytest.skip("static dictionary not available")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_text_par_core.cuh"), os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_text_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.tp_walk.restype = C.c_int64
    lib.tp_walk.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    lib.tp_phases.restype = C.c_int64
    lib.tp_phases.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.tp_inverse_walk.restype = C.c_int64
    lib.tp_inverse_walk.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_uint64]
    lib.tp_inverse_parallel.restype = C.c_int64
    lib.tp_inverse_parallel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_int)]
    lib.tp_parallel.restype = C.c_int64
    lib.tp_parallel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def _same_parallel(tp, x, bs, mode=0):
    x = np.ascontiguousarray(x, np.uint8)
    a = np.zeros(len(x) + 64, np.uint8)
    b = np.zeros(len(x) + 64, np.uint8)
    rounds, nc, nf = C.c_int(0), C.c_int(0), C.c_int(0)
    na = tp.tp_walk(x.ctypes.data, len(x), a.ctypes.data, bs, mode)
    nb = tp.tp_parallel(x.ctypes.data, len(x), b.ctypes.data, bs, mode, C.byref(rounds), C.byref(nc), C.byref(nf))
    assert nb != -2, "the parallel statement declined"
    assert na == nb, (na, nb, rounds.value)
    if na > 0:
        assert np.array_equal(a[:na], b[:nb]), int(np.argmax(a[:na] != b[:nb]))
    return na, rounds.value, nc.value, nf.value


@pytest.mark.parametrize("n,bs", [(1024, 1024), (5000, 4096), (100000, 1 << 16), (1 << 20, 1 << 20), (4 << 20, 4 << 20)])
def test_dictionary_rounds_match_the_walk(tp, n, bs):
    """the data-parallel dictionary (fixed point of the insert decisions) + prefix-sum emission == the serial walk"""
    texts = [synth.markov_text(n, seed=n), _words(n, 50, n + 1), _words(n, 5000, n + 2, crlf=True), synth.english_text(n, seed=n + 3),
             synth.english_text(n, seed=n + 4, crlf=False, html=True), synth.enwik_shaped(n, seed=n + 5), synth.xml_like(n, seed=n + 6)]
    worst = 0
    for x in texts:
        for mode in (0, 0x40):
            na, rounds, nc, nf = _same_parallel(tp, x, bs, mode)
            worst = max(worst, rounds)
    print("n", n, "rounds (max)", worst)
    assert worst <= 24
    spaces = np.concatenate([np.full(50, 0x20, np.uint8), synth.markov_text(n, seed=3)[: n - 50]])
    _same_parallel(tp, spaces, bs)
    dense = np.frombuffer((b"the and that have with " * (n // 23 + 1))[:n], np.uint8)
    _same_parallel(tp, dense, bs)
    caps = np.frombuffer((b"Hello hello HELLO hello Hello world World wOrld world abc Abc abc aBc " * (n // 70 + 1))[:n], np.uint8)
    _same_parallel(tp, caps, bs)
    hard = np.random.default_rng(n).integers(0, 256, n).astype(np.uint8)
    _same_parallel(tp, hard, bs)


def _same(tp, x, bs, mode=0):
    x = np.ascontiguousarray(x, np.uint8)
    a = np.zeros(len(x) + 64, np.uint8)
    b = np.zeros(len(x) + 64, np.uint8)
    nc, nf = C.c_int(0), C.c_int(0)
    na = tp.tp_walk(x.ctypes.data, len(x), a.ctypes.data, bs, mode)
    nb = tp.tp_phases(x.ctypes.data, len(x), b.ctypes.data, bs, mode, C.byref(nc), C.byref(nf))
    assert na == nb, (na, nb)
    if na > 0:
        assert np.array_equal(a[:na], b[:nb])
    return na, nc.value, nf.value


@pytest.mark.parametrize("n,bs", [(1024, 1024), (5000, 4096), (100000, 1 << 16), (1 << 20, 1 << 20), (3 << 20, 4 << 20)])
def test_three_phases_match_the_walk(tp, n, bs):
    for x in (synth.markov_text(n, seed=n), _words(n, 50, n + 1), _words(n, 5000, n + 2, crlf=True)):
        for mode in (0, 0x40, 0x20):
            na, nc, nf = _same(tp, x, bs, mode)
            assert na > 0 and 0 < nf <= nc < n / 2  # the serial pass touches far fewer items than the block has bytes
    spaces = np.concatenate([np.full(50, 0x20, np.uint8), synth.markov_text(n, seed=3)[: n - 50]])
    _same(tp, spaces, bs)
    dense = np.frombuffer((b"the and that have with " * (n // 23 + 1))[:n], np.uint8)  # every word is in the dictionary: single spaces implied
    _same(tp, dense, bs)
    hard = np.random.default_rng(n).integers(0, 256, n).astype(np.uint8)  # the walk itself does not care whether the block is text
    _same(tp, hard, bs)
    incompressible = _words(n, 200000, n + 4)  # few repeats: the output can exceed the bound -> both report the skip
    _same(tp, incompressible, bs)


def test_three_phases_dictionary_wrap(tp):
    na, nc, nf = _same(tp, _words(12 << 20, 700000, 7), 16 << 20)
    assert na > 0


@pytest.mark.parametrize("n,bs", [(1024, 1024), (5000, 4096), (100000, 1 << 16), (1 << 20, 1 << 20), (4 << 20, 4 << 20)])
def test_parallel_inverse_matches_the_walk(tp, n, bs):
    """encode with the walk, decode with the serial inverse walk and with the data-parallel statement: same bytes, and the input again"""
    texts = [synth.markov_text(n, seed=n), _words(n, 50, n + 1), _words(n, 5000, n + 2, crlf=True), synth.english_text(n, seed=n + 3),
             synth.english_text(n, seed=n + 4, crlf=False, html=True), synth.enwik_shaped(n, seed=n + 5), synth.xml_like(n, seed=n + 6),
             np.frombuffer((b"Hello hello HELLO hello Hello world World wOrld world abc Abc abc aBc a I a the " * (n // 80 + 1))[:n], np.uint8),
             np.random.default_rng(n).integers(0, 256, n).astype(np.uint8)]
    worst = 0
    for x in texts:
        x = np.ascontiguousarray(x, np.uint8)
        for mode in (0, 0x40):
            enc = np.zeros(len(x) + 64, np.uint8)
            ne = tp.tp_walk(x.ctypes.data, len(x), enc.ctypes.data, bs, mode)
            if ne <= 0:
                continue
            cap = 2 * len(x) + 1024  # mode 0x40 forced on LF-only text doubles every line end
            a = np.zeros(cap + 64, np.uint8)
            b = np.zeros(cap + 64, np.uint8)
            rounds = C.c_int(0)
            na = tp.tp_inverse_walk(enc.ctypes.data, ne, a.ctypes.data, cap, bs)
            nb = tp.tp_inverse_parallel(enc.ctypes.data, ne, b.ctypes.data, cap, bs, C.byref(rounds))
            assert nb != -2, "the parallel statement declined"
            assert na == nb, (na, nb)
            assert np.array_equal(a[:na], b[:nb]), int(np.argmax(a[:na] != b[:nb]))
            worst = max(worst, rounds.value)
    print("n", n, "inverse rounds (max)", worst)
