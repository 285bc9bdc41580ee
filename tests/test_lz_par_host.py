"""CPU check of the data-parallel LZ / LZX parse (kanzi-go_b200/csrc/kz_lz_par_core.cuh, the code the kernels of kz_lz_par.cu execute): precomputed
candidates + speculative segment parses + stitch pass, iterated to the fixed point of the table membership (skip acceleration), + emission
produce exactly the oracle's bytes, for several segment sizes, including blocks with long incompressible stretches and degenerate data."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

synth = importlib.import_module("kanzi-go_b200.synth")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "lz_par_host.cpp")
HDR = os.path.join(ROOT, "kanzi-go_b200", "csrc", "kz_lz_par_core.cuh")
SO = os.path.join(ROOT, "tests", "host", "_build", "liblz_par_host.so")


@pytest.fixture(scope="module")
def lz():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-x", "c++", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.lz_forward_par.restype = C.c_int64
    lib.lz_forward_par.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    return lib


def cases(n, seed):
    r = np.random.default_rng(seed)
    yield "text", synth.markov_text(n, seed=seed)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 1)
    yield "zipf", synth.zipf_bytes(n, 1.5, seed=seed + 2)
    yield "periodic", np.resize(np.frombuffer(b"abracadabra, ", np.uint8), n)
    yield "runs", np.repeat(r.integers(0, 256, n // 37 + 1).astype(np.uint8), 37)[:n]
    yield "const", np.full(n, 7, np.uint8)
    yield "dna", np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, n)]
    mixed = synth.markov_text(n, seed=seed + 3).copy()
    k = n // 3
    mixed[k: 2 * k] = r.integers(0, 256, k).astype(np.uint8)  # a long incompressible stretch: the skip acceleration (srcInc >> 6) kicks in
    yield "mixed", mixed
    sparse = r.integers(0, 256, n).astype(np.uint8)  # incompressible with islands of repeats: skip mode with matches inside
    for o in range(0, n - 600, 5000):
        sparse[o + 300: o + 600] = sparse[o: o + 300]
    yield "sparse", sparse
    yield "x86", synth.x86_like(n, seed=seed + 4)
    yield "records", synth.records_like(n, seed=seed + 5)


def run(lz, extra, dt, x, seg):
    n = len(x)
    dst = np.zeros(n + n // 64 + 64, np.uint8)
    stats = (C.c_int64 * 8)()
    r = lz.lz_forward_par(extra, dt, x.ctypes.data, n, dst.ctypes.data, len(dst), seg, stats)
    return r, dst, list(stats)


@pytest.mark.parametrize("extra,n", [(0, 24), (1, 24), (0, 100), (1, 100), (0, 4097), (1, 4097), (0, 70000), (1, 70000), (1, 300001)])
def test_parallel_parse_matches_oracle(lz, extra, n):
    ttype = po.T_LZX if extra else po.T_LZ
    for cname, x in cases(n, n):
        x = np.ascontiguousarray(x, np.uint8)
        for dt in (0, 6) if n < 300000 else (0,):
            want, _ = po.transform_forward(ttype, x, dt)
            for seg in (1024, 8192, 1 << 30):
                r, dst, st = run(lz, extra, dt, x, seg)
                if want is None:
                    assert r == -1, (cname, n, dt, seg, r)
                    continue
                assert r == len(want) and np.array_equal(dst[:r], want), (cname, n, dt, seg, r, len(want), st)


def test_large_blocks_and_stitch_work(lz):
    """2 MiB blocks: byte exact, and the stitch pass re-parses only a small part of a text block"""
    n = 2 << 20
    for cname, x in cases(n, 11):
        if cname not in ("text", "mixed", "sparse", "x86", "records", "uniform"):
            continue
        x = np.ascontiguousarray(x, np.uint8)
        want, _ = po.transform_forward(po.T_LZX, x, 0)
        r, dst, st = run(lz, 1, 0, x, 1024)  # SEG of kz_lz_par.cu
        if want is None:
            assert r == -1, cname
            continue
        assert r == len(want) and np.array_equal(dst[:r], want), (cname, r, len(want), st)
        print(cname, "matches", st[0], "fix-up (all rounds)", st[1], "pieces", st[2], "rounds", st[3])
        if cname == "text":
            assert st[1] < st[0] // 10 and st[3] <= 4


def test_junction_cap_falls_back_to_the_serial_stitch(lz):
    """a junction that needs more fix-up matches than its cap sends the block to the serial stitch: same bytes"""
    n = 300000
    hit = 0
    try:
        lz.lz_set_jcap(1)
        for cname, x in cases(n, 5):
            x = np.ascontiguousarray(x, np.uint8)
            want, _ = po.transform_forward(po.T_LZX, x, 0)
            r, dst, st = run(lz, 1, 0, x, 1024)
            if want is None:
                assert r == -1, cname
                continue
            assert r == len(want) and np.array_equal(dst[:r], want), (cname, r, len(want), st)
            hit += st[5]
    finally:
        lz.lz_set_jcap(64)
    assert hit > 0
