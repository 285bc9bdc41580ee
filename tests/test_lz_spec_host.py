"""Next-round preparation, CPU only: a model of the warp-speculative LZ / LZX parse (tests/host/lz_spec_host.cpp: 32 positions of a
literal run probed per batch, hash-slot hazards inside the batch forwarded lane to lane) produces exactly the oracle's bytes, and the
number of serial steps drops from one per visited position to one per batch."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

synth = importlib.import_module("kanzi-go_b200.synth")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "lz_spec_host.cpp")
SO = os.path.join(ROOT, "tests", "host", "_build", "liblz_spec_host.so")


@pytest.fixture(scope="module")
def lz():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.lz_forward_spec.restype = C.c_int64
    lib.lz_forward_spec.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    return lib


def cases(n, seed):
    r = np.random.default_rng(seed)
    yield "text", synth.markov_text(n, seed=seed)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 1)
    yield "zipf", synth.zipf_bytes(n, 1.5, seed=seed + 2)
    yield "periodic", np.resize(np.frombuffer(b"abracadabra, ", np.uint8), n)
    yield "runs", np.repeat(r.integers(0, 256, n // 37 + 1).astype(np.uint8), 37)[:n]  # same hash many times inside a batch
    yield "const", np.full(n, 7, np.uint8)
    yield "dna", np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, n)]
    mixed = synth.markov_text(n, seed=seed + 3).copy()
    k = n // 3
    mixed[k: 2 * k] = r.integers(0, 256, k).astype(np.uint8)  # a long incompressible stretch: the skip acceleration (srcInc >> 6) kicks in
    yield "mixed", mixed


@pytest.mark.parametrize("extra", [0, 1])
@pytest.mark.parametrize("n", [24, 100, 4097, 70000, 300001, 2 << 20])
def test_speculative_parse_matches_oracle(lz, extra, n):
    ttype = po.T_LZX if extra else po.T_LZ
    for cname, x in cases(n, n):
        x = np.ascontiguousarray(x, np.uint8)
        for dt in (0, 6):
            want, _ = po.transform_forward(ttype, x, dt)
            dst = np.zeros(n + n // 64 + 64, np.uint8)
            stats = (C.c_int64 * 3)(0, 0, 0)
            r = lz.lz_forward_spec(extra, dt, x.ctypes.data, n, dst.ctypes.data, len(dst), stats)
            if want is None:
                assert r == -1, (cname, n, dt)
                continue
            assert r == len(want) and np.array_equal(dst[:r], want), (cname, n, dt, r, len(want))


def test_speculation_cuts_the_serial_steps(lz):
    n = 4 << 20
    for kind, x in (("text", synth.markov_text(n, seed=5)), ("uniform", synth.uniform_bytes(n, seed=6))):
        x = np.ascontiguousarray(x, np.uint8)
        dst = np.zeros(n + n // 64 + 64, np.uint8)
        stats = (C.c_int64 * 3)(0, 0, 0)
        lz.lz_forward_spec(1, 0, x.ctypes.data, n, dst.ctypes.data, len(dst), stats)
        batches, literal_probes, matches = stats[0], stats[1], stats[2]
        print(kind, "batches", batches, "literal probes", literal_probes, "matches", matches, "probes per batch %.2f" % (literal_probes / batches))
        assert batches < literal_probes
