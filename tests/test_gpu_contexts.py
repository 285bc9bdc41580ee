"""Contexts are independent: several of them (one CUDA stream and one set of workspaces each), driven by several host threads,
may work on one GPU at the same time - the way goroutines of the Go side would call the C ABI (integration/go/gpu). The streams
they write are the oracle's, whatever runs beside them (bench.py times the -l 3 and -l 5 round trips of a step side by side)."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

JOBS = [("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("LZ", "NONE"), ("BWT+RANK+ZRLT", "ANS0")]


def test_contexts_side_by_side(kz, oracle, synth):
    n = int(os.environ.get("KZ_TEST_CONTEXTS_BYTES", "6000000"))  # smaller under compute-sanitizer
    x = synth.silesia_shaped(n, seed=synth.SEED + 7)
    want = [oracle.compress(x, tr, en, block_size=1 << 20, jobs=8, input_size=n) for tr, en in JOBS]
    ctxs = [kz.Context(0) for _ in JOBS]
    got = [[] for _ in JOBS]
    back = [[] for _ in JOBS]
    errors = []
    gate = threading.Barrier(len(JOBS))

    def worker(i):
        try:
            tr, en = JOBS[i]
            for _ in range(3):
                gate.wait()
                s = ctxs[i].compress(x, tr, en, block_size=1 << 20, input_size=n)
                got[i].append(s)
                back[i].append(ctxs[i].decompress(s, n + 64))
        except BaseException as e:  # noqa: B902 - re-raised below
            errors.append((i, e))
            gate.abort()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(JOBS))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for i, (tr, en) in enumerate(JOBS):
        for s, y in zip(got[i], back[i]):
            assert len(s) == len(want[i]) and np.array_equal(s, want[i]), (tr, en)
            assert len(y) == n and np.array_equal(y, x), (tr, en)
    for c in ctxs:
        c.close()
