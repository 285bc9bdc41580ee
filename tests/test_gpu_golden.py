"""Reference-written streams (tests/golden/ref_<case>.knz, produced by tools/dump_ref.go where a Go toolchain and the kanzi-go sources exist):
the GPU path must write the same bytes and must decode them back to the input. None exists in the build image (no Go: DESIGN.md section 2,
"parity unpinned"), so the loop normally finds nothing; the oracle-side twin is tests/test_oracle_pins.py::test_reference_golden_streams."""
import os

import numpy as np
import pytest

from test_oracle_pins import GOLDEN_CASES, golden_input

pytestmark = pytest.mark.gpu


def test_gpu_against_reference_streams(kz, gpu, oracle):
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    found = 0
    for name, transform, entropy, bs, ck, n in GOLDEN_CASES:
        x = golden_input(n)
        path = os.path.join(here, "ref_%s.knz" % name)
        if os.path.exists(path):
            found += 1
            want = np.fromfile(path, np.uint8)
        else:  # no reference stream: at least the oracle's, so that every case of the list runs through the GPU path
            want = oracle.compress(x, transform, entropy, block_size=bs, checksum_bits=ck, jobs=4, input_size=n)
        try:
            got = gpu.compress(x, transform, entropy, block_size=bs, checksum_bits=ck, input_size=n)
        except kz.KanziError as e:
            pytest.fail("%s: %s" % (name, e))
        assert len(got) == len(want) and np.array_equal(got, want), name
        y = gpu.decompress(want, n + 64)
        assert len(y) == n and np.array_equal(y, x), name
    assert found in (0, len(GOLDEN_CASES)) or found > 0
