"""Runs last on purpose (file name): the confirmation of a fix made after the GPU budget of round 1 was spent — if it were wrong, every
other GPU test has already reported by the time this one fails."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bs", [15936, 16000, 16320, 32 * 1024 - 64])
def test_bwt_inverse_head_count_residues(gpu, oracle, synth, kz, bs):
    """Block lengths whose sub-list head count ceil(n / 64) + 8 is 1..7 modulo the CTA size of ibwt_measure_kernel: the last CTA then has
    fewer live threads than primary ranks to stage. Several different blocks are inverted back to back so that a stale value from the
    previous block would show (round 1 shipped that bug for one session: profiles/README.md)."""
    heads = (bs + 63) // 64 + 8
    assert 1 <= heads % 256 <= 7
    parts = [synth.markov_text(bs, seed=i) if i % 2 else synth.zipf_bytes(bs, 1.2, seed=i) for i in range(12)]
    x = np.concatenate(parts)
    n = len(x)
    want = oracle.compress(x, "BWT", "NONE", block_size=bs, jobs=8, input_size=n)
    got = gpu.compress(x, "BWT", "NONE", block_size=bs, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want)
    for _ in range(3):  # repeated calls: the shared memory of an SM keeps what the previous launch left there
        assert np.array_equal(gpu.decompress(want, n + 64), x)
    for p in parts:
        w, _ = oracle.transform_forward(oracle.T_BWT, p)
        assert np.array_equal(gpu.transform_inverse(kz.T_BWT, w, bs + 64), p)
