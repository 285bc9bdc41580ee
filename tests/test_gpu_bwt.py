"""GPU parity tests for BWT.Forward / BWT.Inverse (+ BWTBlockCodec framing) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cases(synth, n, seed):
    yield "text", synth.markov_text(n, seed=seed)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 1)
    yield "const", np.full(n, 65, np.uint8)
    yield "abab", np.resize(np.array([1, 2], np.uint8), n)
    yield "zipf", synth.zipf_bytes(n, 1.5, seed=seed + 2)
    yield "ramp", (np.arange(n) % 251).astype(np.uint8)
    rep = synth.markov_text(max(n // 7, 1), seed=seed + 3)
    yield "repeat7", np.resize(rep, n)
    z = synth.uniform_bytes(n, seed=seed + 4)
    z[z < 250] = 0
    yield "zeros", z


@pytest.mark.parametrize("n", [2, 3, 11, 255, 256, 257, 1000, 4096, 65536, 100003, 1 << 20])
def test_bwt_block_codec_parity(gpu, oracle, synth, kz, n):
    for name, x in cases(synth, n, seed=n):
        want, _ = oracle.transform_forward(oracle.T_BWT, x)
        got, _ = gpu.transform_forward(kz.T_BWT, x)
        assert got is not None and want is not None
        assert np.array_equal(got, want), (name, n, int(np.argmax(got[: len(want)] != want)) if len(got) == len(want) else (len(got), len(want)))
        back = gpu.transform_inverse(kz.T_BWT, want, n + 64)
        assert np.array_equal(back, x), (name, n)


def test_bwt_doc_example(gpu, kz):
    got, _ = gpu.transform_forward(kz.T_BWT, b"mississippi")
    assert bytes(got[2:]) == b"ipssmpissii" and got[1] == 4 and got[0] == 0  # primary index 5 stored as 5 - 1


def test_byte_transform_mirror(gpu, oracle, synth, kz):
    x = synth.markov_text(300000, seed=3)
    t = kz.ByteTransform("BWT", ctx=gpu)
    f = t.forward(x)
    want, _ = oracle.transform_forward(oracle.T_BWT, x)
    assert np.array_equal(f, want)
    assert np.array_equal(t.inverse(f, len(x)), x)
    assert t.max_encoded_len(1000) == 1033
    assert kz.ByteTransform("BWT", ctx=gpu).forward(b"x") is None  # 1 byte: the codec declines (invalid index size)


@pytest.mark.parametrize("entropy", ["ANS0", "HUFFMAN", "NONE"])
@pytest.mark.parametrize("n,bs", [(10, 1024), (5000, 1024), (1 << 20, 1 << 18), (3 * (1 << 20) + 777, 1 << 20)])
def test_stream_parity_bwt(gpu, oracle, synth, kz, n, bs, entropy):
    x = synth.markov_text(n, seed=n + 1)
    want = oracle.compress(x, "BWT", entropy, block_size=bs, jobs=8, input_size=n)
    got = gpu.compress(x, "BWT", entropy, block_size=bs, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (len(got), len(want))
    assert np.array_equal(gpu.decompress(want, n + 64), x)


def test_block_batch_bwt(gpu, oracle, synth, kz):
    lens = [1, 15, 16, 300, 70000, 200000]
    blocks = [synth.markov_text(m, seed=50 + i) for i, m in enumerate(lens)]
    t48 = kz.transform_type("BWT")
    got = gpu.encode_blocks(t48, kz.E_ANS0, blocks)
    for i, (b, (g, gbits)) in enumerate(zip(blocks, got)):
        w, wbits = oracle.encode_block(b, t48, oracle.E_ANS0)
        assert gbits == wbits and np.array_equal(g, w), (i, lens[i])
    payloads = [oracle.encode_block(b, t48, oracle.E_ANS0) for b in blocks]
    dec = gpu.decode_blocks(t48, kz.E_ANS0, payloads, block_size=1 << 18)
    for b, d in zip(blocks, dec):
        assert np.array_equal(b, d)


def test_config4_block_32mb(gpu, oracle, synth, kz):
    """BASELINE config 4 shape: one 32 MiB text block, BWT + ANS0: GPU stream == oracle stream, and round trip."""
    n = 32 << 20
    x = synth.markov_text(n, seed=77)
    got = gpu.compress(x, "BWT", "ANS0", block_size=n, input_size=n)
    assert np.array_equal(gpu.decompress(got, n + 64), x)
    want = oracle.compress(x, "BWT", "ANS0", block_size=n, jobs=1, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want)


def test_bwt_inverse_corrupt_payload_is_reported_or_decoded(gpu, oracle, synth, kz):
    """a payload whose LF permutation is not one list through all ranks (flipped primary index / data bytes) is an error, never stale device memory"""
    x = synth.markov_text(100000, seed=3)
    enc, _ = oracle.transform_forward(kz.T_BWT, x)
    rng = np.random.default_rng(5)
    for trial in range(16):
        bad = enc.copy()
        pos = 1 + trial if trial < 8 else int(rng.integers(40, len(bad)))  # the 8 primary index fields, then data bytes
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            y = gpu.transform_inverse(kz.T_BWT, bad, len(x) + 64)
            assert len(y) == len(x)
        except kz.KanziError:
            pass
    back = gpu.transform_inverse(kz.T_BWT, enc, len(x) + 64)  # and the engine is fine afterwards
    assert np.array_equal(back, x)
