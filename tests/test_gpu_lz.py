"""GPU parity tests for LZCodec (LZ and LZX): forward output and inverse against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cases(synth, n, seed):
    yield "text", synth.markov_text(n, seed=seed)
    yield "runs", synth.runs_bytes(n, seed=seed + 1)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 2)
    yield "const", np.full(n, 65, np.uint8)
    yield "zipf2", synth.zipf_bytes(n, 2.0, seed=seed + 3)
    yield "repeat", np.resize(synth.uniform_bytes(max(n // 9, 1), seed=seed + 4), n)
    far = synth.uniform_bytes(n, seed=seed + 5)
    if n > 300000:
        far[-100000:] = far[:100000]  # a match further than 65534 bytes away (24-bit distances when the block is large enough)
    yield "far", far


@pytest.mark.parametrize("name", ["LZ", "LZX"])
@pytest.mark.parametrize("n", [10, 23, 24, 25, 100, 1000, 4096, 70000, 300001, 1 << 20])
def test_lz_parity(gpu, oracle, synth, kz, n, name):
    t = kz.TRANSFORM_IDS[name]
    for cname, x in cases(synth, n, seed=n):
        want, _ = oracle.transform_forward(t, x)
        got, _ = gpu.transform_forward(t, x)
        assert (got is None) == (want is None), (name, cname, n)
        if want is None:
            continue
        assert len(got) == len(want) and np.array_equal(got, want), (name, cname, n, len(got), len(want))
        back = gpu.transform_inverse(t, want, n + 16)
        assert np.array_equal(back, x), (name, cname, n)


def test_lz_data_type_rules(gpu, oracle, synth, kz):
    """ctx["dataType"]: DNA -> min match 6, SMALL_ALPHABET -> skip (transform/LZCodec.go:298-311)."""
    x = synth.zipf_bytes(100000, 1.5, seed=4, alphabet=4)
    t = kz.T_LZX
    want, _ = oracle.transform_forward(t, x, data_type=6)
    got, _ = gpu.transform_forward(t, x, data_type=6)
    assert want is not None and np.array_equal(got, want)
    assert (got[12] >> 1) & 7 == 4
    assert gpu.transform_forward(t, x, data_type=9)[0] is None
    assert oracle.transform_forward(t, x, data_type=9)[0] is None


@pytest.mark.parametrize("transform,entropy", [("LZX", "HUFFMAN"), ("LZ", "HUFFMAN"), ("LZX", "NONE"), ("LZ", "ANS0")])
@pytest.mark.parametrize("n,bs", [(5000, 1024), (1 << 20, 1 << 18), (3 * (1 << 20) + 777, 1 << 20)])
def test_stream_parity_lz(gpu, oracle, synth, kz, n, bs, transform, entropy):
    """config 3's GPU-covered sub-pipeline: -t LZX -e HUFFMAN (and -l 1 = LZX&NONE)."""
    x = np.concatenate([synth.markov_text(n // 2, seed=n), synth.uniform_bytes(n - n // 2, seed=n + 1)])
    want = oracle.compress(x, transform, entropy, block_size=bs, jobs=8, input_size=n)
    got = gpu.compress(x, transform, entropy, block_size=bs, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (len(got), len(want))
    assert np.array_equal(gpu.decompress(want, n + 64), x)


def test_block_batch_lz(gpu, oracle, synth, kz):
    lens = [1, 15, 16, 23, 24, 300, 70000, 200000]
    blocks = [synth.markov_text(m, seed=60 + i) for i, m in enumerate(lens)]
    t48 = kz.transform_type("LZX")
    got = gpu.encode_blocks(t48, kz.E_HUFFMAN, blocks)
    for i, (b, (g, gbits)) in enumerate(zip(blocks, got)):
        w, wbits = oracle.encode_block(b, t48, oracle.E_HUFFMAN)
        assert gbits == wbits and np.array_equal(g, w), (i, lens[i])
    payloads = [oracle.encode_block(b, t48, oracle.E_HUFFMAN) for b in blocks]
    dec = gpu.decode_blocks(t48, kz.E_HUFFMAN, payloads, block_size=1 << 18)
    for b, d in zip(blocks, dec):
        assert np.array_equal(b, d)
