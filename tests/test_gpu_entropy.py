"""GPU parity tests for the entropy codecs (order-0 rANS, Huffman): the CUDA encoder must emit the oracle's bits exactly, the CUDA decoder
must invert the oracle's streams, through every level of the C ABI (single block codec, block batch, whole stream)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    if len(d):
        i = int(d[0])
        return "first diff at byte %d (bit %d): got %s want %s (lens %d / %d)" % (i, 8 * i, a[i:i + 8].tolist(), b[i:i + 8].tolist(), len(a), len(b))
    return "lengths differ: %d vs %d" % (len(a), len(b))


def shapes(synth, n, seed):
    yield "zipf1", synth.zipf_bytes(n, 1.0, seed=seed)
    yield "zipf2", synth.zipf_bytes(n, 2.0, seed=seed + 1)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 2)
    yield "text", synth.markov_text(n, seed=seed + 3)
    yield "const", np.full(n, 200, np.uint8)
    yield "two", (synth.uniform_bytes(n, seed=seed + 4) & 1).astype(np.uint8) * 255
    yield "small_alpha", (synth.zipf_bytes(n, 1.2, seed=seed + 5, alphabet=40) + 50).astype(np.uint8)
    yield "ramp", (np.arange(n) % 256).astype(np.uint8)
    # Fibonacci-like counts force Huffman codes longer than 12 bits (limitCodeLengths path)
    fib = [1, 1]
    while sum(fib) < 14000:
        fib.append(fib[-1] + fib[-2])
    pat = np.repeat(np.arange(len(fib), dtype=np.uint8) * 7 + 3, fib)
    rng = np.random.default_rng(seed + 9)
    rng.shuffle(pat)
    yield "fib", np.resize(pat, n)
    geo = np.minimum(rng.geometric(0.35, n), 40).astype(np.uint8)
    yield "geometric", geo


SIZES = [1, 31, 32, 33, 36, 37, 38, 39, 100, 4095, 4096, 4097, 16383, 16384, 16385, 16386, 16387, 16388, 50000, 65536 + 19, 262144, 1 << 20]


CODECS = ["ANS0", "HUFFMAN", "ANS1", "RANGE"]


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("n", SIZES)
def test_entropy_codec_parity(gpu, oracle, synth, kz, n, codec):
    et = oracle.parse_entropy(codec)
    for name, x in shapes(synth, n, seed=n):
        want, want_bits = oracle.entropy_encode(et, x)
        got, got_bits = gpu.entropy_encode(et, x)
        assert got_bits == want_bits, (codec, name, n, got_bits, want_bits, first_diff(got, want))
        assert np.array_equal(got, want), (codec, name, n, first_diff(got, want))
        y, used = gpu.entropy_decode(et, want, n)
        assert np.array_equal(y, x), (codec, name, n, first_diff(y, x))
        assert used == want_bits, (codec, name, n, used, want_bits)


@pytest.mark.parametrize("codec", CODECS)
def test_reference_test_inputs(gpu, oracle, synth, kz, codec):
    et = oracle.parse_entropy(codec)
    for x in synth.reference_test_inputs():
        if len(x) == 0:
            continue
        want, want_bits = oracle.entropy_encode(et, x)
        got, got_bits = gpu.entropy_encode(et, x)
        assert got_bits == want_bits and np.array_equal(got, want), first_diff(got, want)
        y, _ = gpu.entropy_decode(et, want, len(x))
        assert np.array_equal(y, x)


def test_mirror_interfaces(gpu, oracle, synth, kz):
    """EntropyEncoder.Write / EntropyDecoder.Read mirrors (v2/Definitions.go:154-179)."""
    x = synth.zipf_bytes(100000, 1.0, seed=9)
    obs = kz.OutputBitStream()
    ee = kz.EntropyEncoder(obs, "ANS0", ctx=gpu)
    assert ee.write(x) == len(x)
    ee.dispose()
    want, want_bits = oracle.entropy_encode(oracle.E_ANS0, x)
    assert obs.written() == want_bits
    assert np.array_equal(obs.to_bytes(), want)
    ed = kz.EntropyDecoder(want, "ANS0", ctx=gpu)
    assert np.array_equal(ed.read(len(x)), x)
    assert ed.consumed == want_bits


@pytest.mark.parametrize("entropy", ["ANS0", "NONE", "HUFFMAN", "ANS1", "RANGE"])
def test_block_batch_parity(gpu, oracle, synth, kz, entropy):
    """kz_encode_blocks / kz_decode_blocks against encodingTask.encode / decodingTask.decode restated in the oracle."""
    lens = [1, 15, 16, 33, 255, 256, 1000, 65535, 65536, 65537, 100000, 16384 * 3, 5, 70000]
    blocks = [synth.zipf_bytes(n, 1.0, seed=100 + i) for i, n in enumerate(lens)]
    et = oracle.parse_entropy(entropy)
    got = gpu.encode_blocks(0, et, blocks)
    for i, (b, (g, gbits)) in enumerate(zip(blocks, got)):
        w, wbits = oracle.encode_block(b, 0, et)
        assert gbits == wbits, (i, lens[i], gbits, wbits)
        assert np.array_equal(g, w), (i, lens[i], first_diff(g, w))
    payloads = [oracle.encode_block(b, 0, et) for b in blocks]
    dec = gpu.decode_blocks(0, et, payloads, block_size=1 << 20)
    for i, (b, d) in enumerate(zip(blocks, dec)):
        assert np.array_equal(b, d), (i, lens[i], first_diff(d, b))


@pytest.mark.parametrize("n,bs", [(0, 1024), (5, 1024), (1024, 1024), (4096 + 7, 1024), (1 << 20, 65536), (3 * (1 << 20) + 12345, 1 << 20), (40 << 20, 4 << 20)])
@pytest.mark.parametrize("entropy", ["ANS0", "NONE", "HUFFMAN", "ANS1", "RANGE"])
def test_stream_parity(gpu, oracle, synth, kz, n, bs, entropy):
    """Whole CompressedOutputStream / CompressedInputStream byte parity (header, block prefixes, end marker)."""
    x = synth.zipf_bytes(n, 1.0, seed=n + 17) if n else np.zeros(0, np.uint8)
    want = oracle.compress(x, "NONE", entropy, block_size=bs, jobs=8, input_size=n)
    got = gpu.compress(x, "NONE", entropy, block_size=bs, input_size=n)
    assert len(got) == len(want), (len(got), len(want), first_diff(got, want))
    assert np.array_equal(got, want), first_diff(got, want)
    y = gpu.decompress(want, n + 64)
    assert np.array_equal(y, x), first_diff(y, x)
    w = kz.Writer("NONE", entropy, block_size=bs, file_size=n, ctx=gpu)
    w.write(x[: n // 2])
    w.write(x[n // 2:])
    w.close()
    assert np.array_equal(w.getvalue(), want)
    assert np.array_equal(kz.Reader(want, ctx=gpu).read_all(n + 64), x)


def test_config2_full_size(gpu, oracle, synth, kz):
    """BASELINE config 2: -e ANS0 -t NONE, 256 MiB Zipf(1.0), 64 x 4 MiB blocks: stream parity + round trip."""
    n = 256 << 20
    x = synth.zipf_bytes(n, 1.0, seed=synth.SEED + 1)
    got = gpu.compress(x, "NONE", "ANS0", block_size=4 << 20, input_size=n)
    y = gpu.decompress(got, n + 64)
    assert np.array_equal(y, x)
    want = oracle.compress(x, "NONE", "ANS0", block_size=4 << 20, jobs=8, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), first_diff(got, want)


def test_ans1_chunk_boundary(gpu, oracle, synth, kz):
    """order-1 chunks are 4 MiB: two chunks in one block, and the reference's 2/3-byte chunk panic."""
    n = (4 << 20) + 1000
    x = synth.markov_text(n, seed=21)
    want, wbits = oracle.entropy_encode(oracle.E_ANS1, x)
    got, gbits = gpu.entropy_encode(kz.E_ANS1, x)
    assert gbits == wbits and np.array_equal(got, want), first_diff(got, want)
    y, used = gpu.entropy_decode(kz.E_ANS1, want, n)
    assert np.array_equal(y, x) and used == wbits
    bad = synth.markov_text((4 << 20) + 2, seed=22)
    with pytest.raises(oracle.OracleError):
        oracle.entropy_encode(oracle.E_ANS1, bad)
    with pytest.raises(kz.KanziError):
        gpu.entropy_encode(kz.E_ANS1, bad)


def test_config1_huffman_4mb(gpu, oracle, synth, kz):
    """BASELINE config 1: -e HUFFMAN -t NONE, 4 MiB Zipf(1.0) bytes, 1 block."""
    x = synth.zipf_bytes(4 << 20, 1.0, seed=synth.SEED)
    want = oracle.compress(x, "NONE", "HUFFMAN", block_size=4 << 20, input_size=len(x))
    got = gpu.compress(x, "NONE", "HUFFMAN", block_size=4 << 20, input_size=len(x))
    assert np.array_equal(got, want), first_diff(got, want)
    assert np.array_equal(gpu.decompress(want, len(x) + 64), x)


@pytest.mark.parametrize("codec", CODECS)
def test_corrupt_streams_do_not_crash(gpu, oracle, synth, kz, codec):
    x = synth.zipf_bytes(200000, 1.0, seed=5)
    s = oracle.compress(x, "NONE", codec, block_size=65536, jobs=2)
    rng = np.random.default_rng(7)
    for trial in range(12):
        t = s.copy()
        pos = int(rng.integers(21, len(t)))
        t[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            y = gpu.decompress(t, len(x) + 64)
            assert len(y) <= len(x) + 64
        except kz.KanziError:
            pass
    t = s[: len(s) // 2]
    with pytest.raises(kz.KanziError):
        gpu.decompress(t, len(x) + 64)
    # the engine must still be healthy afterwards
    assert np.array_equal(gpu.decompress(s, len(x) + 64), x)
