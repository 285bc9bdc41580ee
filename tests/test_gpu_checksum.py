"""GPU parity tests for block checksums (-x 32 / -x 64): XXHash32/XXHash64 of every block computed on the device
(hash/XXHash32.go, hash/XXHash64.go; io/CompressedStream.go:753-763, :1990-2006)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("transform,entropy", [("NONE", "ANS0"), ("NONE", "NONE"), ("LZX", "HUFFMAN"), ("BWT", "ANS0")])
def test_stream_with_checksums(gpu, oracle, synth, kz, bits, transform, entropy):
    for n, bs in [(5, 1024), (1000, 1024), (65536 + 77, 65536), (3 * 65536 + 31, 65536), (2 * (1 << 20) + 5, 1 << 20)]:
        x = synth.markov_text(n, seed=n + bits)
        want = oracle.compress(x, transform, entropy, block_size=bs, checksum_bits=bits, jobs=4, input_size=n)
        got = gpu.compress(x, transform, entropy, block_size=bs, checksum_bits=bits, input_size=n)
        assert len(got) == len(want) and np.array_equal(got, want), (n, bs, bits, transform, entropy)
        assert np.array_equal(gpu.decompress(want, n + 64), x)


@pytest.mark.parametrize("bits", [32, 64])
def test_block_batch_with_checksums(gpu, oracle, synth, kz, bits):
    lens = [1, 15, 16, 31, 32, 33, 100, 4096, 65536, 65537, 100003]
    blocks = [synth.zipf_bytes(n, 1.0, seed=300 + i) for i, n in enumerate(lens)]
    et = oracle.parse_entropy("ANS0")
    got = gpu.encode_blocks(0, et, blocks, checksum_bits=bits)
    payloads = []
    for i, (b, (g, gbits)) in enumerate(zip(blocks, got)):
        w, wbits = oracle.encode_block(b, 0, et, checksum_bits=bits)
        assert gbits == wbits and np.array_equal(g, w), (i, lens[i])
        payloads.append((w, wbits))
    dec = gpu.decode_blocks(0, et, payloads, block_size=1 << 20, checksum_bits=bits)
    for b, d in zip(blocks, dec):
        assert np.array_equal(b, d)


@pytest.mark.parametrize("bits", [32, 64])
def test_checksum_mismatch_is_reported(gpu, oracle, synth, kz, bits):
    """A flipped payload bit of a NONE/NONE stream decodes 'successfully' but must fail the CRC check (ERR_CRC_CHECK = 19)."""
    n = 100000
    x = synth.zipf_bytes(n, 1.0, seed=bits)
    s = oracle.compress(x, "NONE", "NONE", block_size=65536, checksum_bits=bits, jobs=2, input_size=n)
    t = s.copy()
    t[len(t) // 2] ^= 0x10
    with pytest.raises(kz.KanziError) as ei:
        gpu.decompress(t, n + 64)
    assert ei.value.code == 19
    assert np.array_equal(gpu.decompress(s, n + 64), x)
