"""Complete kanzi levels at BASELINE sizes ("no level chain is tested at BASELINE sizes"): the whole stream the GPU
path writes must equal the oracle's stream byte for byte, and the GPU must decode the oracle's stream back to the input.

 * BASELINE.json metric configuration: 200,000,000-byte silesia.tar-shaped slab (synth.silesia_shaped), 4 MiB
   blocks (48 blocks), -l 3 = TEXT+UTF+PACK+MM+LZX & HUFFMAN and -l 5 = TEXT+UTF+BWT+RANK+ZRLT & ANS0 (app/BlockCompressor.go:665-700)
 * configs[3] shape: one 32 MiB block of enwik-shaped text at -l 5
 * the workload that failed its round trip at the end of round 1: 16 x 4 MiB of markov text at -l 5 (block 6 of it ends up with
   1,425,035 bytes after TEXT: a BWT length whose head count is 3 modulo 256), and such a BWT block on its own
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LEVELS = {"l1": ("LZX", "NONE"), "l2": ("DNA+LZ", "HUFFMAN"), "l3": ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"),
          "l4": ("TEXT+UTF+EXE+PACK+MM+ROLZ", "NONE"), "l5": ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0")}


@pytest.fixture(scope="module")
def silesia(synth):
    return synth.silesia_shaped(200_000_000, seed=synth.SEED + 1)


def _first_diff(a, b):
    k = min(len(a), len(b))
    d = np.nonzero(a[:k] != b[:k])[0]
    return int(d[0]) if len(d) else k


@pytest.mark.parametrize("level", ["l3", "l5"])
def test_silesia_200mb_stream_parity(gpu, oracle, silesia, level):
    tr, en = LEVELS[level]
    x = silesia
    n = len(x)
    want = oracle.compress(x, tr, en, block_size=4 << 20, jobs=64, input_size=n)
    got = gpu.compress(x, tr, en, block_size=4 << 20, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (level, len(got), len(want), _first_diff(got, want))
    y = gpu.decompress(want, n + 64)
    assert len(y) == n and np.array_equal(y, x), (level, _first_diff(y, x))


@pytest.mark.parametrize("level", ["l1", "l2", "l4"])
def test_silesia_40mb_other_levels(gpu, oracle, silesia, level):
    """the other levels the GPU path covers, on the first 40 MB of the slab (text, x86-like and 16-bit segments)"""
    tr, en = LEVELS[level]
    x = silesia[:40_000_000]
    n = len(x)
    want = oracle.compress(x, tr, en, block_size=4 << 20, jobs=64, input_size=n)
    got = gpu.compress(x, tr, en, block_size=4 << 20, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (level, len(got), len(want), _first_diff(got, want))
    y = gpu.decompress(want, n + 64)
    assert len(y) == n and np.array_equal(y, x), (level, _first_diff(y, x))


def test_enwik_32mib_block_l5(gpu, oracle, synth):
    n = 32 << 20
    x = synth.enwik_shaped(n)
    tr, en = LEVELS["l5"]
    want = oracle.compress(x, tr, en, block_size=32 << 20, jobs=8, input_size=n)
    got = gpu.compress(x, tr, en, block_size=32 << 20, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (len(got), len(want), _first_diff(got, want))
    y = gpu.decompress(want, n + 64)
    assert len(y) == n and np.array_equal(y, x)


@pytest.mark.parametrize("level", ["l3", "l5"])
def test_markov_16x4mib(gpu, oracle, synth, level):
    """bench.py --workload l5 of round 1 (seed SEED + 1): failed its round trip before the ibwt_measure_kernel fix"""
    n = 16 * (4 << 20)
    x = synth.markov_text(n, seed=synth.SEED + 1)
    tr, en = LEVELS[level]
    want = oracle.compress(x, tr, en, block_size=4 << 20, jobs=16, input_size=n)
    got = gpu.compress(x, tr, en, block_size=4 << 20, input_size=n)
    assert len(got) == len(want) and np.array_equal(got, want), (level, len(got), len(want), _first_diff(got, want))
    for _ in range(2):
        y = gpu.decompress(got, n + 64)
        assert len(y) == n and np.array_equal(y, x), (level, _first_diff(y, x))


def test_bwt_block_1425035(gpu, oracle, synth, kz):
    """a BWT block of 1,425,035 bytes: ceil(n / 64) + 8 = 22,275 sub-list heads = 3 modulo 256"""
    n = 1425035
    for seed in (1, 2):
        x = synth.markov_text(n, seed=seed)
        want, _ = oracle.transform_forward(oracle.T_BWT, x)
        got, _ = gpu.transform_forward(kz.T_BWT, x)
        assert np.array_equal(got, want)
        assert np.array_equal(gpu.transform_inverse(kz.T_BWT, want, n + 64), x)
    x = np.concatenate([synth.markov_text(n, seed=3), synth.zipf_bytes(n, 1.2, seed=4), synth.markov_text(n, seed=5)])
    want = oracle.compress(x, "BWT+RANK+ZRLT", "ANS0", block_size=n + (16 - n % 16) % 16, jobs=3, input_size=len(x))
    got = gpu.compress(x, "BWT+RANK+ZRLT", "ANS0", block_size=n + (16 - n % 16) % 16, input_size=len(x))
    assert np.array_equal(got, want)
    assert np.array_equal(gpu.decompress(want, len(x) + 64), x)
