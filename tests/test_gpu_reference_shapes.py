"""The input shapes of the reference's own transform test (v2/transform/Transforms_test.go:163-259 generateTransformTestCases: empty, 1-2 bytes,
all 256 values, the specific sequences, 80 000 eights, zero-heavy blocks of 512..2048 bytes, 20 leading zeros + random, run-length blocks), seeded,
through every transform the GPU path has: forward == oracle byte for byte (or both decline), GPU inverse of the oracle's output == input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ["BWT", "BWTS", "LZ", "LZX", "ROLZ", "RANK", "MTFT", "ZRLT", "PACK", "DNA", "MM", "TEXT", "UTF", "EXE"]


def shapes(name, seed=20260923):
    rng = np.random.default_rng(seed)
    small = 5 if name == "ZRLT" else None
    yield "EmptyInput", np.zeros(0, np.uint8)
    yield "SingleByteA", np.frombuffer(b"A", np.uint8)
    yield "TwoIdenticalBytesAA", np.frombuffer(b"AA", np.uint8)
    yield "TwoDifferentBytesAB", np.frombuffer(b"AB", np.uint8)
    yield "All256ByteValues", np.arange(256, dtype=np.uint8)
    yield "SpecificSequence_0", np.array([0, 1, 2, 2, 2, 2, 7, 9, 9, 16, 16, 16, 1] + [3] * 19, np.uint8)
    eights = np.full(80000, 8, np.uint8)
    eights[0] = 1
    yield "AllEights_OneOne_80k", eights
    yield "ShortRepeats_2", np.array([0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3], np.uint8)
    for i in range(3, 6):
        v = rng.integers(0, small or 100, 1 << (i + 6)).astype(np.uint8)
        v[v >= 33] = 0
        yield "LotsOfZeros_%d" % i, v
    x = rng.integers(0, small or 256, 512).astype(np.uint8)
    x[:20] = 0
    yield "Random_WithInitialZeros", x
    for i in range(7, 50):
        out = np.zeros(1024, np.uint8)
        idx = 20
        while idx < 1024:
            length = int(rng.integers(0, 120))
            if length % 3 == 0:
                length = 1
            out[idx: idx + length] = rng.integers(0, small or 256)
            idx += length
        yield "RandomLengthsRandomValues_%d" % i, out


@pytest.mark.parametrize("name", NAMES)
def test_reference_transform_shapes(gpu, oracle, kz, name):
    t = kz.TRANSFORM_IDS[name]
    for cname, x in shapes(name):
        if len(x) == 0:
            continue  # the reference returns (0, 0, nil) before looking at anything; the ABI takes no empty blocks
        want, wdt = oracle.transform_forward(t, x)
        got, gdt = gpu.transform_forward(t, x)
        assert (want is None) == (got is None), (name, cname)
        if want is None:
            continue
        assert gdt == wdt and len(got) == len(want) and np.array_equal(got, want), (name, cname, len(got), len(want))
        back = gpu.transform_inverse(t, want, len(x) + 4096)
        assert np.array_equal(back, x), (name, cname)
