"""Deterministic synthetic inputs shared by tests/ and bench.py. numpy only, seeded."""
import numpy as np

SEED = 0x4B414E5A  # the reference's own test seed (v2/io/CompressedStream_test.go:31)


def zipf_bytes(n, s=1.0, seed=SEED, alphabet=256):
    """n bytes, P(value = r) ~ 1/(r+1)^s over `alphabet` values (identity rank -> value)."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.power(np.arange(1, alphabet + 1, dtype=np.float64), s)
    cdf = np.cumsum(w / w.sum())
    cdf[-1] = 1.0
    out = np.empty(n, np.uint8)
    step = 1 << 24
    for o in range(0, n, step):
        m = min(step, n - o)
        out[o:o + m] = np.searchsorted(cdf, rng.random(m), side="right").astype(np.uint8)
    return out


def uniform_bytes(n, seed=SEED):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def markov_text(n, seed=SEED, order_alphabet=96):
    """English-like order-1 Markov bytes over printable ASCII with word/line structure (cheap, vectorised by blocks)."""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10)).astype(np.uint8)) for _ in range(4096)]
    probs = 1.0 / np.arange(1, len(words) + 1)
    probs /= probs.sum()
    out = bytearray()
    while len(out) < n:
        idx = rng.choice(len(words), 4096, p=probs)
        line = 0
        for i in idx:
            out += words[i]
            line += len(words[i]) + 1
            if line > 72:
                out += b"\r\n"
                line = 0
            else:
                out += b" "
    return np.frombuffer(bytes(out[:n]), np.uint8).copy()


def runs_bytes(n, seed=SEED):
    """run-structured data like the reference's benchmarks (v2/benchmark/Entropy_test.go:101-170)."""
    rng = np.random.default_rng(seed)
    vals = rng.integers(0, 256, n // 4 + 8, dtype=np.uint8)
    lens = rng.integers(1, 9, n // 4 + 8)
    return np.repeat(vals, lens)[:n].copy()


def reference_test_inputs(seed=1234567):
    """The input shapes of the reference's entropy round-trip test (v2/entropy/Entropy_test.go:590-806), seeded."""
    rng = np.random.default_rng(seed)
    cases = [
        np.full(40, 65, np.uint8),
        np.array([0x3d, 0x4d, 0x54, 0x47, 0x5a, 0x36, 0x39, 0x26, 0x72, 0x6f, 0x6c, 0x65, 0x3d, 0x32, 0x26, 0x67], np.uint8),
        np.zeros(0, np.uint8),
        np.arange(256, dtype=np.uint8),
        np.full(1024, ord("*"), np.uint8),
        np.tile(np.array([65, 66], np.uint8), 512),
        rng.integers(0, 256, 4096, dtype=np.uint8),
        (rng.integers(0, 256, 4096, dtype=np.uint8) * (rng.random(4096) < 0.1)).astype(np.uint8),
    ]
    for _ in range(13):
        cases.append((65 + rng.integers(0, 4 + 3 * _, 256)).astype(np.uint8))
    return cases
