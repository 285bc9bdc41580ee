"""GPU parity tests for ROLZ (rolzCodec1, transform/ROLZCodec.go:301-913): the serial match-table replay with its nested rANS
streams (order 0 below 128 KiB, order 1 above; 32 KiB chunks for tokens / lengths / match indexes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T_ROLZ = 11


def cases(synth, n, seed):
    yield "text", synth.markov_text(n, seed=seed)
    yield "uniform", synth.uniform_bytes(n, seed=seed + 1)
    yield "periodic", np.resize(np.frombuffer(b"abracadabra, ", np.uint8), n)
    yield "zipf", synth.zipf_bytes(n, 1.5, seed=seed + 2)
    yield "const", np.full(n, 7, np.uint8)
    dna = np.frombuffer(b"ACGT", np.uint8)[synth.uniform_bytes(n, seed=seed + 3) & 3]
    yield "dna", dna  # DetectSimpleType -> DT_DNA: 8-byte keys, minimum match 7


@pytest.mark.parametrize("n", [10, 63, 64, 65, 100, 513, 1000, 5000, 70000, 131071, 131072, 200000, 600000])
def test_rolz_transform_parity(gpu, oracle, synth, kz, n):
    for cname, x in cases(synth, n, seed=n):
        want, wdt = oracle.transform_forward(T_ROLZ, x)
        got, gdt = gpu.transform_forward(T_ROLZ, x)
        if want is None:
            assert got is None, (cname, n)
            continue
        assert got is not None, (cname, n)
        assert len(got) == len(want) and np.array_equal(got, want), (cname, n, len(got), len(want),
                                                                      int(np.argmax(got[: min(len(got), len(want))] != want[: min(len(got), len(want))])))
        assert gdt == wdt, (cname, n, gdt, wdt)
        back = gpu.transform_inverse(T_ROLZ, want, n + 64)
        assert np.array_equal(back, x), (cname, n)


@pytest.mark.parametrize("entropy", ["NONE", "ANS0"])
@pytest.mark.parametrize("n,bs", [(5000, 1024), (300000, 65536), (2 * (1 << 20) + 77, 1 << 20)])
def test_stream_parity_rolz(gpu, oracle, synth, kz, entropy, n, bs):
    """-t ROLZ (the last stage of kanzi -l 4) through the whole stream path"""
    for kind in ("text", "uniform"):
        x = synth.markov_text(n, seed=n + 5) if kind == "text" else synth.uniform_bytes(n, seed=n + 6)
        want = oracle.compress(x, "ROLZ", entropy, block_size=bs, jobs=8, input_size=n)
        got = gpu.compress(x, "ROLZ", entropy, block_size=bs, input_size=n)
        assert len(got) == len(want) and np.array_equal(got, want), (entropy, n, bs, kind, len(got), len(want))
        assert np.array_equal(gpu.decompress(want, n + 64), x), (entropy, n, bs, kind)
