"""CPU tests of the oracle's pre-transforms (TEXT / UTF / PACK / DNA / MM — oracle side only so far; the
GPU side is next round's work). The reference's own tests for these are round trips (transform/Transforms_test.go:262-425);
on top of that the bit packing / delta coding output is recomputed here with numpy straight from the format description, as an
independent statement of AliasCodec.go:145-207 and FSDCodec.go:262-294."""
import importlib

import numpy as np
import pytest

from oracle import pyoracle as po

synth = importlib.import_module("kanzi-go_b200.synth")
BS = 4 << 20


def _fwd(t, x, dt=0):
    return po.transform_forward(t, np.ascontiguousarray(x, dtype=np.uint8), dt, block_size=BS)


def _inv(t, y, cap):
    return po.transform_inverse(t, y, cap, block_size=BS)


def _roundtrip(t, x, dt=0):
    y, dt_after = _fwd(t, x, dt)
    assert y is not None, "transform skipped"
    z = _inv(t, y, len(x) + 4096)
    z = z[0] if isinstance(z, tuple) else z
    assert np.array_equal(z, x)
    return y, dt_after


def _utf_text(n_words, seed):
    rng = np.random.default_rng(seed)
    words = ["héllo", "wörld", "naïve", "日本語", "テキスト", "данные", "текст", "abc", "xyz", " ", "\n", "😀"]
    return "".join(rng.choice(words, n_words))


# ---------------------------------------------------------------- PACK / DNA
@pytest.mark.parametrize("n", [1024, 1025, 1026, 1027, 50001])
def test_pack_four_symbols_matches_format(n):
    rng = np.random.default_rng(n)
    syms = np.frombuffer(b"ACGT", dtype=np.uint8)
    x = rng.choice(syms, n)
    y, dt = _roundtrip(po.T_PACK, x)
    assert dt == po_dt("DNA")
    # header: n0 = 252 absent symbols, the 4 symbols in increasing order, count & 3, that many raw bytes, then 2 bits per symbol
    c3 = n & 3
    idx = np.searchsorted(np.sort(syms), x[c3:]).astype(np.uint8).reshape(-1, 4)
    packed = (idx[:, 0] << 6) | (idx[:, 1] << 4) | (idx[:, 2] << 2) | idx[:, 3]
    want = np.concatenate([[252], np.sort(syms), [c3], x[:c3], packed]).astype(np.uint8)
    assert np.array_equal(y, want)
    y2, _ = _roundtrip(po.T_DNA, x)  # DNA = PACK restricted to DNA data
    assert np.array_equal(y2, y)


@pytest.mark.parametrize("n", [4096, 4097])
def test_pack_sixteen_symbols_matches_format(n):
    rng = np.random.default_rng(n)
    syms = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    x = rng.choice(syms, n)
    y, _ = _roundtrip(po.T_PACK, x)
    odd = n & 1
    idx = np.searchsorted(np.sort(syms), x[odd:]).astype(np.uint8).reshape(-1, 2)
    want = np.concatenate([[240], np.sort(syms), [odd], x[:odd], (idx[:, 0] << 4) | idx[:, 1]]).astype(np.uint8)
    assert np.array_equal(y, want)


def test_pack_single_symbol_and_skips():
    x = np.full(5000, 0x41, np.uint8)
    y, _ = _roundtrip(po.T_PACK, x)
    assert bytes(y) == bytes([255, 0x41]) + (5000).to_bytes(4, "little")
    assert _fwd(po.T_PACK, np.full(1023, 0x41, np.uint8))[0] is None  # below _ALIAS_MIN_BLOCKSIZE
    rnd = np.random.default_rng(0).integers(0, 256, 1 << 16).astype(np.uint8)
    assert _fwd(po.T_PACK, rnd)[0] is None  # fewer than 16 free byte values
    hexs = np.random.default_rng(1).choice(np.frombuffer(b"0123456789abcdef", dtype=np.uint8), 4096)
    assert _fwd(po.T_DNA, hexs)[0] is None  # DNA only accepts DNA data
    for dt in ("MULTIMEDIA", "UTF8", "EXE", "BIN"):
        assert _fwd(po.T_PACK, x, po_dt(dt))[0] is None


def test_pack_digram_aliasing():
    x = synth.markov_text(1 << 18, seed=11)
    y, _ = _roundtrip(po.T_PACK, x)
    n0 = int(y[0])
    assert 16 <= n0 < 240 and y[1] in (0, 1)
    # the map lists n0 (digram, alias) triples; every alias is a byte value absent from the input
    present = np.zeros(256, bool)
    present[np.unique(x)] = True
    aliases = y[2 + 2 : 2 + 3 * n0 : 3]
    assert not present[aliases].any() and len(set(aliases.tolist())) == n0
    assert len(y) < len(x)


# ---------------------------------------------------------------- MM
def _zigzag(delta):
    return np.where(delta < 0, -2 * delta - 1, 2 * delta).astype(np.uint8)


@pytest.mark.parametrize("kind", ["delta", "xor"])
def test_mm_matches_format(kind):
    rng = np.random.default_rng(5)
    n = 1 << 17
    t = np.arange(n)
    if kind == "delta":  # smooth 8-bit signal: small deltas at distance 1
        x = (128 + 100 * np.sin(t / 40.0) + rng.normal(0, 1.5, n)).astype(np.uint8)
    else:  # interleaved channels with wrap-around jumps: distance 2, large deltas -> xor coding
        a = (t * 37 % 256).astype(np.uint8)
        x = np.empty(n, np.uint8)
        x[0::2] = a[: n // 2]
        x[1::2] = (a[: n // 2] ^ 0x80) + rng.integers(0, 2, n // 2).astype(np.uint8)
    y, dt = _fwd(po.T_MM, x)
    if y is None:
        pytest.skip("sample not accepted by the FSD heuristics")
    assert dt == po_dt("MULTIMEDIA")
    mode, dist = int(y[0]), int(y[1])
    assert dist in (1, 2, 3, 4, 8, 16)
    z = _inv(po.T_MM, y, n)
    z = z[0] if isinstance(z, tuple) else z
    assert np.array_equal(z, x)
    assert np.array_equal(y[2 : 2 + dist], x[:dist])
    if mode == 1:
        assert np.array_equal(y[2 + dist :], x[dist:] ^ x[:-dist])
    else:
        delta = x[dist:].astype(np.int32) - x[:-dist].astype(np.int32)
        esc = (delta < -127) | (delta > 127)
        if not esc.any():
            assert np.array_equal(y[2 + dist :], _zigzag(delta))
        else:  # escaped positions take two bytes: 0xFF, then the xor
            out = []
            for d, e, a, b in zip(delta.tolist(), esc.tolist(), x[dist:].tolist(), x[:-dist].tolist()):
                out += [0xFF, a ^ b] if e else [(-2 * d - 1) if d < 0 else 2 * d]
            assert y[2 + dist :].tolist() == out


def test_mm_skips():
    txt = synth.markov_text(1 << 16, seed=2)
    assert _fwd(po.T_MM, txt)[0] is None  # no gain from any step
    assert _fwd(po.T_MM, txt[:1000])[0] is None  # below _FSD_MIN_BLOCK_LENGTH
    assert _fwd(po.T_MM, txt, po_dt("TEXT"))[0] is None


# ---------------------------------------------------------------- UTF
def test_utf_roundtrip_and_map():
    s = _utf_text(60000, 3)
    x = np.frombuffer(s.encode("utf-8"), dtype=np.uint8)
    y, dt = _roundtrip(po.T_UTF, x)
    assert dt == po_dt("UTF8")
    start, adjust = int(y[0]), int(y[1])
    n = (int(y[2]) << 8) | int(y[3])
    body = x[start : len(x) - 4 + adjust].tobytes().decode("utf-8")
    assert n == len(set(body))  # one map entry per distinct code point of the processed part
    # map entries are ordered by decreasing frequency: the first one is the most frequent code point
    from collections import Counter

    cnt = Counter(body)
    top = max(cnt.items(), key=lambda kv: (kv[1], _pack(kv[0])))[0]
    assert int.from_bytes(y[4:7].tobytes(), "big") == _pack(top)
    assert len(y) < len(x)


def _pack(ch):
    b = ch.encode("utf-8")
    if len(b) == 1:
        return b[0]
    if len(b) == 2:
        return (1 << 19) | (b[0] << 8) | b[1]
    if len(b) == 3:
        return (2 << 19) | ((b[0] & 0x0F) << 12) | ((b[1] & 0x3F) << 6) | (b[2] & 0x3F)
    return (4 << 19) | ((b[0] & 0x07) << 18) | ((b[1] & 0x3F) << 12) | ((b[2] & 0x3F) << 6) | (b[3] & 0x3F)


def test_utf_skips():
    assert _fwd(po.T_UTF, synth.markov_text(1 << 16, seed=1))[0] is None  # pure ASCII: too few multi-byte sequences
    rnd = np.random.default_rng(0).integers(0, 256, 1 << 16).astype(np.uint8)
    assert _fwd(po.T_UTF, rnd)[0] is None
    x = np.frombuffer(_utf_text(5000, 4).encode("utf-8"), dtype=np.uint8)
    assert _fwd(po.T_UTF, x, po_dt("TEXT"))[0] is None  # already typed as something else
    bad = x.copy()
    bad[len(bad) // 2] = 0xC0  # never valid in UTF-8
    assert _fwd(po.T_UTF, bad)[0] is None


# ---------------------------------------------------------------- TEXT (encoding 2)
def test_text_roundtrip_and_flags():
    x = synth.markov_text(1 << 20, seed=3)
    y, dt = _roundtrip(po.T_TEXT, x)
    assert dt == po_dt("TEXT") and (y[0] & 0x80) == 0 and len(y) < len(x)
    # CR LF line ends are folded into LF and flagged
    lf = bytes(x[: 1 << 18]).replace(b"\r", b"")
    xl = np.frombuffer(lf, dtype=np.uint8)
    xc = np.frombuffer(lf.replace(b"\n", b"\r\n"), dtype=np.uint8)
    yl, _ = _roundtrip(po.T_TEXT, xl)
    yc, _ = _roundtrip(po.T_TEXT, xc)
    assert (yc[0] & 0x40) and not (yl[0] & 0x40) and len(xc) > len(xl)
    assert np.array_equal(yc[1:], yl[1:])  # the CR of every CR LF pair is dropped, nothing else changes


def test_text_static_dictionary_word_codes():
    # " the " and friends are in the static dictionary: index + 1 < 64 -> one byte 0x80 | (index + 1); an upper-case first letter
    # is signalled by 0x80 before the index; the space between two dictionary words is implied
    x = np.frombuffer((b"the and that have with " * 200), dtype=np.uint8)
    y, _ = _roundtrip(po.T_TEXT, x)
    the, and_ = y[1], y[2]
    assert the == 0x81  # "the" is word 0 of the dictionary
    assert 0x80 < and_ < 0xC0
    x2 = np.frombuffer((b"The and that have with " * 200), dtype=np.uint8)
    y2, _ = _roundtrip(po.T_TEXT, x2)
    assert y2[1] == 0x80 and y2[2] == 0x81


def test_text_skips():
    rnd = np.random.default_rng(0).integers(0, 256, 1 << 16).astype(np.uint8)
    y, dt = _fwd(po.T_TEXT, rnd)
    assert y is None and dt == po_dt("BIN")
    dna = np.random.default_rng(1).choice(np.frombuffer(b"ACGT", dtype=np.uint8), 1 << 16)
    y, dt = _fwd(po.T_TEXT, dna)
    assert y is None and dt == po_dt("DNA")
    u = np.frombuffer(_utf_text(20000, 5).replace(" ", "").encode("utf-8"), dtype=np.uint8)
    y, dt = _fwd(po.T_TEXT, u)
    assert y is None and dt == po_dt("UTF8")


# ---------------------------------------------------------------- whole streams with the level chains
@pytest.mark.parametrize("chain,entropy", [("DNA+LZ", "HUFFMAN"), ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), ("TEXT+UTF+PACK+MM+ROLZ", "NONE"),
                                           ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0")])
def test_level_chains_roundtrip(chain, entropy):
    rng = np.random.default_rng(9)
    t = np.arange(300000)
    parts = [synth.markov_text(700000, seed=3), np.frombuffer(_utf_text(40000, 6).encode("utf-8"), dtype=np.uint8),
             (128 + 100 * np.sin(t / 50.0) + rng.normal(0, 2, t.size)).astype(np.uint8),
             rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300000), rng.integers(0, 256, 200000).astype(np.uint8)]
    for x in parts:
        s = po.compress(x, chain, entropy, block_size=1 << 18, jobs=4, input_size=len(x))
        y = po.decompress(s, len(x) + 64, jobs=4)
        assert np.array_equal(y, x)


def po_dt(name):
    return ["UNDEFINED", "TEXT", "MULTIMEDIA", "EXE", "NUMERIC", "BASE64", "DNA", "BIN", "UTF8", "SMALL_ALPHABET"].index(name)


# ---------------------------------------------------------------- EXE
def _synth_x86(n, seed):
    """random bytes shaped like machine code for detectExeType (all 256 values, >= 10 % zeros, >= 1 % 0xFF) with relative CALL / JMP
    instructions (E8 / E9 + 32-bit offset whose top byte is 00 or FF) at more than 1 per 200 bytes"""
    r = np.random.default_rng(seed)
    x = r.integers(0, 256, n).astype(np.uint8)
    x[r.random(n) < 0.15] = 0
    x[r.random(n) < 0.03] = 255
    for p in np.sort(r.choice(n - 16, n // 60, replace=False)):
        x[p] = 0xE8 if r.random() < 0.7 else 0xE9
        x[p + 1: p + 5] = np.frombuffer(np.int32(int(r.integers(-200000, 200000))).tobytes(), np.uint8)
    return x


def _synth_arm64(n, seed):
    r = np.random.default_rng(seed)
    n4 = n // 4
    w = r.integers(0, 1 << 32, n4, dtype=np.uint64).astype(np.uint32)
    w[r.random(n4) < 0.25] &= 0x00FF00FF  # zero bytes
    w[r.random(n4) < 0.05] |= 0xFF00FF00
    br = r.random(n4) < 0.03
    off = r.integers(-50000, 50000, n4).astype(np.int64) & ((1 << 26) - 1)
    w[br] = (np.where(r.random(n4) < 0.5, 0x94000000, 0x14000000)[br] | off[br]).astype(np.uint32)
    x = np.frombuffer(w.tobytes(), np.uint8).copy()
    x[:4] = [0x10, 0x20, 0x30, 0x40]  # no magic number
    return np.concatenate([x, r.integers(0, 256, n - 4 * n4).astype(np.uint8)])


@pytest.mark.parametrize("n", [4096, 100003, 1 << 20])
def test_exe_roundtrip(n):
    y, dt = _roundtrip(po.T_EXE, _synth_x86(n, n))
    assert y[0] == 0x40 and dt == po_dt("EXE")
    # header: mode, code start (0 without an executable header), end of the transformed code section
    assert int.from_bytes(y[1:5].tobytes(), "little") == 0 and int.from_bytes(y[5:9].tobytes(), "little") <= len(y)
    a = _synth_arm64(n, n + 1)
    ya, dta = _fwd(po.T_EXE, a)
    if ya is not None:  # the ARM64 heuristics may decline the synthetic sample
        assert ya[0] == 0x20 and dta == po_dt("EXE")
        z = _inv(po.T_EXE, ya, n + 64)
        z = z[0] if isinstance(z, tuple) else z
        assert np.array_equal(z, a)


def test_exe_real_binary_and_skips():
    import sys

    exe = np.fromfile(sys.executable, dtype=np.uint8)[: 2 << 20]
    if len(exe) >= 65536 and bytes(exe[:4]) == b"\x7fELF":  # the interpreter itself: an ELF header tells the code section
        y, dt = _roundtrip(po.T_EXE, exe)
        assert dt == po_dt("EXE") and y[0] in (0x40, 0x20)
    rnd = np.random.default_rng(0).integers(0, 256, 1 << 16).astype(np.uint8)
    y, dt = _fwd(po.T_EXE, rnd)
    assert y is None and dt == po_dt("BIN")
    y, dt = _fwd(po.T_EXE, synth.markov_text(1 << 16, seed=1))
    assert y is None
    assert _fwd(po.T_EXE, _synth_x86(1 << 16, 3), po_dt("TEXT"))[0] is None
    assert _fwd(po.T_EXE, _synth_x86(4000, 3))[0] is None  # below _EXE_MIN_BLOCK_SIZE


def test_level4_chain_roundtrip():
    parts = [synth.markov_text(400000, seed=3), _synth_x86(300000, 5), np.random.default_rng(2).integers(0, 256, 100000).astype(np.uint8)]
    for x in parts:
        s = po.compress(x, "TEXT+UTF+EXE+PACK+MM+ROLZ", "NONE", block_size=1 << 18, jobs=4, input_size=len(x))
        assert np.array_equal(po.decompress(s, len(x) + 64, jobs=4), x)
