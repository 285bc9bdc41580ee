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


# ----------------------------------------------------------------------------------------------------------------------
# Corpus-shaped slabs. The real corpora (silesia.tar, enwik9) are not available offline; these
# generators reproduce their *mix* of content classes with seeded numpy so that every run (GPU, oracle, CPU baseline)
# sees the same bytes. Everything is built from "token streams": a pool of byte strings drawn with a Zipf law, with
# earlier phrases copied forward (long repeats) — vectorised, ~20 MB/s.
# ----------------------------------------------------------------------------------------------------------------------
_COMMON = ("the of and to a in is that it was for on with he be as his at by had this not but from have are which her she or you they "
           "an were there been one all we their has would when if so no will him who more said out up what about into than them can "
           "only other time new some could these two may first then do any like my now over such our man me even most made after also "
           "did many before must through back years where much your way well down should because each just those people how too little "
           "state good very make world still own see men work long get here between both life being under never day same another know "
           "while last might us great old year off come since against go came right used take three").split()


def _syllable_words(rng, count, lo=1, hi=4):
    cons = np.frombuffer(b"bcdfghjklmnprstvwyz", np.uint8)
    vow = np.frombuffer(b"aeiou", np.uint8)
    words = []
    for _ in range(count):
        k = int(rng.integers(lo, hi + 1))
        w = bytearray()
        for _s in range(k):
            w.append(int(cons[rng.integers(0, len(cons))]))
            w.append(int(vow[rng.integers(0, len(vow))]))
            if rng.random() < 0.35:
                w.append(int(cons[rng.integers(0, len(cons))]))
        words.append(bytes(w))
    return words


def _token_stream(rng, tokens, probs, n, repeat_every=0, repeat_len=(4, 40), chunk_tokens=1 << 20):
    """n bytes made of `tokens` drawn i.i.d. with `probs`; every ~repeat_every tokens an earlier phrase of repeat_len tokens is
    copied forward (long-distance repeats)."""
    pool = np.frombuffer(b"".join(tokens), np.uint8)
    tlen = np.array([len(t) for t in tokens], np.int64)
    tstart = np.concatenate([[0], np.cumsum(tlen)[:-1]])
    out = np.empty(n + 256, np.uint8)
    pos = 0
    cdf = np.cumsum(np.asarray(probs, np.float64))
    cdf /= cdf[-1]
    mean_len = float(np.dot(tlen, np.diff(np.concatenate([[0.0], cdf]))))
    while pos < n:
        ct = int(min(chunk_tokens, max(4096, (n - pos) / mean_len * 1.1 + 4096)))
        idx = np.searchsorted(cdf, rng.random(ct), side="right").astype(np.int64)
        np.minimum(idx, len(tokens) - 1, out=idx)
        if repeat_every:
            k = ct // repeat_every
            dsts = np.sort(rng.integers(repeat_len[1] + 1, ct - repeat_len[1], k))
            lens = rng.integers(repeat_len[0], repeat_len[1], k)
            back = (rng.pareto(1.2, k) * 50).astype(np.int64) + lens
            for d, ln, bk in zip(dsts.tolist(), lens.tolist(), back.tolist()):
                s = d - bk
                if s < 0:
                    s = int(rng.integers(0, d - ln)) if d > ln else 0
                idx[d:d + ln] = idx[s:s + ln]
        lens = tlen[idx]
        ends = np.cumsum(lens)
        total = int(ends[-1])
        src = np.repeat(tstart[idx] - (ends - lens), lens) + np.arange(total, dtype=np.int64)
        take = min(total, n + 256 - pos)
        out[pos:pos + take] = pool[src[:take]]
        pos += take
    return out[:n].copy()


def _zipf(k, s=1.0):
    p = 1.0 / np.power(np.arange(1, k + 1, dtype=np.float64), s)
    return p / p.sum()


def english_text(n, seed=SEED, crlf=True, html=False):
    """Word-structured English-like prose: common English words + a Zipf tail of syllable words, sentences with capitals
    and punctuation, CR LF (or LF) line ends, phrase repeats; optionally with HTML-ish markup lines (webster-like)."""
    rng = np.random.default_rng(seed)
    base = [w.encode() for w in _COMMON] + _syllable_words(rng, 6000)
    nl = b"\r\n" if crlf else b"\n"
    toks, weights = [], []
    pw = _zipf(len(base), 1.05)
    seps = [(b" ", 0.80), (b", ", 0.07), (b". ", 0.0), (nl, 0.09), (b"; ", 0.01), (b" - ", 0.01), (b"'s ", 0.02)]
    # a token = word + separator (the separator of ". " is followed by a capitalised word, so sentence starts are tokens of their own)
    for w, p in zip(base, pw):
        for s, q in seps:
            if q > 0:
                toks.append(w + s)
                weights.append(p * q * 0.93)
        toks.append(w + b". " )
        weights.append(p * 0.035)
        toks.append(w.capitalize() + b" ")
        weights.append(p * 0.035)
    if html:
        for t, q in ((b"<p>", 0.004), (b"</p>" + nl, 0.004), (b"<i>", 0.003), (b"</i> ", 0.003), (b"<b>", 0.002), (b"</b> ", 0.002),
                     (b"<h1>", 0.0005), (b"</h1>" + nl, 0.0005), (b"&amp; ", 0.001), (b"<hw>", 0.002), (b"</hw> ", 0.002), (b"<def>", 0.002), (b"</def>" + nl, 0.002)):
            toks.append(t)
            weights.append(q)
    for y in range(1800, 1830):
        toks.append(str(y).encode() + b" ")
        weights.append(0.0001)
    return _token_stream(rng, toks, weights, n, repeat_every=40, repeat_len=(3, 30))


def x86_like(n, seed=SEED):
    """Machine-code-like bytes: Zipf-weighted instruction templates (1..7 bytes), E8/E9 rel32 calls/jumps and 0F 8x jumps with
    small signed displacements, mov-immediates with 4-byte little-endian addresses on a stride, zero padding and an ASCII
    string table now and then (mozilla / ooffice / samba objects in silesia)."""
    rng = np.random.default_rng(seed)
    toks, weights = [], []
    common_ops = [0x8B, 0x89, 0x83, 0xFF, 0x55, 0x5D, 0xC3, 0x50, 0x51, 0x52, 0x53, 0x56, 0x57, 0x85, 0x74, 0x75, 0xEB, 0x33, 0x3B, 0x8D, 0x6A, 0x68, 0xC7, 0x0F, 0x48, 0x4C, 0x90, 0xCC]
    modrm = [0x45, 0x4D, 0x55, 0x5D, 0x75, 0x7D, 0xC0, 0xC1, 0xC8, 0xD0, 0xF6, 0xFF, 0x04, 0x24, 0x44, 0x4C, 0x08, 0x10, 0xEC, 0xE4]
    for i in range(3000):
        k = int(rng.integers(1, 8))
        t = bytearray([common_ops[int(rng.integers(0, len(common_ops)))]])
        if k > 1:
            t.append(modrm[int(rng.integers(0, len(modrm)))])
        while len(t) < k:
            t.append(int(rng.choice([0, 0, 0, 0xFF, 4, 8, 0x10, 1, int(rng.integers(0, 256))])))
        toks.append(bytes(t))
    weights += list(_zipf(3000, 1.1) * 0.80)
    for i in range(600):  # calls / jumps with rel32 displacements (mostly near: high bytes 00 00 or FF FF)
        d = int(rng.normal(0, 40000))
        op = b"\xE8" if i % 4 else b"\xE9"
        toks.append(op + int(d & 0xFFFFFFFF).to_bytes(4, "little"))
    weights += list(_zipf(600, 0.7) * 0.08)
    for i in range(200):
        d = int(rng.normal(0, 3000))
        toks.append(bytes([0x0F, 0x80 + (i & 15)]) + int(d & 0xFFFFFFFF).to_bytes(4, "little"))
    weights += list(_zipf(200, 0.7) * 0.02)
    for i in range(800):  # absolute addresses on a stride
        a = 0x00401000 + 16 * int(rng.integers(0, 1 << 14))
        toks.append(bytes([0xB8 + (i & 7)]) + a.to_bytes(4, "little"))
    weights += list(_zipf(800, 0.8) * 0.06)
    toks.append(b"\x00" * 16)
    weights.append(0.01)
    toks.append(b"\xCC" * 8)
    weights.append(0.005)
    for w in _syllable_words(rng, 300, 2, 5):
        toks.append(w + b"\x00")
    weights += list(_zipf(300, 1.0) * 0.025)
    return _token_stream(rng, toks, weights, n, repeat_every=200, repeat_len=(8, 200))


def records_like(n, seed=SEED):
    """Highly repetitive record data (nci / osdb in silesia): fixed-format numeric lines from a small template set, long verbatim
    repeats with small edits."""
    rng = np.random.default_rng(seed)
    toks, weights = [], []
    atoms = [b"C", b"N", b"O", b"H", b"S", b"Cl", b"P", b"F"]
    for i in range(400):
        x, y = rng.normal(0, 3, 2)
        a = atoms[int(min(rng.integers(0, 12), 7))]
        toks.append(b"%10.4f%10.4f%10.4f %-3s 0  0  0  0  0  0  0  0  0  0\n" % (x, y, 0.0, a))
    weights += list(_zipf(400, 0.9) * 0.6)
    for i in range(300):
        a, b = rng.integers(1, 60, 2)
        toks.append(b"%3d%3d%3d  0  0  0  0\n" % (a, b, 1 + (i & 1)))
    weights += list(_zipf(300, 0.8) * 0.3)
    toks += [b"M  END\n", b"$$$$\n", b"  -ISIS-  \n\n", b"> <NSC>\n", b"> <CAS_RN>\n"]
    weights += [0.01, 0.01, 0.01, 0.01, 0.01]
    for i in range(500):
        toks.append(b"%d\n\n" % int(rng.integers(1000, 999999)))
    weights += list(_zipf(500, 0.3) * 0.05)
    return _token_stream(rng, toks, weights, n, repeat_every=30, repeat_len=(10, 120))


def walk16(n, seed=SEED):
    """16-bit little-endian random-walk samples with 12 significant bits (mr / x-ray in silesia)."""
    rng = np.random.default_rng(seed)
    k = n // 2 + 1
    steps = rng.normal(0, 12, k) + 40 * np.sin(np.arange(k) / 700.0) * (rng.random(k) < 0.02)
    v = (np.cumsum(steps) % 8192 + 2048 * np.sin(np.arange(k) / 5000.0) + 2048).astype(np.int64) & 0x0FFF
    return v.astype("<u2").view(np.uint8)[:n].copy()


def xml_like(n, seed=SEED):
    """XML tag soup with attributes, numbers and indentation (xml in silesia)."""
    rng = np.random.default_rng(seed)
    names = [w.decode() for w in _syllable_words(rng, 60, 2, 4)]
    toks, weights = [], []
    for i, nm in enumerate(names):
        ind = "  " * (i % 5)
        toks += [("%s<%s>" % (ind, nm)).encode(), ("</%s>\n" % nm).encode(), ('%s<%s id="' % (ind, nm)).encode(), ("%s<%s/>\n" % (ind, nm)).encode()]
        w = 1.0 / (i + 1)
        weights += [w, w, 0.4 * w, 0.2 * w]
    wsum = sum(weights)
    weights = [w / wsum * 0.55 for w in weights]
    for i in range(2000):
        toks.append(("%d" % int(rng.integers(0, 100000))).encode())
    weights += list(_zipf(2000, 0.6) * 0.12)
    toks += [b'">', b'" type="', b'" name="', b"\n", b"<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n", b"<!-- ", b" -->\n", b"&lt;", b"&gt;"]
    weights += [0.04, 0.02, 0.02, 0.03, 0.0002, 0.002, 0.002, 0.003, 0.003]
    words = [w.encode() for w in _COMMON[:120]] + _syllable_words(rng, 500)
    toks += [w + b" " for w in words]
    weights += list(_zipf(len(words), 1.0) * 0.2)
    return _token_stream(rng, toks, weights, n, repeat_every=25, repeat_len=(6, 80))


# silesia.tar's content classes in tar order (dickens, mozilla, mr, nci, ooffice, osdb, reymont, samba, sao, webster, xml, x-ray)
# folded into the six classes of SURVEY.md §8d C3: 35 % text, 25 % x86/ELF-like, 15 % records, 10 % 16-bit walks, 10 % XML, 5 % random.
SILESIA_MIX = (("text", 0.10), ("x86", 0.17), ("walk16", 0.05), ("records", 0.15), ("x86", 0.08), ("random", 0.03), ("text", 0.05),
               ("text_html", 0.20), ("xml", 0.10), ("walk16", 0.05), ("random", 0.02))


def silesia_shaped(n=200_000_000, seed=SEED + 3):
    """a `silesia.tar`-shaped slab of n bytes (default 200,000,000): seeded segments in the proportions of SILESIA_MIX."""
    gens = {"text": lambda m, s: english_text(m, s, crlf=True), "text_html": lambda m, s: english_text(m, s, crlf=False, html=True),
            "x86": x86_like, "walk16": walk16, "records": records_like, "xml": xml_like, "random": uniform_bytes}
    out = np.empty(n, np.uint8)
    pos = 0
    for i, (kind, frac) in enumerate(SILESIA_MIX):
        m = n - pos if i == len(SILESIA_MIX) - 1 else min(n - pos, int(round(n * frac)))
        if m <= 0:
            continue
        out[pos:pos + m] = gens[kind](m, seed + 101 * i)
        pos += m
    return out


def enwik_shaped(n=1_000_000_000, seed=SEED + 4):
    """enwik9-shaped text: English prose inside XML page markup with wiki syntax and ~3 % UTF-8 multi-byte runs."""
    rng = np.random.default_rng(seed)
    base = [w.encode() for w in _COMMON] + _syllable_words(rng, 8000)
    toks, weights = [], []
    pw = _zipf(len(base), 1.05)
    for w, p in zip(base, pw):
        toks += [w + b" ", w + b", ", w + b". ", w.capitalize() + b" ", b"[[" + w + b"]] ", w + b"\n"]
        weights += [p * 0.80, p * 0.05, p * 0.04, p * 0.04, p * 0.02, p * 0.02]
    markup = [(b"  <page>\n    <title>", 0.0006), (b"</title>\n    <id>", 0.0006), (b"</id>\n    <revision>\n      <timestamp>", 0.0006),
              (b"</timestamp>\n      <text xml:space=\"preserve\">", 0.0006), (b"</text>\n    </revision>\n  </page>\n", 0.0006), (b"== ", 0.002), (b" ==\n", 0.002),
              (b"'''", 0.003), (b"''", 0.003), (b"{{", 0.002), (b"}}\n", 0.002), (b"&quot;", 0.004), (b"&lt;", 0.002), (b"&gt;", 0.002), (b"* ", 0.004),
              (b"|", 0.004), (b"[http://www.", 0.001), (b".com/ ", 0.001), (b"\n\n", 0.006)]
    for t, q in markup:
        toks.append(t)
        weights.append(q)
    for i in range(1500):
        toks.append(("%d" % int(rng.integers(0, 3000))).encode() + b" ")
    weights += list(_zipf(1500, 0.5) * 0.012)
    # UTF-8 multi-byte words (Latin-1 supplement, Greek, Cyrillic, CJK): ~3 % of the bytes
    utf = []
    for lo, hi, cnt in ((0xC0, 0x17F, 300), (0x391, 0x3C9, 150), (0x410, 0x44F, 300), (0x4E00, 0x4F00, 250)):
        for _ in range(cnt):
            k = int(rng.integers(2, 7))
            utf.append("".join(chr(int(c)) for c in rng.integers(lo, hi, k)).encode("utf-8") + b" ")
    toks += utf
    weights += list(_zipf(len(utf), 0.8) * 0.0085)
    return _token_stream(rng, toks, weights, n, repeat_every=60, repeat_len=(3, 25))
