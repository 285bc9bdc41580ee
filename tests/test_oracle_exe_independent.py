"""An independent pin for the EXE codec (kz_exe_core.cuh and oracle/kzo_exe.cpp are one transliteration in two places, so GPU ==
oracle proves nothing about fidelity to EXECodec.go). This file holds a THIRD statement of the x86 path — written in Python straight from
v2/transform/EXECodec.go (detectExeType :710-810 for blocks without an executable header, forwardX86 :174-300, inverseX86 :333-449), sharing no
text with either C++ version — and checks the oracle against it byte for byte on blocks built to pass the detector. A misreading shared by the
two C++ files shows up here unless this third reading repeats it."""
import numpy as np
import pytest

from oracle import pyoracle as po

ESC, PREFIX, MASK_ADDR, ADDR_MASK = 0x9B, 0x0F, 0xF0F0F0F0, (1 << 24) - 1


def detect(src):
    """detectExeType for a block whose first bytes are no known magic: the ad-hoc thresholds (:734-810). src = block[:count-4], code = [0, count-8).
    Returns "x86", "arm64" or None."""
    end = len(src) - 4
    histo = [0] * 256
    jumps_x86 = jumps_arm = 0
    i = 0
    while i < end:
        histo[src[i]] += 1
        counted = False
        if (src[i] & 0xFE) == 0xE8:
            if src[i + 4] in (0, 0xFF):
                jumps_x86 += 1
                counted = True
        elif src[i] == PREFIX:
            i += 1
            if src[i] in (0x38, 0x3A):
                i += 1
            if (src[i] & 0xF0) == 0x80:
                jumps_x86 += 1
                counted = True
        if not counted and (i & 3) == 0:
            instr = int.from_bytes(bytes(src[i: i + 4]), "little")
            if (instr & 0xFC000000) in (0x14000000, 0x94000000) or (instr & 0x7F000000) in (0x34000000, 0x3500000):
                jumps_arm += 1
        i += 1
    if sum(1 for h in histo if h) != 256:  # DetectSimpleType: anything but "all 256 values present" is not DT_BIN here
        return None
    if histo[0] < end // 10 or sum(histo[:16]) > end // 2 or histo[255] < end // 100:
        return None
    if jumps_x86 >= end // 200:
        return "x86"
    if jumps_arm >= end // 200:
        return "arm64"
    return None


def detect_is_x86(src):
    return detect(src) == "x86"


def forward_arm(src):
    """forwardARM (:507-613) with codeStart = 0, codeEnd = count - 8."""
    count = len(src)
    code_end = count - 8
    dst = bytearray(count + count // 32 + 64)
    dst_end = len(dst) - 8
    dst[0] = 0x20
    s, d, matches = 0, 9, 0
    while s + 4 <= code_end and d < dst_end:
        instr = int.from_bytes(bytes(src[s: s + 4]), "little")
        opcode1 = instr & 0xFC000000
        if opcode1 not in (0x14000000, 0x94000000):
            dst[d: d + 4] = bytes(src[s: s + 4])
            s += 4
            d += 4
            continue
        offset = instr & 0x03FFFFFF
        if instr & (1 << 25) == 0:
            addr = s + 4 * offset
        else:
            addr = s - 4 * ((-offset) & 0x03FFFFFF)
        if addr < 0:
            addr = 0
        val = (opcode1 | (addr >> 2)) & 0xFFFFFFFF
        dst[d: d + 4] = val.to_bytes(4, "little")
        if addr == 0:  # 0 address as escape: the original instruction follows
            dst[d + 4: d + 8] = bytes(src[s: s + 4])
            s += 4
            d += 8
            continue
        s += 4
        d += 4
        matches += 1
    if matches < 16:
        return None
    if (s + 4 <= code_end and d >= dst_end) or d + (count - s) > dst_end:
        return None
    dst[1:5] = (0).to_bytes(4, "little")
    dst[5:9] = d.to_bytes(4, "little")
    dst[d: d + count - s] = bytes(src[s:count])
    d += count - s
    if d > count + count // 50:
        return None
    return np.frombuffer(bytes(dst[:d]), np.uint8)


def inverse_arm(enc, cap):
    code_end = int.from_bytes(bytes(enc[5:9]), "little")
    out = bytearray()
    s = 9
    while s < code_end:
        instr = int.from_bytes(bytes(enc[s: s + 4]), "little")
        opcode1 = instr & 0xFC000000
        if opcode1 not in (0x14000000, 0x94000000):
            out += bytes(enc[s: s + 4])
            s += 4
            continue
        addr = (instr & 0x03FFFFFF) << 2
        if addr == 0:
            out += bytes(enc[s + 4: s + 8])
            s += 8
            continue
        offset = (addr - len(out)) >> 2
        out += ((opcode1 | (offset & 0x03FFFFFF)) & 0xFFFFFFFF).to_bytes(4, "little")
        s += 4
    out += bytes(enc[s:])
    assert len(out) <= cap
    return np.frombuffer(bytes(out), np.uint8)


def make_arm_block(n, seed):
    """AArch64-like words without an executable header: B / BL with small offsets, every byte value present, enough zeros and 0xFF bytes."""
    r = np.random.default_rng(seed)
    x = r.integers(0, 256, n).astype(np.uint8)
    x[r.random(n) < 0.15] = 0
    x[r.random(n) < 0.03] = 0xFF
    x[:256] = np.arange(256, dtype=np.uint8)
    x[0:4] = [0xFD, 0x7B, 0xBF, 0xA9]  # stp x29, x30, [sp, #-16]!: no magic number
    for pos in range(0, n - 3, 4):
        if pos >= 512 and pos < n - 32 and r.random() < 0.08:
            off = int(r.integers(-(pos >> 2) + 1, 1 << 18)) & 0x03FFFFFF  # targets inside the block or ahead of it (address 0 is the escape)
            instr = (0x94000000 if r.random() < 0.7 else 0x14000000) | off
            x[pos: pos + 4] = np.frombuffer(instr.to_bytes(4, "little"), np.uint8)
        elif pos >= 256 and (int(x[pos + 3]) & 0xFC) in (0x14, 0x94) and r.random() < 0.9:
            x[pos + 3] ^= 0x40  # most accidental B / BL patterns of the filler go away (the rest exercises the escapes)
    return x


def forward_x86(src):
    """forwardX86 with codeStart = 0, codeEnd = count - 8. Returns the encoded block or None (the transform declines)."""
    count = len(src)
    code_end = count - 8
    dst = bytearray(count + count // 32 + 64)  # MaxEncodedLen is generous; only dstEnd = len(dst) - 5 matters and it is never reached here
    dst_end = len(dst) - 5
    dst[0] = 0x40
    s, d, matches, boundary = 0, 9, 0, False
    while s < code_end and d < dst_end:
        if src[s] == PREFIX:
            if s + 1 >= code_end:
                boundary = True
                break
            if (src[s + 1] & 0xF0) == 0x80 and s + 5 >= code_end:
                boundary = True
                break
            dst[d] = src[s]
            s += 1
            d += 1
            if (src[s] & 0xF0) != 0x80:
                if src[s] == ESC:
                    dst[d] = ESC
                    d += 1
                dst[d] = src[s]
                s += 1
                d += 1
                continue
            if s + 4 >= code_end:
                boundary = True
                break
        elif (src[s] & 0xFE) != 0xE8:
            if src[s] == ESC:
                dst[d] = ESC
                d += 1
            dst[d] = src[s]
            s += 1
            d += 1
            continue
        elif s + 4 >= code_end:
            boundary = True
            break
        sgn = src[s + 4]
        offset = int.from_bytes(bytes(src[s + 1: s + 5]), "little")
        if (sgn != 0 and sgn != 0xFF) or offset == 0xFF000000:
            dst[d] = ESC
            dst[d + 1] = src[s]
            s += 1
            d += 2
            continue
        addr = s + offset if sgn == 0 else s - ((-offset) & ADDR_MASK)
        dst[d] = src[s]
        dst[d + 1: d + 5] = ((addr ^ MASK_ADDR) & 0xFFFFFFFF).to_bytes(4, "big")
        s += 5
        d += 5
        matches += 1
    if matches < 16:
        return None
    if s < code_end and not boundary:
        return None
    if d + (count - s) > dst_end:
        return None
    dst[1:5] = (0).to_bytes(4, "little")
    dst[5:9] = d.to_bytes(4, "little")
    dst[d: d + count - s] = bytes(src[s:count])
    d += count - s
    if d > count + count // 50:
        return None
    return np.frombuffer(bytes(dst[:d]), np.uint8)


def inverse_x86(enc, cap):
    code_start = int.from_bytes(bytes(enc[1:5]), "little")
    code_end = int.from_bytes(bytes(enc[5:9]), "little")
    assert code_start == 0 and 9 <= code_end <= len(enc)
    out = bytearray()
    s = 9
    while s < code_end:
        if enc[s] == PREFIX:
            if s + 1 >= code_end:
                out.append(enc[s])
                s += 1
                break
            out.append(enc[s])
            s += 1
            if (enc[s] & 0xF0) != 0x80:
                if enc[s] == ESC:
                    s += 1
                out.append(enc[s])
                s += 1
                continue
        elif (enc[s] & 0xFE) != 0xE8:
            if enc[s] == ESC:
                s += 1
            out.append(enc[s])
            s += 1
            continue
        addr = int.from_bytes(bytes(enc[s + 1: s + 5]), "big") ^ MASK_ADDR
        offset = addr - len(out)
        out.append(enc[s])
        if offset >= 0:
            out += (offset & 0xFFFFFFFF).to_bytes(4, "little")
        else:
            out += ((-((-offset) & ADDR_MASK)) & 0xFFFFFFFF).to_bytes(4, "little")
        s += 5
    out += bytes(enc[s:])
    assert len(out) <= cap
    return np.frombuffer(bytes(out), np.uint8)


def make_block(n, seed, escapes=True):
    """Machine-code-like bytes without an executable header: every byte value present, > 10 % zeros, > 1 % 0xFF, calls / jumps with small offsets."""
    r = np.random.default_rng(seed)
    x = r.integers(0, 256, n).astype(np.uint8)
    x[r.random(n) < 0.15] = 0
    x[r.random(n) < 0.03] = 0xFF
    x[:256] = np.arange(256, dtype=np.uint8)
    x[0:4] = [0x55, 0x48, 0x89, 0xE5]  # no magic number
    pos = 300
    while pos + 8 < n - 16:
        kind = r.integers(0, 10)
        if kind < 6:  # call / jmp rel32 with a small forward or backward offset
            x[pos] = 0xE8 if kind < 4 else 0xE9
            off = int(r.integers(-(1 << 20), 1 << 20))
            x[pos + 1: pos + 5] = np.frombuffer((off & 0xFFFFFFFF).to_bytes(4, "little"), np.uint8)
            pos += 5
        elif kind < 8:  # jcc rel32
            x[pos], x[pos + 1] = 0x0F, 0x80 + int(r.integers(0, 16))
            off = int(r.integers(-(1 << 16), 1 << 16))
            x[pos + 2: pos + 6] = np.frombuffer((off & 0xFFFFFFFF).to_bytes(4, "little"), np.uint8)
            pos += 6
        elif escapes and kind == 8:
            x[pos] = ESC  # a literal escape byte, and a 0F followed by one
            x[pos + 1], x[pos + 2] = PREFIX, ESC
            pos += 3
        pos += int(r.integers(4, 40))
    return x


@pytest.mark.parametrize("n,seed", [(4096, 1), (5000, 2), (65536, 3), (300001, 4), (1 << 20, 5)])
def test_oracle_exe_x86_matches_the_independent_statement(n, seed):
    x = make_block(n, seed)
    assert detect_is_x86(x[: n - 4]), "the generator must produce blocks the detector takes for x86 code"
    want = forward_x86(x)
    got, dt = po.transform_forward(po.T_EXE, x)
    assert (want is None) == (got is None)
    assert want is not None, "the block should be transformed"
    assert len(got) == len(want) and np.array_equal(got, want), int(np.argmax(got[: min(len(got), len(want))] != want[: min(len(got), len(want))]))
    back = po.transform_inverse(po.T_EXE, want, n + 64)
    assert np.array_equal(back, x)
    assert np.array_equal(inverse_x86(want, n + 64), x)  # and the third statement is self-consistent


def test_oracle_exe_declines_like_the_independent_statement():
    r = np.random.default_rng(9)
    x = r.integers(0, 256, 8192).astype(np.uint8)  # uniform bytes: too few zeros -> not an executable
    assert not detect_is_x86(x[:-4])
    got, _ = po.transform_forward(po.T_EXE, x)
    assert got is None
    y = make_block(8192, 11)
    y[300:] = np.where(np.isin(y[300:], [0xE8, 0xE9, 0x0F]), 0x90, y[300:])  # no jumps left: below the threshold
    assert not detect_is_x86(y[:-4])
    got, _ = po.transform_forward(po.T_EXE, y)
    # random words hold enough ARM64 B / BL opcode patterns (2 of 64 values of the top six bits) to pass the ARM threshold: the block is
    # either declined or coded as ARM64 (mode byte 0x20, not restated here) — never as x86
    assert got is None or got[0] == 0x20


@pytest.mark.parametrize("n,seed", [(4096, 21), (65536, 22), (300004, 23)])
def test_oracle_exe_arm64_matches_the_independent_statement(n, seed):
    x = make_arm_block(n, seed)
    assert detect(x[: n - 4]) == "arm64", "the generator must produce blocks the detector takes for AArch64 code"
    want = forward_arm(x)
    got, dt = po.transform_forward(po.T_EXE, x)
    assert (want is None) == (got is None)
    assert want is not None, "the block should be transformed"
    assert len(got) == len(want) and np.array_equal(got, want), int(np.argmax(got[: min(len(got), len(want))] != want[: min(len(got), len(want))]))
    back = po.transform_inverse(po.T_EXE, want, n + 64)
    assert np.array_equal(back, x)
    assert np.array_equal(inverse_arm(want, n + 64), x)
