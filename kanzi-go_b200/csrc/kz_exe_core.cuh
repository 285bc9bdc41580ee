// EXE transform (EXECodec): relative call / jump targets of x86 and ARM64 code become absolute addresses. The per-block logic, written
// once for host and device (kernels in kz_exe.cu, CPU check in tests/test_exe_core_host.py). Reference: v2/transform/EXECodec.go —
// Forward :109-172, forwardX86 :174-297, Inverse :302-331, inverseX86 :333-449, forwardARM :507-613, inverseARM :615-698,
// MaxEncodedLen :701-708, detectExeType :710-810, parseExeHeader :812-1021 (the pre-v3 bitstream layout :451-505 is not handled).
// A slice access outside its bounds panics in Go and the block task turns that into a block error (io/CompressedStream.go:735-743):
// here such an access raises a flag and the entry points report it as a hard error rather than a skip.
#pragma once
#include <stdint.h>

#ifndef KZ_EXE_HD
#ifdef __CUDACC__
#define KZ_EXE_HD __host__ __device__ __forceinline__
#else
#define KZ_EXE_HD inline
#endif
#endif

namespace kz {
namespace exec {

KZ_EXE_HD int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
KZ_EXE_HD int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
KZ_EXE_HD void copy_bytes(uint8_t* d, const uint8_t* s, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) d[i] = s[i];
}
// internal/Magic.go:73-112 GetMagicType (0 = no magic)
KZ_EXE_HD uint32_t magic_key(const uint8_t* p, int64_t n) {
    if (n < 4) return 0;
    const uint32_t key = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    if ((key & ~0x0Fu) == 0xFFD8FFE0u) return key;
    if ((key >> 8) == 0x425A68u || (key >> 8) == 0x494433u) return key >> 8;
    const uint32_t k32[18] = {0x47494638u, 0x25504446u, 0x504B0304u, 0x377ABCAFu, 0x89504E47u, 0x7F454C46u, 0xFEEDFACEu, 0xCEFAEDFEu, 0xFEEDFACFu,
                              0xCFFAEDFEu, 0x28B52FFDu, 0x81CFB2CEu, 0x4D534346u, 0x52494646u, 0x664C6143u, 0xFD377A58u, 0x4B414E5Au, 0x52617221u};
    for (int i = 0; i < 18; i++)
        if (key == k32[i]) return key;
    const uint32_t k16 = key >> 16;
    if (k16 == 0x1F8Bu || k16 == 0x424Du || k16 == 0x4D5Au) return k16;
    if (k16 == 0x5034u || k16 == 0x5035u || k16 == 0x5036u) {
        const uint32_t sub = (key >> 8) & 0xFF;
        if (sub == 0x07 || sub == 0x0A || sub == 0x0D || sub == 0x20) return k16;
    }
    return 0;
}
// internal/Global.go:346-419 DetectSimpleType
KZ_EXE_HD int detect_simple_type_i(uint32_t count, const int* f) {
    if (count == 0) return 0;
    const char* DNA = "acgntuACGNTU";
    const char* NUM = "0123456789+-*/=,.:; ";
    const char* B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    int64_t sum = 0;
    for (int i = 0; i < 12; i++) sum += f[(uint8_t)DNA[i]];
    if (sum > (int64_t)count - (int64_t)count / 12) return 6;
    sum = 0;
    for (int i = 0; i < 20; i++) sum += f[(uint8_t)NUM[i]];
    if (sum == (int64_t)count) return 4;
    sum = 0;
    for (int i = 0; i < 64; i++) sum += f[(uint8_t)B64[i]];
    if (sum + f[0x3D] == (int64_t)count) return 5;
    sum = 0;
    for (int i = 0; i < 256; i++)
        if (f[i] > 0) sum++;
    if (sum == 256) return 7;
    if (sum <= 4) return 9;
    return 0;
}

static const uint8_t X86_MASK_JUMP = 0xFE, X86_INSTRUCTION_JUMP = 0xE8, X86_INSTRUCTION_JCC = 0x80, X86_TWO_BYTE_PREFIX = 0x0F, X86_MASK_JCC = 0xF0,
              X86_ESCAPE = 0x9B;
static const uint8_t NOT_EXE = 0x80, EXE_X86 = 0x40, EXE_ARM64 = 0x20, MASK_DT = 0x0F;
static const int64_t X86_ADDR_MASK = (1 << 24) - 1;
static const uint32_t MASK_ADDRESS = 0xF0F0F0F0u;
static const int64_t ARM_B_ADDR_MASK = (1 << 26) - 1;
static const int64_t ARM_B_OPCODE_MASK = 0xFFFFFFFFll ^ ARM_B_ADDR_MASK;
static const int64_t ARM_B_ADDR_SGN_MASK = 1 << 25;
static const int64_t ARM_OPCODE_B = 0x14000000, ARM_OPCODE_BL = 0x94000000;
static const int64_t ARM_CB_OPCODE_MASK = 0x7F000000, ARM_OPCODE_CBZ = 0x34000000, ARM_OPCODE_CBNZ = 0x3500000;
static const int64_t WIN_PE = 0x00004550, WIN_X86_ARCH = 0x014C, WIN_AMD64_ARCH = 0x8664, WIN_ARM64_ARCH = 0xAA64;
static const int64_t ELF_X86_ARCH = 0x03, ELF_AMD64_ARCH = 0x3E, ELF_ARM64_ARCH = 0xB7;
static const int64_t MAC_AMD64_ARCH = 0x01000007, MAC_ARM64_ARCH = 0x0100000C, MAC_MH_EXECUTE = 0x02, MAC_LC_SEGMENT = 0x01, MAC_LC_SEGMENT64 = 0x19;
static const int64_t MIN_BLOCK_SIZE = 4096, MAX_BLOCK_SIZE = (1 << (26 + 2)) - 1;

// a Go slice: indexing outside [0, len) panics; here the access returns 0 and raises *panic, which the callers turn into a block error
struct Slice {
    const uint8_t* p;
    int64_t len;
    bool* panic;
    KZ_EXE_HD uint8_t at(int64_t i) const {
        if (i < 0 || i >= len) {
            *panic = true;
            return 0;
        }
        return p[i];
    }
    KZ_EXE_HD bool need(int64_t i, int64_t n) const {  // src[i:] followed by an n-byte read
        if (i < 0 || i > len || i + n > len) {
            *panic = true;
            return false;
        }
        return true;
    }
    KZ_EXE_HD uint32_t le32(int64_t i) const {
        if (!need(i, 4)) return 0;
        return uint32_t(p[i]) | (uint32_t(p[i + 1]) << 8) | (uint32_t(p[i + 2]) << 16) | (uint32_t(p[i + 3]) << 24);
    }
    KZ_EXE_HD uint32_t be32(int64_t i) const {
        if (!need(i, 4)) return 0;
        return (uint32_t(p[i]) << 24) | (uint32_t(p[i + 1]) << 16) | (uint32_t(p[i + 2]) << 8) | uint32_t(p[i + 3]);
    }
    KZ_EXE_HD uint32_t le16(int64_t i) const {
        if (!need(i, 2)) return 0;
        return uint32_t(p[i]) | (uint32_t(p[i + 1]) << 8);
    }
    KZ_EXE_HD uint32_t be16(int64_t i) const {
        if (!need(i, 2)) return 0;
        return (uint32_t(p[i]) << 8) | uint32_t(p[i + 1]);
    }
    KZ_EXE_HD uint64_t le64(int64_t i) const {
        if (!need(i, 8)) return 0;
        return uint64_t(le32(i)) | (uint64_t(le32(i + 4)) << 32);
    }
    KZ_EXE_HD uint64_t be64(int64_t i) const {
        if (!need(i, 8)) return 0;
        return (uint64_t(be32(i)) << 32) | uint64_t(be32(i + 4));
    }
};

static const uint32_t WIN_MAGIC = 0x4D5A, ELF_MAGIC = 0x7F454C46u, MAC_MAGIC32 = 0xFEEDFACEu, MAC_CIGAM32 = 0xCEFAEDFEu, MAC_MAGIC64 = 0xFEEDFACFu,
               MAC_CIGAM64 = 0xCFFAEDFEu;

// :812-1021
KZ_EXE_HD bool parse_exe_header(const Slice& src, uint32_t magic, int64_t* arch, int64_t* code_start, int64_t* code_end) {
    const int64_t count = src.len;
    if (magic == WIN_MAGIC) {
        if (count >= 64) {
            const int64_t pos_pe = int64_t(src.le32(60));
            if (pos_pe > 0 && pos_pe <= count - 48 && int64_t(src.le32(pos_pe)) == WIN_PE) {
                *code_start = imin(int64_t(src.le32(pos_pe + 44)), count);
                *code_end = imin(*code_start + int64_t(src.le32(pos_pe + 28)), count);
                *arch = int64_t(src.le16(pos_pe + 4));
            }
            return true;
        }
    } else if (magic == ELF_MAGIC) {
        const bool little = src.at(5) == 1;
        if (count >= 64) {
            *code_start = 0;
            const bool is64 = src.at(4) == 2;
            auto r16 = [&](int64_t i) { return int64_t(little ? src.le16(i) : src.be16(i)); };
            auto r32 = [&](int64_t i) { return int64_t(little ? src.le32(i) : src.be32(i)); };
            auto r64 = [&](int64_t i) { return int64_t(little ? src.le64(i) : src.be64(i)); };  // int(uint64): wraps like Go
            const int64_t nb = is64 ? r16(0x3C) : r16(0x30);
            const int64_t sz = is64 ? r16(0x3A) : r16(0x2E);
            const int64_t pos = is64 ? r64(0x28) : r32(0x20);
            for (int64_t i = 0; i < nb; i++) {
                if (*src.panic) return false;
                const int64_t start = pos + i * sz;
                if (start + (is64 ? 0x28 : 0x18) >= count) return false;
                const int64_t type = r32(start + 4);
                const int64_t off = is64 ? r64(start + 0x18) : r32(start + 0x10);
                const int64_t len = is64 ? r64(start + 0x20) : r32(start + 0x14);
                if (type == 1 && len >= 64) {
                    if (*code_start == 0) *code_start = off;
                    *code_end = off + len;
                }
            }
            *arch = r16(18);
            *code_start = imin(*code_start, count);
            *code_end = imin(*code_end, count);
            return true;
        }
    } else if (magic == MAC_MAGIC32 || magic == MAC_CIGAM32 || magic == MAC_MAGIC64 || magic == MAC_CIGAM64) {
        const bool is64 = magic == MAC_MAGIC64 || magic == MAC_CIGAM64;
        *code_start = 0;
        if (count >= 64) {
            if (int64_t(src.le32(12)) != MAC_MH_EXECUTE) return false;
            *arch = int64_t(src.le32(4));
            const int64_t nb_cmds = int64_t(src.le32(0x10));
            int64_t cmd = 0, pos = is64 ? 0x20 : 0x1C;
            while (cmd < nb_cmds) {
                if (*src.panic) return false;
                const int64_t ld_cmd = int64_t(src.le32(pos));
                const int64_t sz_cmd = int64_t(src.le32(pos + 4));
                const int64_t sz_seg_hdr = is64 ? 0x48 : 0x38;
                if (ld_cmd == MAC_LC_SEGMENT || ld_cmd == MAC_LC_SEGMENT64) {
                    if (pos + 14 >= count) return false;
                    const uint64_t name_segment = src.be64(pos + 8) >> 16;
                    if (name_segment == 0x5F5F54455854ull) {  // "__TEXT"
                        const int64_t pos_section = pos + sz_seg_hdr;
                        if (pos_section + 0x34 >= count) return false;
                        const uint64_t name_section = src.be64(pos_section) >> 16;
                        if (name_section == 0x5F5F74657874ull) {  // "__text"
                            if (is64) {
                                *code_start = int64_t(int32_t(uint32_t(src.le64(pos_section + 0x30))));
                                *code_end = *code_start + int64_t(int32_t(src.le32(pos_section + 0x28)));
                            } else {
                                *code_start = int64_t(int32_t(src.le32(pos_section + 0x2C)));
                                *code_end = *code_start + int64_t(int32_t(src.le32(pos_section + 0x28)));
                            }
                            break;
                        }
                    }
                }
                cmd++;
                pos += sz_cmd;
            }
            *code_start = imin(*code_start, count);
            *code_end = imin(*code_end, count);
            return true;
        }
    }
    return false;
}

// :710-810
KZ_EXE_HD uint8_t detect_exe_type(const Slice& src, int64_t* code_start, int64_t* code_end) {
    const uint32_t magic = magic_key(src.p, src.len);
    int64_t arch = 0;
    if (parse_exe_header(src, magic, &arch, code_start, code_end)) {
        if (arch == ELF_X86_ARCH || arch == ELF_AMD64_ARCH) return EXE_X86;
        if (arch == WIN_X86_ARCH || arch == WIN_AMD64_ARCH) return EXE_X86;
        if (arch == MAC_AMD64_ARCH) return EXE_X86;
        if (arch == ELF_ARM64_ARCH || arch == WIN_ARM64_ARCH) return EXE_ARM64;
        if (arch == MAC_ARM64_ARCH) return EXE_ARM64;
    }
    int64_t jumps_x86 = 0, jumps_arm64 = 0;
    const int64_t count = *code_end - *code_start;
    int histo[256] = {0};
    for (int64_t i = *code_start; i < *code_end; i++) {
        if (*src.panic) return NOT_EXE;
        histo[src.at(i)]++;
        if ((src.at(i) & X86_MASK_JUMP) == X86_INSTRUCTION_JUMP) {
            if (src.at(i + 4) == 0 || src.at(i + 4) == 0xFF) {
                jumps_x86++;
                continue;
            }
        } else if (src.at(i) == X86_TWO_BYTE_PREFIX) {
            i++;
            if (src.at(i) == 0x38 || src.at(i) == 0x3A) i++;
            if ((src.at(i) & X86_MASK_JCC) == X86_INSTRUCTION_JCC) {
                jumps_x86++;
                continue;
            }
        }
        if ((i & 3) != 0) continue;
        const int64_t instr = int64_t(src.le32(i));
        const int64_t opcode1 = instr & ARM_B_OPCODE_MASK, opcode2 = instr & ARM_CB_OPCODE_MASK;
        if (opcode1 == ARM_OPCODE_B || opcode1 == ARM_OPCODE_BL || opcode2 == ARM_OPCODE_CBZ || opcode2 == ARM_OPCODE_CBNZ) jumps_arm64++;
    }
    const int dt = detect_simple_type_i(uint32_t(imax(count, 0)), histo);
    if (dt != 7 /*DT_BIN*/) return uint8_t(NOT_EXE | dt);
    int64_t small_vals = 0;
    for (int i = 0; i < 16; i++) small_vals += histo[i];
    if (histo[0] < count / 10 || small_vals > count / 2 || histo[255] < count / 100) return uint8_t(NOT_EXE | dt);
    if (jumps_x86 >= count / 200) return EXE_X86;
    if (jumps_arm64 >= count / 200) return EXE_ARM64;
    return uint8_t(NOT_EXE | dt);
}

KZ_EXE_HD void put_le32(uint8_t* p, uint32_t v) {
    p[0] = uint8_t(v);
    p[1] = uint8_t(v >> 8);
    p[2] = uint8_t(v >> 16);
    p[3] = uint8_t(v >> 24);
}
KZ_EXE_HD void put_be32(uint8_t* p, uint32_t v) {
    p[0] = uint8_t(v >> 24);
    p[1] = uint8_t(v >> 16);
    p[2] = uint8_t(v >> 8);
    p[3] = uint8_t(v);
}

// :174-297
KZ_EXE_HD bool forward_x86(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t code_start, int64_t code_end, int64_t* out_n, bool* panic) {
    int64_t src_idx = code_start, dst_idx = 9, matches = 0;
    const int64_t dst_end = dst_len - 5;
    dst[0] = EXE_X86;
    bool boundary = false;
    if (code_start < 0 || code_end < code_start || code_end > count) return false;
    if (code_start > 0) {
        copy_bytes(dst + dst_idx, src, (uint64_t)(code_start));
        dst_idx += code_start;
    }
    const Slice S{src, count, panic};
    while (src_idx < code_end && dst_idx < dst_end) {
        if (*panic) return false;
        if (S.at(src_idx) == X86_TWO_BYTE_PREFIX) {
            if (src_idx + 1 >= code_end) {
                boundary = true;
                break;
            }
            if ((S.at(src_idx + 1) & X86_MASK_JCC) == X86_INSTRUCTION_JCC) {
                if (src_idx + 5 >= code_end) {
                    boundary = true;
                    break;
                }
            }
            dst[dst_idx++] = src[src_idx++];
            if ((S.at(src_idx) & X86_MASK_JCC) != X86_INSTRUCTION_JCC) {
                if (src[src_idx] == X86_ESCAPE) dst[dst_idx++] = X86_ESCAPE;
                dst[dst_idx++] = src[src_idx++];
                continue;
            }
            if (src_idx + 4 >= code_end) {
                boundary = true;
                break;
            }
        } else if ((S.at(src_idx) & X86_MASK_JUMP) != X86_INSTRUCTION_JUMP) {
            if (src[src_idx] == X86_ESCAPE) dst[dst_idx++] = X86_ESCAPE;
            dst[dst_idx++] = src[src_idx++];
            continue;
        } else if (src_idx + 4 >= code_end) {
            boundary = true;
            break;
        }
        const uint8_t sgn = S.at(src_idx + 4);
        const int64_t offset = int64_t(S.le32(src_idx + 1));
        if ((sgn != 0 && sgn != 0xFF) || offset == 0xFF000000ll) {
            dst[dst_idx] = X86_ESCAPE;
            dst[dst_idx + 1] = src[src_idx];
            src_idx++;
            dst_idx += 2;
            continue;
        }
        int64_t addr = src_idx;
        if (sgn == 0) addr += offset;
        else addr -= (-offset & X86_ADDR_MASK);
        dst[dst_idx] = src[src_idx];
        put_be32(dst + dst_idx + 1, uint32_t(addr) ^ MASK_ADDRESS);
        src_idx += 5;
        dst_idx += 5;
        matches++;
    }
    if (matches < 16) return false;
    if (src_idx < code_end && !boundary) return false;
    if (dst_idx + (count - src_idx) > dst_end) return false;
    put_le32(dst + 1, uint32_t(code_start));
    put_le32(dst + 5, uint32_t(dst_idx));
    copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(count - src_idx));
    dst_idx += count - src_idx;
    if (dst_idx > count + count / 50) return false;
    *out_n = dst_idx;
    return true;
}

// :507-613 (the conditional-branch form is disabled in the reference: isCB is always false)
KZ_EXE_HD bool forward_arm(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t code_start, int64_t code_end, int64_t* out_n, bool* panic) {
    int64_t src_idx = code_start, dst_idx = 9, matches = 0;
    const int64_t dst_end = dst_len - 8;
    dst[0] = EXE_ARM64;
    if (code_start < 0 || code_end < code_start || code_end > count) return false;
    if (code_start > 0) {
        copy_bytes(dst + dst_idx, src, (uint64_t)(code_start));
        dst_idx += code_start;
    }
    const Slice S{src, count, panic};
    while (src_idx + 4 <= code_end && dst_idx < dst_end) {
        if (*panic) return false;
        const int64_t instr = int64_t(S.le32(src_idx));
        const int64_t opcode1 = instr & ARM_B_OPCODE_MASK;
        const bool is_bl = opcode1 == ARM_OPCODE_B || opcode1 == ARM_OPCODE_BL;
        if (!is_bl) {
            copy_bytes(dst + dst_idx, src + src_idx, 4);
            src_idx += 4;
            dst_idx += 4;
            continue;
        }
        const int64_t offset = int64_t(int32_t(instr & ARM_B_ADDR_MASK));
        int64_t addr;
        if ((instr & ARM_B_ADDR_SGN_MASK) == 0) addr = src_idx + 4 * offset;
        else addr = src_idx - 4 * int64_t(int32_t(-offset & ARM_B_ADDR_MASK));
        if (addr < 0) addr = 0;
        const int64_t val = opcode1 | (addr >> 2);
        if (addr == 0) {
            put_le32(dst + dst_idx, uint32_t(val));
            copy_bytes(dst + dst_idx + 4, src + src_idx, 4);
            src_idx += 4;
            dst_idx += 8;
            continue;
        }
        put_le32(dst + dst_idx, uint32_t(val));
        src_idx += 4;
        dst_idx += 4;
        matches++;
    }
    if (matches < 16) return false;
    if ((src_idx + 4 <= code_end && dst_idx >= dst_end) || dst_idx + (count - src_idx) > dst_end) return false;
    put_le32(dst + 1, uint32_t(code_start));
    put_le32(dst + 5, uint32_t(dst_idx));
    copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(count - src_idx));
    dst_idx += count - src_idx;
    if (dst_idx > count + count / 50) return false;
    *out_n = dst_idx;
    return true;
}

// :333-449
KZ_EXE_HD bool inverse_x86(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t* out_n, bool* panic) {
    const Slice S{src, count, panic};
    int64_t src_idx = 9, dst_idx = 0;
    const int64_t code_start = int64_t(S.le32(1)), code_end = int64_t(S.le32(5));
    if (code_start < 0 || code_end < src_idx || code_end > count || code_start > code_end - src_idx || code_start > dst_len - dst_idx) return false;
    if (code_start > 0) {
        copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(code_start));
        dst_idx += code_start;
        src_idx += code_start;
    }
    while (src_idx < code_end) {
        if (src[src_idx] == X86_TWO_BYTE_PREFIX) {
            if (src_idx + 1 >= code_end) {
                if (dst_idx >= dst_len) return false;
                dst[dst_idx++] = src[src_idx++];
                break;
            }
            if (dst_idx >= dst_len) return false;
            dst[dst_idx++] = src[src_idx++];
            if ((src[src_idx] & X86_MASK_JCC) != X86_INSTRUCTION_JCC) {
                if (src[src_idx] == X86_ESCAPE) {
                    src_idx++;
                    if (src_idx >= code_end) return false;
                }
                if (dst_idx >= dst_len) return false;
                dst[dst_idx++] = src[src_idx++];
                continue;
            }
        } else if ((src[src_idx] & X86_MASK_JUMP) != X86_INSTRUCTION_JUMP) {
            if (src[src_idx] == X86_ESCAPE) {
                src_idx++;
                if (src_idx >= code_end) return false;
            }
            if (dst_idx >= dst_len) return false;
            dst[dst_idx++] = src[src_idx++];
            continue;
        }
        if (src_idx + 4 >= code_end) return false;
        if (dst_idx + 5 > dst_len) return false;
        const int64_t addr = int64_t(S.be32(src_idx + 1) ^ MASK_ADDRESS);
        const int64_t offset = addr - dst_idx;
        dst[dst_idx++] = src[src_idx++];
        if (offset >= 0) put_le32(dst + dst_idx, uint32_t(offset));
        else put_le32(dst + dst_idx, uint32_t(-(-offset & X86_ADDR_MASK)));
        src_idx += 4;
        dst_idx += 4;
    }
    if (dst_idx + (count - src_idx) > dst_len) return false;
    if (src_idx < count) {
        copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(count - src_idx));
        dst_idx += count - src_idx;
    }
    *out_n = dst_idx;
    return true;
}

// :615-698
KZ_EXE_HD bool inverse_arm(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t* out_n, bool* panic) {
    const Slice S{src, count, panic};
    int64_t src_idx = 9, dst_idx = 0;
    const int64_t code_start = int64_t(S.le32(1)), code_end = int64_t(S.le32(5));
    if (code_start < 0 || code_end < src_idx || code_end > count || code_start > code_end - src_idx || code_start > dst_len - dst_idx) return false;
    if (code_start > 0) {
        copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(code_start));
        dst_idx += code_start;
        src_idx += code_start;
    }
    while (src_idx < code_end) {
        if (src_idx + 4 > code_end) return false;
        if (dst_idx + 4 > dst_len) return false;
        const int64_t instr = int64_t(S.le32(src_idx));
        const int64_t opcode1 = instr & ARM_B_OPCODE_MASK;
        const bool is_bl = opcode1 == ARM_OPCODE_B || opcode1 == ARM_OPCODE_BL;
        if (!is_bl) {
            copy_bytes(dst + dst_idx, src + src_idx, 4);
            src_idx += 4;
            dst_idx += 4;
            continue;
        }
        const int64_t addr = (instr & ARM_B_ADDR_MASK) << 2;
        const int64_t offset = (addr - dst_idx) >> 2;
        const int64_t val = opcode1 | (offset & ARM_B_ADDR_MASK);
        if (addr == 0) {
            if (src_idx + 8 > code_end) return false;
            copy_bytes(dst + dst_idx, src + src_idx + 4, 4);
            src_idx += 8;
            dst_idx += 4;
            continue;
        }
        put_le32(dst + dst_idx, uint32_t(val));
        src_idx += 4;
        dst_idx += 4;
    }
    if (dst_idx + (count - src_idx) > dst_len) return false;
    if (src_idx < count) {
        copy_bytes(dst + dst_idx, src + src_idx, (uint64_t)(count - src_idx));
        dst_idx += count - src_idx;
    }
    *out_n = dst_idx;
    return true;
}


KZ_EXE_HD int64_t max_encoded_len(int64_t n) { return n <= 256 ? n + 32 : n + n / 8; }  // :701-708

// EXECodec.Forward :109-172. Returns the encoded length, -1 = skip (the reference returns an error), -2 = the reference panics
// (block error). *dt is ctx["dataType"], updated like the reference does. cap = len(dst).
KZ_EXE_HD int64_t forward(const uint8_t* src, int64_t count, uint8_t* dst, int64_t cap, int* dt) {
    if (count == 0) return 0;
    if (count < MIN_BLOCK_SIZE || count > MAX_BLOCK_SIZE) return -1;
    if (cap < max_encoded_len(count)) return -1;
    if (*dt != 0 && *dt != 3 && *dt != 7) return -1;  // DT_UNDEFINED / DT_EXE / DT_BIN
    bool panic = false;
    int64_t code_start = 0, code_end = count - 8;
    const Slice head{src, code_end + 4, &panic};  // src[:codeEnd+4]
    const uint8_t mode0 = detect_exe_type(head, &code_start, &code_end);
    if (panic) return -2;
    if (mode0 & NOT_EXE) {
        *dt = mode0 & MASK_DT;
        return -1;
    }
    const uint8_t mode = (uint8_t)(mode0 & ~MASK_DT);
    int64_t out_n = 0;
    bool ok;
    if (mode == EXE_X86) ok = forward_x86(src, count, dst, cap, code_start, code_end, &out_n, &panic);
    else if (mode == EXE_ARM64) ok = forward_arm(src, count, dst, cap, code_start, code_end, &out_n, &panic);
    else return -1;
    if (panic) return -2;
    if (!ok) return -1;
    *dt = 3;  // DT_EXE
    return out_n;
}

// EXECodec.Inverse :302-331 (bitstream version >= 3). Returns the decoded length or -1 on invalid data.
KZ_EXE_HD int64_t inverse(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
    if (n == 0 || cap == 0) return 0;
    if (n < 9) return -1;
    bool panic = false;
    int64_t out_n = 0;
    bool ok = false;
    if (src[0] == EXE_X86) ok = inverse_x86(src, n, dst, cap, &out_n, &panic);
    else if (src[0] == EXE_ARM64) ok = inverse_arm(src, n, dst, cap, &out_n, &panic);
    return ok && !panic ? out_n : -1;
}

}  // namespace exec
}  // namespace kz
