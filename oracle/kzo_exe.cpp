// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). EXE transform (EXECodec): relative call / jump targets of x86 and ARM64 code
// become absolute addresses. Restated from v2/transform/EXECodec.go: Forward :109-172, forwardX86 :174-297, Inverse :302-331,
// inverseX86 :333-449, forwardARM :507-613, inverseARM :615-698, MaxEncodedLen :701-708, detectExeType :710-810,
// parseExeHeader :812-1021. The pre-v3 bitstream layout (inverseV2 :451-505) is not restated.
// An out-of-range slice access panics in Go and the block task turns the panic into a block error (io/CompressedStream.go:735-743):
// the checked accessors below throw kzo::Error(ERR_PROCESS_BLOCK) at the same places.
#include <algorithm>
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {

namespace {

const uint8_t X86_MASK_JUMP = 0xFE, X86_INSTRUCTION_JUMP = 0xE8, X86_INSTRUCTION_JCC = 0x80, X86_TWO_BYTE_PREFIX = 0x0F, X86_MASK_JCC = 0xF0,
              X86_ESCAPE = 0x9B;
const uint8_t NOT_EXE = 0x80, EXE_X86 = 0x40, EXE_ARM64 = 0x20, MASK_DT = 0x0F;
const int64_t X86_ADDR_MASK = (1 << 24) - 1;
const uint32_t MASK_ADDRESS = 0xF0F0F0F0u;
const int64_t ARM_B_ADDR_MASK = (1 << 26) - 1;
const int64_t ARM_B_OPCODE_MASK = 0xFFFFFFFFll ^ ARM_B_ADDR_MASK;
const int64_t ARM_B_ADDR_SGN_MASK = 1 << 25;
const int64_t ARM_OPCODE_B = 0x14000000, ARM_OPCODE_BL = 0x94000000;
const int64_t ARM_CB_OPCODE_MASK = 0x7F000000, ARM_OPCODE_CBZ = 0x34000000, ARM_OPCODE_CBNZ = 0x3500000;
const int64_t WIN_PE = 0x00004550, WIN_X86_ARCH = 0x014C, WIN_AMD64_ARCH = 0x8664, WIN_ARM64_ARCH = 0xAA64;
const int64_t ELF_X86_ARCH = 0x03, ELF_AMD64_ARCH = 0x3E, ELF_ARM64_ARCH = 0xB7;
const int64_t MAC_AMD64_ARCH = 0x01000007, MAC_ARM64_ARCH = 0x0100000C, MAC_MH_EXECUTE = 0x02, MAC_LC_SEGMENT = 0x01, MAC_LC_SEGMENT64 = 0x19;
const int64_t MIN_BLOCK_SIZE = 4096, MAX_BLOCK_SIZE = (1 << (26 + 2)) - 1;

// a Go slice: indexing outside [0, len) panics
struct Slice {
    const uint8_t* p;
    int64_t len;
    uint8_t at(int64_t i) const {
        if (i < 0 || i >= len) throw Error(ERR_PROCESS_BLOCK, "EXE codec: slice index out of range (the reference panics here)");
        return p[i];
    }
    void need(int64_t i, int64_t n) const {  // src[i:] followed by an n-byte read
        if (i < 0 || i > len || i + n > len) throw Error(ERR_PROCESS_BLOCK, "EXE codec: slice index out of range (the reference panics here)");
    }
    uint32_t le32(int64_t i) const {
        need(i, 4);
        return uint32_t(p[i]) | (uint32_t(p[i + 1]) << 8) | (uint32_t(p[i + 2]) << 16) | (uint32_t(p[i + 3]) << 24);
    }
    uint32_t be32(int64_t i) const {
        need(i, 4);
        return (uint32_t(p[i]) << 24) | (uint32_t(p[i + 1]) << 16) | (uint32_t(p[i + 2]) << 8) | uint32_t(p[i + 3]);
    }
    uint32_t le16(int64_t i) const {
        need(i, 2);
        return uint32_t(p[i]) | (uint32_t(p[i + 1]) << 8);
    }
    uint32_t be16(int64_t i) const {
        need(i, 2);
        return (uint32_t(p[i]) << 8) | uint32_t(p[i + 1]);
    }
    uint64_t le64(int64_t i) const {
        need(i, 8);
        return uint64_t(le32(i)) | (uint64_t(le32(i + 4)) << 32);
    }
    uint64_t be64(int64_t i) const {
        need(i, 8);
        return (uint64_t(be32(i)) << 32) | uint64_t(be32(i + 4));
    }
};

const uint32_t WIN_MAGIC = 0x4D5A, ELF_MAGIC = 0x7F454C46u, MAC_MAGIC32 = 0xFEEDFACEu, MAC_CIGAM32 = 0xCEFAEDFEu, MAC_MAGIC64 = 0xFEEDFACFu,
               MAC_CIGAM64 = 0xCFFAEDFEu;

// :812-1021
bool parse_exe_header(const Slice& src, uint32_t magic, int64_t* arch, int64_t* code_start, int64_t* code_end) {
    const int64_t count = src.len;
    if (magic == WIN_MAGIC) {
        if (count >= 64) {
            const int64_t pos_pe = int64_t(src.le32(60));
            if (pos_pe > 0 && pos_pe <= count - 48 && int64_t(src.le32(pos_pe)) == WIN_PE) {
                *code_start = std::min<int64_t>(int64_t(src.le32(pos_pe + 44)), count);
                *code_end = std::min<int64_t>(*code_start + int64_t(src.le32(pos_pe + 28)), count);
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
            *code_start = std::min(*code_start, count);
            *code_end = std::min(*code_end, count);
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
            *code_start = std::min(*code_start, count);
            *code_end = std::min(*code_end, count);
            return true;
        }
    }
    return false;
}

// :710-810
uint8_t detect_exe_type(const Slice& src, int64_t* code_start, int64_t* code_end) {
    const uint32_t magic = get_magic_type(src.p, size_t(src.len));
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
    const int dt = detect_simple_type(size_t(std::max<int64_t>(count, 0)), histo);
    if (dt != DT_BIN) return uint8_t(NOT_EXE | dt);
    int64_t small_vals = 0;
    for (int i = 0; i < 16; i++) small_vals += histo[i];
    if (histo[0] < count / 10 || small_vals > count / 2 || histo[255] < count / 100) return uint8_t(NOT_EXE | dt);
    if (jumps_x86 >= count / 200) return EXE_X86;
    if (jumps_arm64 >= count / 200) return EXE_ARM64;
    return uint8_t(NOT_EXE | dt);
}

void put_le32(uint8_t* p, uint32_t v) {
    p[0] = uint8_t(v);
    p[1] = uint8_t(v >> 8);
    p[2] = uint8_t(v >> 16);
    p[3] = uint8_t(v >> 24);
}
void put_be32(uint8_t* p, uint32_t v) {
    p[0] = uint8_t(v >> 24);
    p[1] = uint8_t(v >> 16);
    p[2] = uint8_t(v >> 8);
    p[3] = uint8_t(v);
}

// :174-297
bool forward_x86(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t code_start, int64_t code_end, size_t* out_n) {
    int64_t src_idx = code_start, dst_idx = 9, matches = 0;
    const int64_t dst_end = dst_len - 5;
    dst[0] = EXE_X86;
    bool boundary = false;
    if (code_start < 0 || code_end < code_start || code_end > count) return false;
    if (code_start > 0) {
        std::memcpy(dst + dst_idx, src, size_t(code_start));
        dst_idx += code_start;
    }
    const Slice S{src, count};
    while (src_idx < code_end && dst_idx < dst_end) {
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
    std::memcpy(dst + dst_idx, src + src_idx, size_t(count - src_idx));
    dst_idx += count - src_idx;
    if (dst_idx > count + count / 50) return false;
    *out_n = size_t(dst_idx);
    return true;
}

// :507-613 (the conditional-branch form is disabled in the reference: isCB is always false)
bool forward_arm(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, int64_t code_start, int64_t code_end, size_t* out_n) {
    int64_t src_idx = code_start, dst_idx = 9, matches = 0;
    const int64_t dst_end = dst_len - 8;
    dst[0] = EXE_ARM64;
    if (code_start < 0 || code_end < code_start || code_end > count) return false;
    if (code_start > 0) {
        std::memcpy(dst + dst_idx, src, size_t(code_start));
        dst_idx += code_start;
    }
    const Slice S{src, count};
    while (src_idx + 4 <= code_end && dst_idx < dst_end) {
        const int64_t instr = int64_t(S.le32(src_idx));
        const int64_t opcode1 = instr & ARM_B_OPCODE_MASK;
        const bool is_bl = opcode1 == ARM_OPCODE_B || opcode1 == ARM_OPCODE_BL;
        if (!is_bl) {
            std::memcpy(dst + dst_idx, src + src_idx, 4);
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
            std::memcpy(dst + dst_idx + 4, src + src_idx, 4);
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
    std::memcpy(dst + dst_idx, src + src_idx, size_t(count - src_idx));
    dst_idx += count - src_idx;
    if (dst_idx > count + count / 50) return false;
    *out_n = size_t(dst_idx);
    return true;
}

// :333-449
bool inverse_x86(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, size_t* out_n) {
    const Slice S{src, count};
    int64_t src_idx = 9, dst_idx = 0;
    const int64_t code_start = int64_t(S.le32(1)), code_end = int64_t(S.le32(5));
    if (code_start < 0 || code_end < src_idx || code_end > count || code_start > code_end - src_idx || code_start > dst_len - dst_idx) return false;
    if (code_start > 0) {
        std::memcpy(dst + dst_idx, src + src_idx, size_t(code_start));
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
        std::memcpy(dst + dst_idx, src + src_idx, size_t(count - src_idx));
        dst_idx += count - src_idx;
    }
    *out_n = size_t(dst_idx);
    return true;
}

// :615-698
bool inverse_arm(const uint8_t* src, int64_t count, uint8_t* dst, int64_t dst_len, size_t* out_n) {
    const Slice S{src, count};
    int64_t src_idx = 9, dst_idx = 0;
    const int64_t code_start = int64_t(S.le32(1)), code_end = int64_t(S.le32(5));
    if (code_start < 0 || code_end < src_idx || code_end > count || code_start > code_end - src_idx || code_start > dst_len - dst_idx) return false;
    if (code_start > 0) {
        std::memcpy(dst + dst_idx, src + src_idx, size_t(code_start));
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
            std::memcpy(dst + dst_idx, src + src_idx, 4);
            src_idx += 4;
            dst_idx += 4;
            continue;
        }
        const int64_t addr = (instr & ARM_B_ADDR_MASK) << 2;
        const int64_t offset = (addr - dst_idx) >> 2;
        const int64_t val = opcode1 | (offset & ARM_B_ADDR_MASK);
        if (addr == 0) {
            if (src_idx + 8 > code_end) return false;
            std::memcpy(dst + dst_idx, src + src_idx + 4, 4);
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
        std::memcpy(dst + dst_idx, src + src_idx, size_t(count - src_idx));
        dst_idx += count - src_idx;
    }
    *out_n = size_t(dst_idx);
    return true;
}

}  // namespace

size_t exe_max_encoded_len(size_t n) { return n <= 256 ? n + 32 : n + n / 8; }  // :701-708

bool exe_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    const int64_t count = int64_t(n);
    if (count < MIN_BLOCK_SIZE || count > MAX_BLOCK_SIZE) return false;
    if (cap < exe_max_encoded_len(n)) return false;
    if (ctx.data_type != DT_UNDEFINED && ctx.data_type != DT_EXE && ctx.data_type != DT_BIN) return false;
    int64_t code_start = 0, code_end = count - 8;
    const Slice head{src, code_end + 4};  // src[:codeEnd+4]
    const uint8_t mode0 = detect_exe_type(head, &code_start, &code_end);
    if (mode0 & NOT_EXE) {
        ctx.data_type = mode0 & MASK_DT;
        return false;
    }
    const uint8_t mode = uint8_t(mode0 & ~MASK_DT);
    // the reference hands dst with its full length to the coders: len(dst) is what the sequence allocated, MaxEncodedLen at least
    bool ok;
    if (mode == EXE_X86) ok = forward_x86(src, count, dst, int64_t(cap), code_start, code_end, out_n);
    else if (mode == EXE_ARM64) ok = forward_arm(src, count, dst, int64_t(cap), code_start, code_end, out_n);
    else return false;
    if (ok) ctx.data_type = DT_EXE;
    else *out_n = 0;
    return ok;
}

bool exe_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0 || cap == 0) return true;
    if (n < 9) return false;
    if (src[0] == EXE_X86) return inverse_x86(src, int64_t(n), dst, int64_t(cap), out_n);
    if (src[0] == EXE_ARM64) return inverse_arm(src, int64_t(n), dst, int64_t(cap), out_n);
    return false;
}

}  // namespace kzo
