// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Reduced offset LZ with ANS coded streams (kanzi ROLZ, rolzCodec1).
// Follows v2/transform/ROLZCodec.go line by line: :57-86 keys / hash / emitCopy, :246-296 findMatch, :301-596 Forward,
// :601-913 Inverse, :916-971 MaxEncodedLen / emitLengthROLZ / readLengthROLZ; internal/Global.go:346-419 DetectSimpleType.
#include <algorithm>
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {
namespace {
const uint32_t HASH_SIZE = 1 << 16;                       // :34
const int MIN_MATCH3 = 3, MIN_MATCH4 = 4, MIN_MATCH7 = 7;  // :35-37
const int MAX_MATCH1 = MIN_MATCH3 + 65535;                // :38
const int LOG_POS_CHECKS1 = 4;                            // :40
const int CHUNK_SIZE = 16 * 1024 * 1024;                  // :42
const uint32_t HASH_MASK = ~uint32_t(CHUNK_SIZE - 1);     // :43
const uint64_t HASH_SEED = 200002979;                     // :48
const size_t MAX_BLOCK = size_t(1) << 30, MIN_BLOCK = 64;  // :49-50

inline uint32_t le16(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8); }
inline uint32_t le32(const uint8_t* p) {
    uint32_t v;
    std::memcpy(&v, p, 4);
    return v;
}
inline uint64_t le64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;
}
inline uint32_t get_key1(const uint8_t* p) { return le16(p); }                                             // :57-59
inline uint32_t get_key2(const uint8_t* p) { return uint32_t((le64(p) * HASH_SEED) >> 40) & 0xFFFF; }      // :61-63
inline uint32_t rolz_hash(const uint8_t* p) { return ((le32(p) << 8) * uint32_t(HASH_SEED)) & HASH_MASK; }  // :65-67

int emit_length(uint8_t* block, int lit_len) {  // :921-943
    int idx = 0;
    if (lit_len >= 1 << 7) {
        if (lit_len >= 1 << 14) {
            if (lit_len >= 1 << 21) block[idx++] = uint8_t(0x80 | (lit_len >> 21));
            block[idx++] = uint8_t(0x80 | (lit_len >> 14));
        }
        block[idx++] = uint8_t(0x80 | (lit_len >> 7));
    }
    block[idx] = uint8_t(lit_len & 0x7F);
    return idx + 1;
}
int read_length(const uint8_t* buf, int* idx_out) {  // :946-971
    uint8_t next = buf[0];
    int idx = 1;
    int lit_len = next & 0x7F;
    if (next >= 128) {
        next = buf[idx++];
        lit_len = (lit_len << 7) | (next & 0x7F);
        if (next >= 128) {
            next = buf[idx++];
            lit_len = (lit_len << 7) | (next & 0x7F);
            if (next >= 128) {
                next = buf[idx++];
                lit_len = (lit_len << 7) | (next & 0x7F);
            }
        }
    }
    *idx_out = idx;
    return lit_len;
}

struct Rolz1 {
    std::vector<uint32_t> matches;
    std::vector<int32_t> counters;
    unsigned log_pos_checks = LOG_POS_CHECKS1;
    int32_t mask_checks = (1 << LOG_POS_CHECKS1) - 1, pos_checks = 1 << LOG_POS_CHECKS1;
    int min_match = MIN_MATCH3;

    // :246-296. buf_len = len(buf) of the reference's chunk slice
    void find_match(const uint8_t* buf, int buf_len, int pos, uint32_t hash32, int32_t counter, const uint32_t* m, int* idx, int* len) const {
        int max_match = std::min(MAX_MATCH1, buf_len - pos);
        *idx = -1;
        *len = -1;
        if (max_match < min_match) return;
        max_match -= 8;
        int best_len = 0, best_idx = -1;
        const uint8_t* cur = buf + pos;
        for (int32_t i = counter; i > counter - pos_checks; i--) {
            uint32_t ref = m[i & mask_checks];
            if ((ref & HASH_MASK) != hash32) continue;
            ref &= ~HASH_MASK;
            const uint8_t* rb = buf + ref;
            if (rb[best_len] != cur[best_len]) continue;
            int n = 0;
            while (n < max_match) {
                const uint64_t diff = le64(rb + n) ^ le64(cur + n);
                if (diff != 0) {
                    n += __builtin_ctzll(diff) >> 3;
                    break;
                }
                n += 8;
            }
            if (n > best_len) {
                best_idx = int(i);
                best_len = n;
            }
        }
        if (best_len < min_match) return;
        *idx = int(counter) - best_idx;
        *len = best_len - min_match;
    }
};

size_t rolz_max_len(size_t n) { return n <= 512 ? n + 64 : n; }  // :916-918
}  // namespace

int detect_simple_type(size_t count, const int* freqs0) {  // internal/Global.go:346-419
    if (count == 0) return DT_UNDEFINED;
    static const char DNA[] = "acgntuACGNTU\"";
    static const char NUMERIC[] = "0123456789+-*/=,.:; ";
    static const char BASE64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    int64_t sum = 0;
    for (int i = 0; i < 12; i++) sum += freqs0[uint8_t(DNA[i])];
    if (sum > int64_t(count) - int64_t(count) / 12) return DT_DNA;
    sum = 0;
    for (int i = 0; i < 20; i++) sum += freqs0[uint8_t(NUMERIC[i])];
    if (sum == int64_t(count)) return DT_NUMERIC;
    sum = 0;
    for (int i = 0; i < 64; i++) sum += freqs0[uint8_t(BASE64[i])];
    if (sum + freqs0[0x3D] == int64_t(count)) return DT_BASE64;
    sum = 0;
    for (int i = 0; i < 256; i++)
        if (freqs0[i] > 0) sum++;
    if (sum == 256) return DT_BIN;
    if (sum <= 4) return DT_SMALL_ALPHABET;
    return DT_UNDEFINED;
}

size_t rolz_max_encoded_len(size_t n) { return rolz_max_len(n); }

// ROLZCodec.Forward :150-168 + rolzCodec1.Forward :301-596
bool rolz_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (n < MIN_BLOCK || n > MAX_BLOCK) return false;
    if (cap < rolz_max_len(n)) return false;
    Rolz1 c;
    c.counters.assign(1 << 16, 0);
    const int len_src = int(n);
    const int src_end = len_src - 4;
    dst[0] = uint8_t(n >> 24), dst[1] = uint8_t(n >> 16), dst[2] = uint8_t(n >> 8), dst[3] = uint8_t(n);
    int size_chunk = std::min(len_src, CHUNK_SIZE);
    int start_chunk = 0;
    std::vector<uint8_t> lit_buf(rolz_max_len(size_t(size_chunk))), len_buf(size_t(size_chunk) / 5 + 8), midx_buf(size_t(size_chunk) / 4 + 8),
        tk_buf(size_t(size_chunk) / 4 + 8);
    unsigned lit_order = len_src < (1 << 17) ? 0 : 1;
    uint8_t flags = uint8_t(lit_order);
    c.min_match = MIN_MATCH3;
    int delta = 2;
    {  // ctx != nil in every factory-made codec (Factory.go:143-147)
        int dt = ctx.data_type;
        if (dt == DT_UNDEFINED) {
            int freqs0[256];
            std::memset(freqs0, 0, sizeof(freqs0));
            histogram_order0(src, n, freqs0);
            dt = detect_simple_type(n, freqs0);
            if (dt != DT_UNDEFINED) ctx.data_type = dt;
        }
        if (dt == DT_EXE) {
            delta = 3;
            flags |= 8;
        } else if (dt == DT_DNA) {
            delta = 8;
            c.min_match = MIN_MATCH7;
            flags |= 4;
        } else if (dt == DT_MULTIMEDIA) {
            delta = 8;
            c.min_match = MIN_MATCH4;
            flags |= 2;
        }
    }
    flags |= uint8_t(c.log_pos_checks << 4);
    dst[4] = flags;
    int src_idx = 0;
    size_t dst_idx = 5;
    c.matches.assign(size_t(HASH_SIZE) << c.log_pos_checks, 0);
    bool ok = true;
    while (start_chunk < src_end) {
        int lit_idx = 0, len_idx = 0, m_idx = 0, tk_idx = 0;
        std::fill(c.matches.begin(), c.matches.end(), 0u);
        int end_chunk = start_chunk + size_chunk;
        if (end_chunk >= src_end) {
            end_chunk = src_end;
            size_chunk = end_chunk - start_chunk;
        }
        const uint8_t* buf = src + start_chunk;  // len(buf) = size_chunk, readable up to the end of src
        src_idx = 0;
        const int nfirst = std::min(src_end - start_chunk, 8);
        for (int j = 0; j < nfirst; j++) lit_buf[size_t(lit_idx++)] = buf[src_idx++];
        int first_lit_idx = src_idx;
        int src_inc = 0;
        while (src_idx < size_chunk) {
            uint32_t key = c.min_match == MIN_MATCH3 ? get_key1(buf + src_idx - delta) : get_key2(buf + src_idx - delta);
            uint32_t* m = &c.matches[size_t(key) << c.log_pos_checks];
            uint32_t hash32 = rolz_hash(buf + src_idx);
            int match_idx, match_len;
            c.find_match(buf, size_chunk, src_idx, hash32, c.counters[key], m, &match_idx, &match_len);
            c.counters[key] = (c.counters[key] + 1) & c.mask_checks;  // register current position
            m[c.counters[key]] = hash32 | uint32_t(src_idx);
            if (match_idx < 0) {
                src_idx++;
                src_idx += src_inc >> 6;
                src_inc++;
                continue;
            }
            const int src_idx1 = src_idx + 1;  // check if better match at next position
            key = c.min_match == MIN_MATCH3 ? get_key1(buf + src_idx1 - delta) : get_key2(buf + src_idx1 - delta);
            m = &c.matches[size_t(key) << c.log_pos_checks];
            hash32 = rolz_hash(buf + src_idx1);
            int match_idx1, match_len1;
            c.find_match(buf, size_chunk, src_idx1, hash32, c.counters[key], m, &match_idx1, &match_len1);
            if (match_idx1 >= 0 && match_len1 > match_len) {
                match_idx = match_idx1;
                match_len = match_len1;
                src_idx = src_idx1;
                c.counters[key] = (c.counters[key] + 1) & c.mask_checks;
                m[c.counters[key]] = hash32 | uint32_t(src_idx);
            }
            const int lit_len = src_idx - first_lit_idx;  // token LLLLLMMM
            uint8_t token;
            if (match_len >= 7) {
                token = 7;
                len_idx += emit_length(&len_buf[size_t(len_idx)], match_len - 7);
            } else {
                token = uint8_t(match_len);
            }
            if (lit_len > 0) {
                if (lit_len >= 31) {
                    token |= 0xF8;
                    len_idx += emit_length(&len_buf[size_t(len_idx)], lit_len - 31);
                } else {
                    token |= uint8_t(lit_len << 3);
                }
                std::memcpy(&lit_buf[size_t(lit_idx)], buf + first_lit_idx, size_t(lit_len));
                lit_idx += lit_len;
            }
            tk_buf[size_t(tk_idx++)] = token;
            midx_buf[size_t(m_idx++)] = uint8_t(match_idx);
            src_idx += match_len + c.min_match;
            first_lit_idx = src_idx;
            src_inc = 0;
        }
        src_idx = size_chunk;  // emit last chunk literals
        const int lit_len = src_idx - first_lit_idx;
        if (tk_idx != 0) tk_buf[size_t(tk_idx++)] = lit_len >= 31 ? uint8_t(0xF8) : uint8_t(lit_len << 3);
        if (lit_len > 0) {
            if (lit_len >= 31) len_idx += emit_length(&len_buf[size_t(len_idx)], lit_len - 31);
            std::memcpy(&lit_buf[size_t(lit_idx)], buf + first_lit_idx, size_t(lit_len));
            lit_idx += lit_len;
        }
        BitWriter obs;  // private bitstream of the chunk (:509-555)
        obs.write_bits(uint64_t(lit_idx), 32);
        obs.write_bits(uint64_t(tk_idx), 32);
        obs.write_bits(uint64_t(len_idx), 32);
        obs.write_bits(uint64_t(m_idx), 32);
        ans_encode(obs, lit_buf.data(), size_t(lit_idx), int(lit_order));
        ans_encode(obs, tk_buf.data(), size_t(tk_idx), 0, 32768);
        ans_encode(obs, len_buf.data(), size_t(len_idx), 0, 32768);
        ans_encode(obs, midx_buf.data(), size_t(m_idx), 0, 32768);
        const std::vector<uint8_t>& bytes = obs.close();
        if (dst_idx + bytes.size() > cap) {
            ok = false;
            break;
        }
        std::memcpy(dst + dst_idx, bytes.data(), bytes.size());
        dst_idx += bytes.size();
        start_chunk = end_chunk;
    }
    if (!ok) return false;
    if (dst_idx + 4 > cap) return false;
    src_idx += start_chunk - size_chunk;  // emit last literals
    dst[dst_idx] = src[src_idx], dst[dst_idx + 1] = src[src_idx + 1], dst[dst_idx + 2] = src[src_idx + 2], dst[dst_idx + 3] = src[src_idx + 3];
    src_idx += 4;
    dst_idx += 4;
    if (src_idx != len_src) return false;
    if (dst_idx >= n) return false;  // no compression
    *out_n = dst_idx;
    return true;
}

// ROLZCodec.Inverse :173-191 + rolzCodec1.Inverse :601-913 (bitstream version 6)
bool rolz_inverse(Ctx&, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0 || cap == 0) return true;
    if (n < 5 || n > MAX_BLOCK) return false;
    const int len_src = int(n), len_dst = int(std::min<size_t>(cap, MAX_BLOCK + 4096));
    const int dst_end = int((uint32_t(src[0]) << 24) | (uint32_t(src[1]) << 16) | (uint32_t(src[2]) << 8) | src[3]) - 4;
    if (dst_end <= 0 || dst_end > len_dst) return false;
    int start_chunk = 0, src_idx = 5, dst_idx = 0;
    int size_chunk = std::min(len_dst, CHUNK_SIZE);
    std::vector<uint8_t> lit_buf(size_t(size_chunk) + 8), mlen_buf(size_t(size_chunk) / 5 + 8), midx_buf(size_t(size_chunk) / 4 + 8),
        tk_buf(size_t(size_chunk) / 4 + 8);
    const int lit_cap = size_chunk, mlen_cap = size_chunk / 5, midx_cap = size_chunk / 4, tk_cap = size_chunk / 4;
    Rolz1 c;
    c.counters.assign(1 << 16, 0);
    const uint8_t flags = src[4];
    const unsigned lit_order = flags & 1;
    int delta = 2;
    c.min_match = MIN_MATCH3;
    if ((flags & 0x0E) == 2) {
        c.min_match = MIN_MATCH4;
        delta = 8;
    } else if ((flags & 0x0E) == 4) {
        c.min_match = MIN_MATCH7;
        delta = 8;
    } else if ((flags & 0x0E) == 8) {
        delta = 3;
    }
    c.log_pos_checks = unsigned(flags >> 4);
    if (c.log_pos_checks < 2 || c.log_pos_checks > 8) return false;
    c.pos_checks = 1 << c.log_pos_checks;
    c.mask_checks = c.pos_checks - 1;
    c.matches.assign(size_t(HASH_SIZE) << c.log_pos_checks, 0);
    try {
        while (start_chunk < dst_end) {
            int m_idx = 0, len_idx = 0, lit_idx = 0, tk_idx = 0;
            std::fill(c.matches.begin(), c.matches.end(), 0u);
            int end_chunk = start_chunk + size_chunk;
            if (end_chunk > dst_end) end_chunk = dst_end;
            size_chunk = end_chunk - start_chunk;
            uint8_t* buf = dst + start_chunk;
            bool only_literals;
            int lit_len_decoded;
            {
                BitReader ibs = BitReader::over_bytes(src + src_idx, size_t(len_src - src_idx));
                const int lit_len = int(int32_t(ibs.read_bits(32)));
                const int tk_len = int(int32_t(ibs.read_bits(32)));
                const int mlen_len = int(int32_t(ibs.read_bits(32)));
                const int midx_len = int(int32_t(ibs.read_bits(32)));
                const int first_lit_len = std::min(size_chunk, 8);
                if (lit_len < 0 || lit_len > lit_cap) return false;
                if (tk_len < 0 || tk_len > tk_cap) return false;
                if (mlen_len < 0 || mlen_len > mlen_cap) return false;
                if (midx_len < 0 || midx_len > midx_cap) return false;
                if (lit_len < first_lit_len || lit_len > size_chunk) return false;
                if ((tk_len == 0 && midx_len != 0) || (tk_len > 0 && midx_len + 1 != tk_len)) return false;
                lit_len_decoded = lit_len;
                ans_decode(ibs, lit_buf.data(), size_t(lit_len), int(lit_order));
                ans_decode(ibs, tk_buf.data(), size_t(tk_len), 0, 32768);
                ans_decode(ibs, mlen_buf.data(), size_t(mlen_len), 0, 32768);
                ans_decode(ibs, midx_buf.data(), size_t(midx_len), 0, 32768);
                only_literals = tk_len == 0;
                src_idx += int((ibs.read() + 7) >> 3);
            }
            if (only_literals) {
                if (lit_len_decoded != size_chunk) return false;
                std::memcpy(buf, lit_buf.data(), size_t(size_chunk));
                dst_idx = size_chunk;
                start_chunk = end_chunk;
                continue;
            }
            dst_idx = 0;
            int mm = 8;
            if (start_chunk >= dst_end) mm = dst_end - start_chunk;
            for (int j = 0; j < mm; j++) buf[dst_idx++] = lit_buf[size_t(lit_idx++)];
            while (dst_idx < size_chunk) {
                const uint8_t token = tk_buf[size_t(tk_idx++)];
                int match_len = token & 0x07;
                if (match_len == 7) {
                    int d;
                    match_len = read_length(&mlen_buf[size_t(len_idx)], &d) + 7;
                    len_idx += d;
                }
                int lit_len;
                if (token < 0xF8) {
                    lit_len = token >> 3;
                } else {
                    int d;
                    lit_len = read_length(&mlen_buf[size_t(len_idx)], &d) + 31;
                    len_idx += d;
                }
                if (lit_len > 0) {
                    if (dst_idx + lit_len > lit_cap) return false;
                    int src_inc = 0;
                    uint8_t* d = buf + dst_idx - delta;
                    std::memcpy(d + delta, &lit_buf[size_t(lit_idx)], size_t(lit_len));
                    for (int k = 0; k < lit_len; k++) {
                        const uint32_t key = c.min_match == MIN_MATCH3 ? get_key1(d + k) : get_key2(d + k);
                        const int32_t cc = (c.counters[key] + 1) & c.mask_checks;
                        c.matches[(size_t(key) << c.log_pos_checks) + size_t(cc)] = uint32_t(dst_idx + k);
                        c.counters[key] = cc;
                        k += src_inc >> 6;
                        src_inc++;
                    }
                    lit_idx += lit_len;
                    dst_idx += lit_len;
                    if (dst_idx >= size_chunk) {
                        if (dst_idx == size_chunk) break;  // last chunk literals not followed by match
                        return false;
                    }
                }
                if (dst_idx + match_len + c.min_match > dst_end) return false;  // sanity check
                const int32_t match_idx = int32_t(midx_buf[size_t(m_idx++)] & 0xFF);
                const uint32_t key = c.min_match == MIN_MATCH3 ? get_key1(buf + dst_idx - delta) : get_key2(buf + dst_idx - delta);
                uint32_t* m = &c.matches[size_t(key) << c.log_pos_checks];
                int ref = int(m[(c.counters[key] - match_idx) & c.mask_checks]);
                c.counters[key] = (c.counters[key] + 1) & c.mask_checks;
                m[c.counters[key]] = uint32_t(dst_idx);
                int ml = match_len + c.min_match;  // emitCopy :69-84
                if (dst_idx >= ref + ml) {
                    std::memcpy(buf + dst_idx, buf + ref, size_t(ml));
                    dst_idx += ml;
                } else {
                    while (ml != 0) {
                        buf[dst_idx++] = buf[ref++];
                        ml--;
                    }
                }
            }
            start_chunk = end_chunk;
        }
    } catch (const Error&) {
        return false;
    }
    dst_idx += start_chunk - size_chunk;  // emit last literals
    if (dst_idx + 4 > len_dst || len_src - src_idx != 4) return false;
    dst[dst_idx] = src[src_idx], dst[dst_idx + 1] = src[src_idx + 1], dst[dst_idx + 2] = src[src_idx + 2], dst[dst_idx + 3] = src[src_idx + 3];
    src_idx += 4;
    dst_idx += 4;
    if (src_idx != len_src) return false;
    *out_n = size_t(dst_idx);
    return true;
}

}  // namespace kzo
