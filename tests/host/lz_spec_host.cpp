// NEXT ROUND — CPU model of the warp-speculative LZ / LZX parse planned for lz_parse_kernel (tests/test_lz_spec_host.py compares it with the
// oracle byte for byte). The serial parse probes one position at a time (hash slot, candidate bytes: two dependent memory round trips per
// literal); here a batch of 32 lanes probes the next 32 positions of the literal run at once, against the table as it was at the start
// of the batch plus the inserts of the earlier lanes of the batch (a lane that shares its hash with an earlier lane sees that lane's
// position, like the serial parse would), the first lane with a match wins, the lanes before it commit their inserts, and the match
// itself (lazy evaluation, backward extension, tokens, in-match inserts) is processed as before.
// Same format logic as the parse of kz_lz.cu (reference: v2/transform/LZCodec.go:249-591).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
const uint64_t HASH_SEED = 0x1E35A7BD;
const int MAX_DISTANCE1 = (1 << 16) - 2, MAX_DISTANCE2 = (1 << 24) - 2, MAX_MATCH = 65535 + 254 + 4, MIN_BLOCK_LENGTH = 24, LANES = 32;
inline uint64_t le64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t le32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint32_t lz_hash(uint64_t v, bool extra) { return uint32_t(((v << 24) * HASH_SEED) >> (64 - (extra ? 19 : 16))); }
int emit_length(uint8_t* block, int length) {
    if (length < 254) { block[0] = uint8_t(length); return 1; }
    if (length < 65536 + 254) { length -= 254; block[0] = 254; block[1] = uint8_t(length >> 8); block[2] = uint8_t(length); return 3; }
    length -= 255; block[0] = 255; block[1] = uint8_t(length >> 16); block[2] = uint8_t(length >> 8); block[3] = uint8_t(length); return 4;
}
int find_match(const uint8_t* src, int src_idx, int ref, int max_match) {
    int best = 0;
    while (best + 8 <= max_match) {
        const uint64_t diff = le64(src + src_idx + best) ^ le64(src + ref + best);
        if (diff != 0) { best += __builtin_ctzll(diff) >> 3; break; }
        best += 8;
    }
    return best;
}
}  // namespace

extern "C" {
// stats[0] = batches, stats[1] = lanes that were needed (positions the serial parse visits as literals), stats[2] = matches
int64_t lz_forward_spec(int extra_i, int data_type, const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap, int64_t* stats) {
    const bool extra = extra_i != 0;
    if (n == 0) return 0;
    const int count = int(n);
    if (cap < (n <= 1024 ? n + 16 : n + n / 64)) return -1;
    if (count < MIN_BLOCK_LENGTH) return -1;
    std::vector<int32_t> hashes(size_t(1) << (extra ? 19 : 16), 0);
    std::vector<uint8_t> m_len_buf(size_t(n) + 64), m_buf(size_t(n) + 64), tk_buf(size_t(n) + 64);
    const int src_end = count - 16 - 2;
    int max_dist = MAX_DISTANCE2;
    dst[12] = 1;
    if (src_end < 4 * MAX_DISTANCE1) { max_dist = MAX_DISTANCE1; dst[12] = 0; }
    int min_match = 4;
    if (data_type == 6) min_match = 6;
    else if (data_type == 9) return -1;
    dst[12] |= uint8_t(((min_match - 2) & 7) << 1);
    int src_idx = 0, dst_idx = 13, anchor = 0, m_len_idx = 0, m_idx = 0, tk_idx = 0;
    int repd[2] = {count, count};
    int repd_idx = 0, src_inc = 0;
    while (src_idx < src_end) {
        // ---- speculative batch: the positions the serial parse visits if every earlier lane misses
        int P[LANES], nl = 0;
        uint32_t H[LANES];
        int ref0[LANES];
        {
            int pos = src_idx, inc = src_inc;
            while (nl < LANES && pos < src_end) {
                P[nl] = pos;
                pos = pos + 1 + (inc >> 6);
                inc++;
                nl++;
            }
        }
        for (int j = 0; j < nl; j++) {
            H[j] = lz_hash(le64(src + P[j]), extra);
            ref0[j] = hashes[H[j]];  // table at the start of the batch ...
        }
        for (int j = 1; j < nl; j++)  // ... unless an earlier lane of the batch owns the same slot (on the GPU: __match_any_sync)
            for (int i = j - 1; i >= 0; i--)
                if (H[i] == H[j]) { ref0[j] = P[i]; break; }
        int w = -1, w_best = 0, w_ref = 0;
        bool w_rep = false;
        for (int j = 0; j < nl && w < 0; j++) {  // every lane probes on its own; the first hit wins (ballot + ffs)
            const int pj = P[j], rj = j == 0 ? repd_idx : 0;
            const uint64_t p = le64(src + pj);
            const int pj1 = pj + 1;
            const int max_match = std::min(src_end - pj1, MAX_MATCH);
            const int min_ref = std::max(pj - max_dist, 0);
            int best_len = 0;
            int ref = pj1 - repd[rj];
            if (ref > min_ref && uint32_t(p >> 8) == le32(src + ref)) {
                best_len = find_match(src, pj1, ref, max_match);
            } else {
                ref = pj1 - repd[rj ^ 1];
                if (ref > min_ref && uint32_t(p >> 8) == le32(src + ref)) best_len = find_match(src, pj1, ref, max_match);
            }
            if (best_len >= min_match) {
                w = j; w_best = best_len; w_ref = ref; w_rep = true;
            } else {
                ref = ref0[j];
                if (ref > min_ref && uint32_t(p) == le32(src + ref)) {
                    best_len = find_match(src, pj, ref, std::min(src_end - pj, MAX_MATCH));
                    if (best_len >= min_match) { w = j; w_best = best_len; w_ref = ref; w_rep = false; }
                }
            }
        }
        const int ncommit = w < 0 ? nl : w + 1;
        for (int i = 0; i < ncommit; i++) hashes[H[i]] = P[i];  // in lane order: the last lane sharing a slot wins (atomicMax on the GPU)
        stats[0]++;
        stats[1] += ncommit;
        if (w < 0) {  // thirty-two misses
            src_idx = P[nl - 1] + 1 + ((src_inc + nl - 1) >> 6);
            src_inc += nl;
            repd_idx = 0;
            continue;
        }
        stats[2]++;
        src_inc += w;
        if (w > 0) repd_idx = 0;
        src_idx = P[w];
        // ---- the match, exactly as in the serial parse
        int best_len = w_best, ref = w_ref;
        const int src_idx1 = src_idx + 1;
        const int max_match = std::min(src_end - src_idx1, MAX_MATCH);
        const int min_ref = std::max(src_idx - max_dist, 0);
        if (!w_rep) {
            if (ref != src_idx - repd[0] && ref != src_idx - repd[1]) {
                const uint32_t h1 = lz_hash(le64(src + src_idx1), extra);
                const int ref1 = hashes[h1];
                hashes[h1] = src_idx1;
                if (ref1 > min_ref + 1 && le32(src + src_idx1 + best_len - 3) == le32(src + ref1 + best_len - 3)) {
                    const int bl1 = find_match(src, src_idx1, ref1, max_match);
                    if (bl1 >= best_len) { ref = ref1; best_len = bl1; src_idx = src_idx1; }
                }
                if (extra) {
                    const int src_idx2 = src_idx1 + 1;
                    const uint32_t h2 = lz_hash(le64(src + src_idx2), extra);
                    const int ref2 = hashes[h2];
                    hashes[h2] = src_idx2;
                    if (ref2 > min_ref + 2 && le32(src + src_idx2 + best_len - 3) == le32(src + ref2 + best_len - 3)) {
                        const int bl2 = find_match(src, src_idx2, ref2, std::min(src_end - src_idx2, MAX_MATCH));
                        if (bl2 >= best_len) { ref = ref2; best_len = bl2; src_idx = src_idx2; }
                    }
                }
            }
            while (src_idx > anchor && ref > min_ref && src[src_idx - 1] == src[ref - 1]) { best_len++; ref--; src_idx--; }
            if (best_len > MAX_MATCH) { src_idx += best_len - MAX_MATCH; ref += best_len - MAX_MATCH; best_len = MAX_MATCH; }
        } else {
            if (src[src_idx] == src[ref - 1] && best_len < MAX_MATCH) { best_len++; ref--; }
            else { src_idx++; hashes[lz_hash(le64(src + src_idx), extra)] = src_idx; }
        }
        src_inc = 0;
        const int dist = src_idx - ref;
        const int m_len = best_len - min_match;
        int token, m_len_th;
        if (dist == repd[0]) { token = 0x00; m_len_th = 3; }
        else if (dist == repd[1]) { token = 0x04; m_len_th = 3; }
        else {
            m_len_th = 7;
            if (dist >= 256) {
                if (dist >= 65536) { m_buf[m_idx] = uint8_t(dist >> 16); m_buf[m_idx + 1] = uint8_t(dist >> 8); m_idx += 2; token = 0x18; }
                else { m_buf[m_idx++] = uint8_t(dist >> 8); token = 0x10; }
            } else token = 0x08;
            m_buf[m_idx++] = uint8_t(dist);
        }
        if (m_len >= m_len_th) { token += m_len_th; m_len_idx += emit_length(&m_len_buf[m_len_idx], m_len - m_len_th); }
        else token += m_len;
        repd[1] = repd[0];
        repd[0] = dist;
        repd_idx = 1;
        const int lit_len = src_idx - anchor;
        if (lit_len == 0) tk_buf[tk_idx++] = uint8_t(token);
        else {
            if (lit_len >= 7) {
                if (lit_len >= (1 << 24)) return -1;
                tk_buf[tk_idx++] = uint8_t((7 << 5) | token);
                dst_idx += emit_length(dst + dst_idx, lit_len - 7);
            } else tk_buf[tk_idx++] = uint8_t((lit_len << 5) | token);
            std::memcpy(dst + dst_idx, src + anchor, size_t(lit_len));
            dst_idx += lit_len;
        }
        anchor = src_idx + best_len;
        src_idx++;
        while (src_idx < anchor) { hashes[lz_hash(le64(src + src_idx), extra)] = src_idx; src_idx++; }
    }
    const int lit_len = count - anchor;
    if (dst_idx + lit_len + tk_idx + m_idx >= count) return -1;
    if (lit_len >= 7) { tk_buf[tk_idx++] = uint8_t(7 << 5); dst_idx += emit_length(dst + dst_idx, lit_len - 7); }
    else tk_buf[tk_idx++] = uint8_t(lit_len << 5);
    std::memcpy(dst + dst_idx, src + anchor, size_t(lit_len));
    dst_idx += lit_len;
    const uint32_t a = uint32_t(dst_idx), b = uint32_t(tk_idx), c = uint32_t(m_idx);
    std::memcpy(dst, &a, 4); std::memcpy(dst + 4, &b, 4); std::memcpy(dst + 8, &c, 4);
    std::memcpy(dst + dst_idx, tk_buf.data(), size_t(tk_idx)); dst_idx += tk_idx;
    std::memcpy(dst + dst_idx, m_buf.data(), size_t(m_idx)); dst_idx += m_idx;
    std::memcpy(dst + dst_idx, m_len_buf.data(), size_t(m_len_idx)); dst_idx += m_len_idx;
    if (dst_idx > count - count / 100) return -1;
    return dst_idx;
}
}
