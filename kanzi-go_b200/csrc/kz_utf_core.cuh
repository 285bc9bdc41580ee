// UTF transform (UTFCodec): code points of a UTF-8 block are replaced by 1 or 2 byte ranks ordered by frequency. The per-block logic,
// written once for host and device (same arrangement as kz_text_core.cuh: kernels in kz_utf.cu, CPU check in
// tests/test_utf_core_host.py). Reference: v2/transform/UTFCodec.go — Forward :87-262, Inverse :267-383, validateUTF :393-515,
// packUTF :517-545, unpackUTF1 :577-610 (bitstream >= 4), MaxEncodedLen :386-388.
#pragma once
#include <stdint.h>

#ifndef KZ_UTF_HD
#ifdef __CUDACC__
#define KZ_UTF_HD __host__ __device__ __forceinline__
#else
#define KZ_UTF_HD inline
#endif
#endif

namespace kz {
namespace utfc {

static const int MIN_BLOCKSIZE = 1024, MAX_SYMBOLS = 32768;
static const uint32_t ALIAS_MAP_SIZE = 1u << 22;

// _UTF_SIZES :31-48: bytes of the sequence a leading byte introduces, 0 = not a leading byte
KZ_UTF_HD int utf_size(uint8_t b) {
    if (b < 0x80) return 1;
    if (b < 0xC2) return 0;
    if (b < 0xE0) return 2;
    if (b < 0xF0) return 3;
    if (b < 0xF5) return 4;
    return 0;
}
KZ_UTF_HD bool forbidden_byte(uint8_t b) { return b == 0xC0 || b == 0xC1 || b >= 0xF5; }

KZ_UTF_HD int pack_utf(const uint8_t* in, uint32_t* out) {  // :517-545
    const int s = utf_size(in[0]);
    switch (s) {
        case 1: *out = in[0]; break;
        case 2: *out = (1u << 19) | ((uint32_t)in[0] << 8) | in[1]; break;
        case 3: *out = (2u << 19) | (((uint32_t)in[0] & 0x0F) << 12) | (((uint32_t)in[1] & 0x3F) << 6) | ((uint32_t)in[2] & 0x3F); break;
        case 4:
            *out = (4u << 19) | (((uint32_t)in[0] & 0x07) << 18) | (((uint32_t)in[1] & 0x3F) << 12) | (((uint32_t)in[2] & 0x3F) << 6) | ((uint32_t)in[3] & 0x3F);
            break;
        default: *out = 0; break;
    }
    return s;
}
KZ_UTF_HD int unpack_utf1(uint32_t in, uint8_t* out) {  // :577-610 (out[4]; bytes past the length are zero like the reference's zeroed array)
    out[0] = out[1] = out[2] = out[3] = 0;
    const uint32_t sz = in >> 19;
    if (sz == 0) {
        out[0] = (uint8_t)in;
        return 1;
    }
    if (sz == 1) {
        out[0] = (uint8_t)(in >> 8);
        out[1] = (uint8_t)in;
        return 2;
    }
    if (sz == 2) {
        out[0] = (uint8_t)(((in >> 12) & 0x0F) | 0xE0);
        out[1] = (uint8_t)(((in >> 6) & 0x3F) | 0x80);
        out[2] = (uint8_t)((in & 0x3F) | 0x80);
        return 3;
    }
    if (sz >= 4 && sz <= 7) {
        out[0] = (uint8_t)(((in >> 18) & 0x07) | 0xF0);
        out[1] = (uint8_t)(((in >> 12) & 0x3F) | 0x80);
        out[2] = (uint8_t)(((in >> 6) & 0x3F) | 0x80);
        out[3] = (uint8_t)((in & 0x3F) | 0x80);
        return 4;
    }
    return 0;
}

// start of the regular processing (:117-127): 3 after a byte order mark at bytes 1..3 (sic), else the first <= 4 non-leading bytes are skipped
KZ_UTF_HD int utf_start(const uint8_t* src) {
    if (src[1] == 0xEF && src[2] == 0xBB && src[3] == 0xBF) return 3;
    int start = 0;
    while (start < 4 && utf_size(src[start]) == 0) start++;
    return start;
}

// validateUTF :393-515 over block = src[start : count - 4], given its byte histogram, its (previous, byte) histogram (previous = 0 for
// the first byte) and the number of forbidden bytes (C0, C1, F5..FF) among the first checked_len bytes, where checked_len is what the
// reference's periodic 1-byte check covers: everything when the length is not a multiple of 4, else up to the last multiple of 4096 + 4
KZ_UTF_HD uint32_t validate_checked_len(uint32_t len) {
    const uint32_t end4 = len & ~3u;
    if (end4 != len) return len;
    return end4 > 0 ? ((end4 - 4) & ~0xFFFu) + 4 : 0;
}
KZ_UTF_HD bool validate_utf(const uint32_t* freqs0, const uint32_t* freqs1, uint32_t len, uint32_t forbidden_in_checked) {
    if (forbidden_in_checked != 0) return false;
    uint64_t sum = 0, sum2 = 0;
    for (int i = 0; i < 256; i++) {
        if (i < 0xA0 || i > 0xBF) sum += freqs1[(0xE0 << 8) | i];
        if (i < 0x80 || i > 0x9F) sum += freqs1[(0xED << 8) | i];
        if (i < 0x90 || i > 0xBF) sum += freqs1[(0xF0 << 8) | i];
        if (i < 0x80 || i > 0x8F) sum += freqs1[(0xF4 << 8) | i];
        if (i < 0x80 || i > 0xBF) {
            for (int j = 0xC2; j <= 0xDF; j++) sum += freqs1[(j << 8) | i];
            for (int j = 0xE1; j <= 0xEC; j++) sum += freqs1[(j << 8) | i];
            sum += (uint64_t)freqs1[(0xF1 << 8) | i] + freqs1[(0xF2 << 8) | i] + freqs1[(0xF3 << 8) | i];
            sum += freqs1[(0xEE << 8) | i];
            sum += freqs1[(0xEF << 8) | i];
        } else {
            sum2 += freqs0[i];
        }
        if (sum != 0) return false;
    }
    return sum2 >= (uint64_t)(len / 8);
}

// in-place heap sort of 64-bit keys, ascending (slices.SortStableFunc over a total order = any sort)
KZ_UTF_HD void sort_keys(uint64_t* a, int n) {
    for (int root0 = n / 2 - 1; root0 >= 0; root0--) {
        int root = root0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            const uint64_t t = a[root];
            a[root] = a[child];
            a[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        const uint64_t t = a[0];
        a[0] = a[end];
        a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            const uint64_t u = a[root];
            a[root] = a[child];
            a[child] = u;
            root = child;
        }
    }
}

// Forward :129-262 after the validation. alias_map: ALIAS_MAP_SIZE int32 zeroed; keys: MAX_SYMBOLS uint64 scratch.
// Returns the encoded length, or -1 when the reference returns an error (skip).
KZ_UTF_HD int64_t forward_walk(const uint8_t* src, int count, int start, uint8_t* dst, int32_t* alias_map, uint64_t* keys) {
    int n = 0;
    for (int i = start; i < count - 4;) {  // :146-170
        uint32_t val;
        const int s = pack_utf(src + i, &val);
        bool res = s != 0;
        res = res && (s != 3 || (src[i + 2] & 0xC0) == 0x80);
        res = res && (s != 4 || (((((uint32_t)src[i + 2]) << 8) | src[i + 3]) & 0xC0C0u) == 0x8080u);
        if (alias_map[val] == 0) {
            if (n < MAX_SYMBOLS) keys[n] = val;
            n++;
            res = res && n < MAX_SYMBOLS;
        }
        if (!res) return -1;
        alias_map[val]++;
        i += s;
    }
    if (n == 0) return -1;
    const int max_target = count - count / 10;
    if (3 * n + 6 >= max_target) return -1;
    for (int i = 0; i < n; i++) keys[i] = ((uint64_t)(uint32_t)alias_map[(uint32_t)keys[i]] << 22) | (keys[i] & 0x3FFFFFu);  // (freq, symbol)
    sort_keys(keys, n);  // increasing frequency, then increasing symbol (:190-197)
    int64_t dst_idx = 2;
    dst[dst_idx++] = (uint8_t)(n >> 8);
    dst[dst_idx++] = (uint8_t)n;
    int64_t estimate = dst_idx + 6;
    for (int i = 0; i < n; i++) {
        const uint64_t k = keys[n - 1 - i];
        const int32_t s = (int32_t)(k & 0x3FFFFFu);
        const int64_t freq = (int64_t)(k >> 22);
        dst[dst_idx] = (uint8_t)(s >> 16);
        dst[dst_idx + 1] = (uint8_t)(s >> 8);
        dst[dst_idx + 2] = (uint8_t)s;
        dst_idx += 3;
        if (i < 128) {
            estimate += freq;
            alias_map[s] = i;
        } else {
            estimate += 2 * freq;
            alias_map[s] = 0x10080 | ((i << 1) & 0xFF00) | (i & 0x7F);
        }
    }
    if (estimate >= max_target) return -1;
    for (int i = 0; i < start; i++) dst[dst_idx++] = src[i];
    int src_idx = start;
    while (src_idx < count - 4) {
        uint32_t val;
        src_idx += pack_utf(src + src_idx, &val);
        const int32_t alias = alias_map[val];
        dst[dst_idx++] = (uint8_t)alias;
        dst[dst_idx] = (uint8_t)(alias >> 8);
        dst_idx += alias >> 16;
    }
    dst[0] = (uint8_t)start;
    dst[1] = (uint8_t)(src_idx - (count - 4));
    while (src_idx < count) dst[dst_idx++] = src[src_idx++];
    if (dst_idx >= max_target) return -1;
    return dst_idx;
}

// Inverse :267-383 (bitstream version >= 4). sym: MAX_SYMBOLS scratch entries {4 bytes, length}. Returns the length or -1.
struct Sym {
    uint8_t value[4];
    uint32_t length;
};
KZ_UTF_HD int64_t inverse_walk(const uint8_t* src, int count, uint8_t* dst, int64_t cap, Sym* m) {
    if (count < 4) return -1;
    const int start = src[0] & 3, adjust = src[1] & 3;
    const int n = ((int)src[2] << 8) + src[3];
    if (n == 0 || n >= MAX_SYMBOLS || 4 + 3 * n > count) return -1;
    int src_idx = 4;
    for (int i = 0; i < n; i++) {
        const uint32_t s = ((uint32_t)src[src_idx] << 16) | ((uint32_t)src[src_idx + 1] << 8) | src[src_idx + 2];
        const int sl = unpack_utf1(s, m[i].value);
        if (sl == 0) return -1;
        m[i].length = (uint32_t)sl;
        src_idx += 3;
    }
    const int src_end = count - 4 + adjust;
    int64_t dst_idx = 0;
    const int64_t dst_end = cap - 4;
    if (dst_end < 0) return -1;
    if (src_end < src_idx || src_end > count || src_idx + start > count) return -1;
    for (int i = 0; i < start; i++) dst[dst_idx++] = src[src_idx++];
    while (src_idx < src_end && dst_idx < dst_end) {
        int alias = src[src_idx++];
        if (alias >= 128) {
            if (src_idx >= src_end) return -1;
            alias = ((int)src[src_idx] << 7) + (alias & 0x7F);
            src_idx++;
        }
        if (alias >= n) {  // entries past the map are empty symbols in the reference: four zero bytes, length 0
            dst[dst_idx] = dst[dst_idx + 1] = dst[dst_idx + 2] = dst[dst_idx + 3] = 0;
            continue;
        }
        const Sym& s = m[alias];
        dst[dst_idx] = s.value[0];
        dst[dst_idx + 1] = s.value[1];
        dst[dst_idx + 2] = s.value[2];
        dst[dst_idx + 3] = s.value[3];
        dst_idx += s.length;
    }
    if (src_idx < src_end || dst_idx > cap - count + src_end) return -1;
    for (int i = src_end; i < count; i++) dst[dst_idx++] = src[src_idx++];
    return dst_idx;
}

}  // namespace utfc
}  // namespace kz
