// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Sort-by-rank transforms (RANK / MTFT) and zero run length transform.
// Follows v2/transform/SBRT.go:127-226 and v2/transform/ZRLT.go:58-225 line by line.
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {

// SBRT.go:30-38, :93-111: mode -> (mask1, mask2, shift)
static void sbrt_params(int mode, int64_t& m1, int64_t& m2, unsigned& s) {
    m1 = mode == 3 ? 0 : -1;
    m2 = mode == 1 ? 0 : -1;
    s = mode == 2 ? 1 : 0;
}

// SBRT.Forward :127-172
bool sbrt_forward(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (cap < n + 33) return false;  // MaxEncodedLen :229-231
    uint8_t s2r[256], r2s[256];
    for (int i = 0; i < 256; i++) s2r[i] = r2s[i] = uint8_t(i);
    int64_t m1, m2;
    unsigned s;
    sbrt_params(mode, m1, m2, s);
    int64_t p[256], q[256];
    std::memset(p, 0, sizeof(p));
    std::memset(q, 0, sizeof(q));
    for (size_t i = 0; i < n; i++) {
        uint8_t c = src[i];
        uint8_t r = s2r[c];
        dst[i] = r;
        int64_t qc = ((int64_t(i) & m1) + (p[c] & m2)) >> s;
        p[c] = int64_t(i);
        q[c] = qc;
        while (r > 0 && q[r2s[r - 1]] <= qc) {
            uint8_t t = r2s[r - 1];
            r2s[r] = t;
            s2r[t] = r;
            r--;
        }
        r2s[r] = c;
        s2r[c] = r;
    }
    *out_n = n;
    return true;
}

// SBRT.Inverse :177-226
bool sbrt_inverse(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (n > cap) return false;
    uint8_t r2s[256];
    for (int i = 0; i < 256; i++) r2s[i] = uint8_t(i);
    int64_t m1, m2;
    unsigned s;
    sbrt_params(mode, m1, m2, s);
    int64_t p[256], q[256];
    std::memset(p, 0, sizeof(p));
    std::memset(q, 0, sizeof(q));
    for (size_t i = 0; i < n; i++) {
        uint8_t r = src[i];
        uint8_t c = r2s[r];
        dst[i] = c;
        int64_t qc = ((int64_t(i) & m1) + (p[c] & m2)) >> s;
        p[c] = int64_t(i);
        q[c] = qc;
        while (r > 0 && q[r2s[r - 1]] <= qc) {
            r2s[r] = r2s[r - 1];
            r--;
        }
        r2s[r] = c;
    }
    *out_n = n;
    return true;
}

// ZRLT.Forward :58-141
bool zrlt_forward(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (cap < n) return false;
    const uint64_t src_end = n, dst_end = n;  // do not expand
    uint64_t si = 0, di = 0;
    bool res = true;
    while (si < src_end) {
        if (src[si] == 0) {
            const uint64_t run_start = si - 1;  // wraps for si == 0 exactly like the uint arithmetic of the reference
            si++;
            while (si + 1 < src_end && (src[si] | src[si + 1]) == 0) si += 2;
            while (si < src_end && src[si] == 0) si++;
            const uint64_t run_length = si - run_start;
            unsigned lg = log2_floor(uint32_t(run_length));
            if (di >= dst_end - uint64_t(lg)) {
                res = false;
                break;
            }
            while (lg > 0) {
                lg--;
                dst[di++] = uint8_t((run_length >> lg) & 1);
            }
            continue;
        }
        if (src[si] >= 0xFE) {
            if (di >= dst_end - 1) {
                res = false;
                break;
            }
            dst[di++] = 0xFF;
            dst[di] = uint8_t(src[si] - 0xFE);
        } else {
            if (di >= dst_end) {
                res = false;
                break;
            }
            dst[di] = uint8_t(src[si] + 1);
        }
        si++;
        di++;
    }
    *out_n = size_t(di);
    return si == src_end && res;
}

// ZRLT.Inverse :146-225
bool zrlt_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0 || cap == 0) return true;
    const uint64_t src_end = n, dst_end = cap;
    uint64_t si = 0, di = 0, run_length = 0;
    bool ok = true;
    for (;;) {
        if (src[si] <= 1) {
            run_length = 1;
            bool ended = false;
            while (src[si] <= 1) {
                run_length += run_length + src[si];
                si++;
                if (si >= src_end) {
                    ended = true;
                    break;
                }
            }
            if (ended) break;  // goto End
            run_length--;
            if (run_length >= dst_end - di) break;
            while (run_length > 0) {
                run_length--;
                dst[di++] = 0;
            }
        }
        if (src[si] == 0xFF) {
            si++;
            if (si >= src_end) break;
            dst[di] = uint8_t(0xFE + src[si]);
        } else {
            dst[di] = uint8_t(src[si] - 1);
        }
        si++;
        di++;
        if (si >= src_end || di >= dst_end) break;
    }
    if (run_length > 0) {  // End:
        run_length--;
        if (run_length > dst_end - di) {
            ok = false;
        } else {
            while (run_length > 0) {
                run_length--;
                dst[di++] = 0;
            }
        }
    }
    if (si < src_end) ok = false;
    *out_n = size_t(di);
    return ok;
}

}  // namespace kzo
