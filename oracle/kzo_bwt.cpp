// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). BWT + BWTBlockCodec.
// Follows v2/transform/BWT.go:132-208 (Forward/Inverse entry), :211-358 (inverseMergeTPSI semantics),
// v2/transform/DivSufSort.go:179-311 (ComputeBWT output layout + primary indexes),
// v2/transform/BWTBlockCodec.go:78-136, :141-225 (block header).
// The BWT is canonical: any correct suffix sort yields the reference's bytes, so the suffix array is built with
// SA-IS (Nong/Zhang/Chan 2009) instead of restating DivSufSort's 2.7 kLoC. Suffix order = the reference's:
// no sentinel in the data, a suffix that is a proper prefix of another sorts first (implicit smallest terminator).
#include <algorithm>
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {
namespace {

// ---- SA-IS over a generic character accessor. s has n characters, the last one being a unique smallest 0.
struct IntAcc {
    const int32_t* p;
    int32_t operator()(int32_t i) const { return p[i]; }
};
template <class S>
struct Sais {
    const S& s;
    int32_t* SA;
    int32_t n, K;
    std::vector<uint8_t> t;  // 1 = S-type
    std::vector<int32_t> bkt;

    Sais(const S& str, int32_t* sa, int32_t len, int32_t alpha) : s(str), SA(sa), n(len), K(alpha), t(size_t(len)), bkt(size_t(alpha)) {}

    inline bool lms(int32_t i) const { return i > 0 && t[size_t(i)] && !t[size_t(i - 1)]; }

    void buckets(bool end) {
        std::fill(bkt.begin(), bkt.end(), 0);
        for (int32_t i = 0; i < n; i++) bkt[size_t(s(i))]++;
        int32_t sum = 0;
        for (int32_t c = 0; c < K; c++) {
            sum += bkt[size_t(c)];
            bkt[size_t(c)] = end ? sum : sum - bkt[size_t(c)];
        }
    }
    void induce_l() {
        buckets(false);
        for (int32_t i = 0; i < n; i++) {
            int32_t j = SA[i] - 1;
            if (SA[i] > 0 && !t[size_t(j)]) SA[bkt[size_t(s(j))]++] = j;
        }
    }
    void induce_s() {
        buckets(true);
        for (int32_t i = n - 1; i >= 0; i--) {
            int32_t j = SA[i] - 1;
            if (SA[i] > 0 && t[size_t(j)]) SA[--bkt[size_t(s(j))]] = j;
        }
    }

    void run() {
        if (n == 1) { SA[0] = 0; return; }
        t[size_t(n - 1)] = 1;
        t[size_t(n - 2)] = 0;
        for (int32_t i = n - 3; i >= 0; i--) {
            auto a = s(i), b = s(i + 1);
            t[size_t(i)] = (a < b || (a == b && t[size_t(i + 1)])) ? 1 : 0;
        }
        // stage 1: sort the LMS substrings
        buckets(true);
        std::fill(SA, SA + n, -1);
        for (int32_t i = 1; i < n; i++)
            if (lms(i)) SA[--bkt[size_t(s(i))]] = i;
        induce_l();
        induce_s();
        int32_t n1 = 0;
        for (int32_t i = 0; i < n; i++)
            if (lms(SA[i])) SA[n1++] = SA[i];
        std::fill(SA + n1, SA + n, -1);
        int32_t name = 0, prev = -1;
        for (int32_t i = 0; i < n1; i++) {
            int32_t pos = SA[i];
            bool diff = false;
            for (int32_t d = 0; d < n; d++) {
                if (prev == -1 || s(pos + d) != s(prev + d) || t[size_t(pos + d)] != t[size_t(prev + d)]) {
                    diff = true;
                    break;
                }
                if (d > 0 && (lms(pos + d) || lms(prev + d))) break;
            }
            if (diff) {
                name++;
                prev = pos;
            }
            SA[n1 + (pos >> 1)] = name - 1;
        }
        for (int32_t i = n - 1, j = n - 1; i >= n1; i--)
            if (SA[i] >= 0) SA[j--] = SA[i];
        // stage 2: reduced problem
        int32_t* SA1 = SA;
        int32_t* s1 = SA + n - n1;
        if (name < n1) {
            IntAcc acc{s1};
            Sais<IntAcc> sub(acc, SA1, n1, name);
            sub.run();
        } else {
            for (int32_t i = 0; i < n1; i++) SA1[s1[i]] = i;
        }
        // stage 3: induce the final order
        buckets(true);
        for (int32_t i = 1, j = 0; i < n; i++)
            if (lms(i)) s1[j++] = i;
        for (int32_t i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
        std::fill(SA + n1, SA + n, -1);
        for (int32_t i = n1 - 1; i >= 0; i--) {
            int32_t j = SA[i];
            SA[i] = -1;
            SA[--bkt[size_t(s(j))]] = j;
        }
        induce_l();
        induce_s();
    }
};

// suffix array of src[0..n) under the reference's order; sa has n entries
void suffix_array(const uint8_t* src, int32_t n, std::vector<int32_t>& sa) {
    std::vector<int32_t> work(size_t(n) + 1);
    auto acc = [src, n](int32_t i) -> int32_t { return i == n ? 0 : int32_t(src[i]) + 1; };
    Sais<decltype(acc)> sais(acc, work.data(), n + 1, 257);
    sais.run();
    sa.assign(work.begin() + 1, work.end());  // work[0] is the terminator
}
}  // namespace

int bwt_chunks(size_t n) { return n < 256 ? 1 : 8; }  // BWT.go:631-637

// DivSufSort.ComputeBWT layout (:179-198): dst[0] = src[n-1]; then src[SA[r]-1] for every rank r in order, skipping the
// rank of suffix 0. primary[k] = rank(k*step)+1, step = ceil(n/chunks) (:200-206, :221-223, :273-275, :309).
void bwt_forward_raw(const uint8_t* src, uint8_t* dst, size_t n, uint32_t* primary) {
    for (int k = 0; k < 8; k++) primary[k] = 0;
    if (n == 0) return;
    if (n == 1) {  // BWT.go:151-154 (primary indexes untouched)
        dst[0] = src[0];
        return;
    }
    std::vector<int32_t> sa;
    suffix_array(src, int32_t(n), sa);
    int chunks = bwt_chunks(n);
    size_t step = n / size_t(chunks);
    if (step * size_t(chunks) != n) step++;
    dst[0] = src[n - 1];
    size_t o = 1;
    for (size_t r = 0; r < n; r++) {
        size_t s = size_t(sa[r]);
        if (s % step == 0) primary[s / step] = uint32_t(r + 1);
        if (s != 0) dst[o++] = src[s - 1];
    }
}

// Same mapping as inverseMergeTPSI (BWT.go:231-258): stable counting sort of the BWT bytes gives, for every rank,
// (first byte of that suffix, rank of the next suffix). The walk starts at rank primary[0]-1. The chunked
// variants (:259-356, biPSIv2 :361-628) produce the same bytes for a valid stream.
bool bwt_inverse_raw(const uint8_t* src, uint8_t* dst, size_t n, const uint32_t* primary) {
    if (n == 0) return true;
    if (n == 1) {
        dst[0] = src[0];
        return true;
    }
    size_t pidx = primary[0];
    if (pidx == 0 || pidx > n) return false;
    if (bwt_chunks(n) == 8)
        for (int k = 0; k < 8; k++)
            if (primary[k] == 0 || primary[k] > n) return false;
    std::vector<uint32_t> next(n);
    std::vector<uint8_t> first(n);
    size_t buckets[256];
    std::memset(buckets, 0, sizeof(buckets));
    for (size_t i = 0; i < n; i++) buckets[src[i]]++;
    size_t sum = 0;
    for (int c = 0; c < 256; c++) {
        size_t tmp = buckets[c];
        buckets[c] = sum;
        sum += tmp;
    }
    const uint32_t END = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        uint8_t v = src[i];
        size_t slot = buckets[v]++;
        first[slot] = v;
        next[slot] = i == 0 ? END : uint32_t(i < pidx ? i - 1 : i);
    }
    uint32_t t = uint32_t(pidx - 1);
    for (size_t i = 0; i < n; i++) {
        if (t == END || t >= n) return false;
        dst[i] = first[t];
        t = next[t];
    }
    return true;
}

// BWTBlockCodec.Forward :78-136
bool bwt_block_forward(Ctx&, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if (n == 0) { *out_n = 0; return true; }
    if (cap < n + 33) return false;
    uint32_t log_bs = log2_floor(uint32_t(n));
    if (n & (n - 1)) log_bs++;
    int pidx_size = int(log_bs + 7) >> 3;
    if (pidx_size <= 0 || pidx_size >= 5) return false;
    int chunks = bwt_chunks(n);
    uint32_t log_chunks = log2_floor(uint32_t(chunks));
    size_t header = size_t(chunks * pidx_size + 1);
    uint32_t primary[8];
    bwt_forward_raw(src, dst + header, n, primary);
    dst[0] = uint8_t((log_chunks << 2) | uint32_t(pidx_size - 1));
    size_t idx = 1;
    for (int i = 0; i < chunks; i++) {
        uint32_t p = primary[i] - 1;
        for (int shift = (pidx_size - 1) << 3; shift >= 0; shift -= 8) dst[idx++] = uint8_t(p >> shift);
    }
    *out_n = n + header;
    return true;
}

// BWTBlockCodec.Inverse :141-225 (bitstream version 6 branch)
bool bwt_block_inverse(Ctx&, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if (n == 0) { *out_n = 0; return true; }
    if (n == 1) return false;
    uint8_t mode = src[0];
    unsigned log_chunks = (mode >> 2) & 7;
    int pidx_size = int(mode & 3) + 1;
    int chunks = 1 << log_chunks;
    size_t header = size_t(chunks * pidx_size + 1);
    if (n < header) return false;
    size_t bs = n - header;
    if (chunks != bwt_chunks(bs)) return false;
    if (chunks > 8) return false;
    uint32_t primary[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t idx = 1;
    for (int i = 0; i < chunks; i++) {
        uint32_t p = 0;
        for (int k = 0; k < pidx_size; k++) p = (p << 8) | src[idx++];
        primary[i] = p + 1;
    }
    if (bs > cap) return false;
    if (!bwt_inverse_raw(src + header, dst, bs, primary)) return false;
    *out_n = bs;
    return true;
}


// ---- BWTS (bijective BWT, transform/BWTS.go). ComputeSuffixArray (DivSufSort.go:161-177) is any correct suffix sort.
namespace {
// BWTS.go:172-209 moveLyndonWordHead
int32_t move_lyndon_word_head(std::vector<int32_t>& sa, std::vector<int32_t>& isa, const uint8_t* data, int32_t count, int32_t start, int32_t size,
                              int32_t rank) {
    const int32_t end = start + size;
    while (rank + 1 < count) {
        const int32_t next_start0 = sa[size_t(rank) + 1];
        if (next_start0 <= end) break;
        int32_t next_start = next_start0;
        int32_t k = 0;
        while (k < size && next_start < count && data[start + k] == data[next_start]) {
            k++;
            next_start++;
        }
        if (k == size && rank < isa[size_t(next_start)]) break;
        if (k < size && next_start < count && data[start + k] < data[next_start]) break;
        sa[size_t(rank)] = next_start0;
        isa[size_t(next_start0)] = rank;
        rank++;
    }
    sa[size_t(rank)] = start;
    isa[size_t(start)] = rank;
    return rank;
}
}  // namespace

// BWTS.Forward :58-170
bool bwts_forward(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (cap < n) return false;
    if (n > (size_t(1) << 30)) return false;
    if (n < 2) {
        dst[0] = src[0];
        *out_n = n;
        return true;
    }
    const int32_t count = int32_t(n);
    std::vector<int32_t> sa, isa(n);
    suffix_array(src, count, sa);
    for (int32_t i = 0; i < count; i++) isa[size_t(sa[size_t(i)])] = i;
    int32_t mn = isa[0];
    int32_t idx_min = 0;
    for (int32_t i = 1; i < count && mn > 0; i++) {
        if (isa[size_t(i)] >= mn) continue;
        int32_t ref_rank = move_lyndon_word_head(sa, isa, src, count, idx_min, i - idx_min, mn);
        for (int32_t j = i - 1; j > idx_min; j--) {  // iterate through the new lyndon word from end to start
            int32_t test_rank = isa[size_t(j)];
            const int32_t start_rank = test_rank;
            while (test_rank < count - 1) {
                const int32_t next_rank_start = sa[size_t(test_rank) + 1];
                if (j > next_rank_start || src[j] != src[next_rank_start] || ref_rank < isa[size_t(next_rank_start) + 1]) break;
                sa[size_t(test_rank)] = next_rank_start;
                isa[size_t(next_rank_start)] = test_rank;
                test_rank++;
            }
            sa[size_t(test_rank)] = j;
            isa[size_t(j)] = test_rank;
            ref_rank = test_rank;
            if (start_rank == test_rank) break;
        }
        mn = isa[size_t(i)];
        idx_min = i;
    }
    mn = count;
    for (int32_t i = 0; i < count; i++) {
        if (isa[size_t(i)] >= mn) {
            dst[isa[size_t(i)]] = src[i - 1];
            continue;
        }
        if (mn < count) dst[mn] = src[i - 1];
        mn = isa[size_t(i)];
    }
    dst[0] = src[count - 1];
    *out_n = n;
    return true;
}

// BWTS.Inverse :211-288
bool bwts_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return true;
    if (n > (size_t(1) << 30) || n > cap) return false;
    if (n < 2) {
        dst[0] = src[0];
        *out_n = n;
        return true;
    }
    const int32_t count = int32_t(n);
    std::vector<int32_t> lf(n);
    int32_t buckets[256] = {0};
    for (int32_t i = 0; i < count; i++) buckets[src[i]]++;
    int32_t sum = 0;
    for (int i = 0; i < 256; i++) {
        sum += buckets[i];
        buckets[i] = sum - buckets[i];
    }
    for (int32_t i = 0; i < count; i++) lf[size_t(i)] = buckets[src[i]]++;
    for (int32_t i = 0, j = count - 1; j >= 0; i++) {
        if (lf[size_t(i)] < 0) continue;
        int32_t p = i;
        for (;;) {
            dst[j] = src[p];
            j--;
            const int32_t t = lf[size_t(p)];
            lf[size_t(p)] = -1;
            p = t;
            if (lf[size_t(p)] < 0) break;
        }
    }
    *out_n = n;
    return true;
}

}  // namespace kzo
