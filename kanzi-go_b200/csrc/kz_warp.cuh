// Warp-level helpers shared by the entropy kernels.
#pragma once
#include "kz_common.cuh"

namespace kz {

KZ_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ORs the low n bits of value (n in 1..32) into an MSB-first word array in shared memory at bit position bitpos
KZ_D void smem_put_bits(uint32_t* words, uint32_t bitpos, uint32_t value, uint32_t n) {
    if (n < 32) value &= (1u << n) - 1u;
    uint32_t w = bitpos >> 5, o = bitpos & 31;
    if (o + n <= 32) {
        atomicOr(&words[w], value << (32 - o - n));
    } else {
        uint32_t r = o + n - 32;  // bits spilling into the next word
        atomicOr(&words[w], value >> r);
        atomicOr(&words[w + 1], value << (32 - r));
    }
}

KZ_D uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Order-0 histogram of src[0..len) by one warp (internal/Global.go:220-251). hist = 8 x 256 counters in shared memory
// (8 replicas to spread same-address atomics); src must be 16-byte aligned. On return f[j] = count of symbol lane + 32 j.
KZ_D void warp_histogram(const uint8_t* __restrict__ src, uint32_t len, uint32_t* hist, int lane, uint32_t (&f)[8]) {
    for (int i = lane; i < 8 * 256; i += 32) hist[i] = 0;
    __syncwarp();
    uint32_t* my = hist + (lane & 7) * 256;
    const uint32_t n16 = len >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    for (uint32_t i = lane; i < n16; i += 32) {
        uint4 v = __ldg(&src4[i]);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&my[w[k] & 0xFF], 1u);
            atomicAdd(&my[(w[k] >> 8) & 0xFF], 1u);
            atomicAdd(&my[(w[k] >> 16) & 0xFF], 1u);
            atomicAdd(&my[w[k] >> 24], 1u);
        }
    }
    for (uint32_t i = (n16 << 4) + lane; i < len; i += 32) atomicAdd(&my[src[i]], 1u);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t s = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) s += hist[r * 256 + lane + 32 * j];
        f[j] = s;
    }
}

// lr: log range (12 for ANS0). Shared by the order-1 path which calls it per context with lr = 11.
// freq8[j] holds the raw count of symbol (lane + 32 j); on return the normalised frequency. Returns alphabet size.
// Warp-parallel restatement of NormalizeFrequencies (EntropyUtils.go:123-260); all lanes must call.
KZ_D int warp_normalize(uint32_t (&f)[8], uint32_t total, uint32_t lr, int lane) {
    const uint32_t scale = 1u << lr;
    if (total == 0) return 0;
    uint32_t present = 0;  // bit j: symbol lane+32j present
    int asz = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t b = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        asz += __popc(b);
        if (f[j] != 0) present |= 1u << j;
    }
    if (total == scale) return asz;  // :139-148 shortcut, frequencies untouched
    // :155-186 scale
    uint32_t sum_scaled = 0;
    uint32_t best = 0, best_idx = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (f[j] != 0) {
            uint64_t sf = (uint64_t)f[j] * scale;
            uint32_t s = (sf <= total) ? 1u : (uint32_t)((sf + (total >> 1)) / total);
            f[j] = s;
            sum_scaled += s;
            uint32_t idx = (uint32_t)lane + 32u * j;
            if (s > best || (s == best && idx < best_idx)) {  // first maximum in symbol order
                best = s;
                best_idx = idx;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum_scaled += __shfl_xor_sync(0xFFFFFFFFu, sum_scaled, d);
        uint32_t ob = __shfl_xor_sync(0xFFFFFFFFu, best, d);
        uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, best_idx, d);
        if (ob > best || (ob == best && oi < best_idx)) {
            best = ob;
            best_idx = oi;
        }
    }
    if (asz == 0) return 0;
    const int max_lane = (int)(best_idx & 31), max_j = (int)(best_idx >> 5);
    if (asz == 1) {  // :192-195
        if (lane == max_lane) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j == max_j) f[j] = scale;
        }
        return 1;
    }
    if (sum_scaled == scale) return asz;
    int delta = (int)sum_scaled - (int)scale;
    const int err_thr = (int)(best >> 4);
    int fmax = (int)best;  // running value of freqs[idxMax], tracked uniformly by every lane
    const int abs_delta = delta < 0 ? -delta : delta;
    if (abs_delta <= err_thr) {  // :211-215
        fmax -= delta;
    } else {
        int inc;
        if (delta < 0) {  // :217-226
            delta += err_thr;
            fmax += err_thr;
            inc = 1;
            delta = -delta;
        } else {
            delta -= err_thr;
            fmax -= err_thr;
            inc = -1;
        }
        if (lane == max_lane) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j == max_j) f[j] = (uint32_t)fmax;
        }
        // :232-256 up to 5 rounds; in a round every symbol is visited once in alphabet (= symbol) order and the
        // first `delta` symbols whose frequency is > 2 are adjusted -> prefix count over the eligibility mask.
        for (int round = 1; round < 6 && delta > 0; round++) {
            int before = 0, adjustments = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                bool elig = ((present >> j) & 1u) && f[j] > 2;
                uint32_t b = __ballot_sync(0xFFFFFFFFu, elig);
                int rank = before + __popc(b & ((1u << lane) - 1u));
                if (elig && rank < delta) f[j] = (uint32_t)((int)f[j] + inc);
                before += __popc(b);
            }
            adjustments = before < delta ? before : delta;
            delta -= adjustments;
            if (adjustments == 0) break;
        }
        // re-read freqs[idxMax] (it may have been adjusted in the rounds)
        uint32_t cur = 0;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j == max_j) cur = f[j];
        fmax = (int)__shfl_sync(0xFFFFFFFFu, cur, max_lane);
        fmax = fmax - delta;  // :258
        if (fmax < 1) fmax = 1;
    }
    if (lane == max_lane) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j == max_j) f[j] = (uint32_t)fmax;
    }
    return asz;
}


}  // namespace kz
