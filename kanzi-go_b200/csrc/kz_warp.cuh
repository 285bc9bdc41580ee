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

}  // namespace kz
