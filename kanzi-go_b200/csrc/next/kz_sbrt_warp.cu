// NEXT ROUND — not part of the build (compile-checked only: nvcc -c). Warp-resident RANK / MTFT inverse, the CUDA form of the lane model
// tests/host/sbrt_inv_warp_host.cpp (checked against the oracle by tests/test_sbrt_inv_warp_host.py). One warp per block; the 256-entry
// list lives in registers: lane l holds ranks 8 l .. 8 l + 7 as 64-bit entries q << 35 | p << 8 | symbol (blocks < 128 MiB).
// Reference: v2/transform/SBRT.go:177-226, modes :93-111.
#include "../kz_sbrt.cuh"

namespace kz {
namespace sbrtwarp {

KZ_D uint64_t pack(uint64_t q, uint64_t p, uint32_t sym) { return (q << 35) | (p << 8) | sym; }
KZ_D uint64_t ent_q(uint64_t e) { return e >> 35; }
KZ_D uint64_t ent_p(uint64_t e) { return (e >> 8) & ((1ull << 27) - 1); }

// mode: 1 = MTFT, 2 = RANK. Input: ranks (len bytes at src_off); output: symbols at dst_off.
__global__ void __launch_bounds__(32) sbrt_inverse_warp_kernel(const uint8_t* __restrict__ data, const SbrtBlock* __restrict__ blocks, int nblocks, int mode,
                                                                uint8_t* __restrict__ out) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const SbrtBlock blk = blocks[b];
    if (!blk.active) return;
    const uint8_t* src = data + blk.src_off;
    uint8_t* dst = out + blk.dst_off;
    const uint32_t n = blk.len;
    const bool rank_mode = mode == 2;
    uint64_t e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) e[k] = pack(0, 0, (uint32_t)(8 * lane + k));
    const uint32_t base_rank = 8u * (uint32_t)lane;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        // 32 input ranks per round: one byte per lane, handed out by shuffles; the 32 decoded symbols leave as one coalesced store
        const uint32_t mine = i0 + lane < n ? src[i0 + lane] : 0u;
        uint32_t out_sym = 0;
        const uint32_t cnt = min(32u, n - i0);
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t i = i0 + j;
            const uint32_t r = __shfl_sync(0xFFFFFFFFu, mine, (int)j);
            // the owner of rank r selects its register (no dynamic register indexing: a select chain)
            uint64_t sel = e[0];
#pragma unroll
            for (int k = 1; k < 8; k++)
                if ((r & 7u) == (uint32_t)k) sel = e[k];
            const uint64_t cur = __shfl_sync(0xFFFFFFFFu, sel, (int)(r >> 3));
            if ((uint32_t)lane == j) out_sym = (uint32_t)(cur & 0xFF);
            const uint64_t qc = rank_mode ? (((uint64_t)i + ent_p(cur)) >> 1) : (uint64_t)i;
            uint32_t c = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) c += (base_rank + k < r && ent_q(e[k]) > qc) ? 1u : 0u;
            const uint32_t nr = __reduce_add_sync(0xFFFFFFFFu, c);
            // move ranks nr .. r-1 down by one: register shift inside the lane, the lane's old last entry goes up one lane
            const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, e[7], 1);
            const uint64_t fresh = pack(qc, (uint64_t)i, (uint32_t)(cur & 0xFF));
#pragma unroll
            for (int k = 7; k >= 1; k--) {
                const uint32_t rank = base_rank + k;
                if (rank > nr && rank <= r) e[k] = e[k - 1];
            }
            if (base_rank > nr && base_rank <= r) e[0] = up;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (base_rank + k == nr) e[k] = fresh;
        }
        if (i0 + lane < n) dst[i0 + lane] = (uint8_t)out_sym;
    }
}

}  // namespace sbrtwarp
}  // namespace kz
