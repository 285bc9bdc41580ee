// MM transform (FSDCodec: fixed step delta) on the GPU.
// Reference: v2/transform/FSDCodec.go — Forward :135-312 samples three sub-blocks, compares the order-0 entropy of the bytes with that
// of the xor against the byte 1, 2, 3, 4, 8 or 16 positions back (internal.ComputeFirstOrderEntropy1024, internal/Global.go:196-214),
// and when a distance wins codes the block as zig-zag deltas (escape 0xFF + xor for deltas outside [-127, 127]) or plain xors;
// Inverse :317-404.
//
// Forward kernels (whole batch, no host synchronisation): sample (7 histograms) -> plan (checks, entropies, distance, coding mode,
// header) -> tiles (bytes produced per 4096 positions) -> scan (tile offsets, completion test) -> emit -> post_hist -> final
// (the "does it still look better" test of :296-310). Inverse: one thread per block walks the tokens (a block's bytes form `dist`
// dependent chains and escapes change the operation, so there is little to share inside a block; blocks run side by side).
#include <algorithm>
#include <cmath>

#include "kz_alias.cuh"
#include "kz_fsd.cuh"

namespace kz {

namespace {

const int DT_UNDEFINED = 0, DT_MULTIMEDIA = 2, DT_BIN = 7;
const uint32_t SAMPLE_SLICES = 16;

struct FsdInfo {
    uint32_t go;       // 1 while the block is still a candidate
    uint32_t mode;     // 0 delta, 1 xor
    uint32_t dist;
    int32_t ent0;
    uint32_t out_len;
    uint32_t pad[3];
};

struct Ws {
    uint32_t* log2tab;  // [257] LOG2_4096
    uint32_t* hist7;    // [nblocks][7][256]
    uint32_t* hpost;    // [nblocks][256]
    FsdInfo* info;      // [nblocks]
    uint32_t* tile_sum; // [nblocks][ntiles]
    uint32_t* tile_off; // [nblocks][ntiles]
    uint32_t ntiles;
};
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
Ws carve(uint8_t* ws, uint32_t nblocks, uint32_t max_len) {
    Ws w;
    w.ntiles = (max_len + FSD_TILE - 1) / FSD_TILE + 1;
    uint8_t* p = ws;
    w.log2tab = reinterpret_cast<uint32_t*>(p);
    p += align256(257 * 4);
    w.hist7 = reinterpret_cast<uint32_t*>(p);
    p += align256((size_t)nblocks * 7 * 256 * 4);
    w.hpost = reinterpret_cast<uint32_t*>(p);
    p += align256((size_t)nblocks * 256 * 4);
    w.info = reinterpret_cast<FsdInfo*>(p);
    p += align256((size_t)nblocks * sizeof(FsdInfo));
    w.tile_sum = reinterpret_cast<uint32_t*>(p);
    p += align256((size_t)nblocks * w.ntiles * 4);
    w.tile_off = reinterpret_cast<uint32_t*>(p);
    return w;
}

// internal/Global.go:60-93 LOG2_4096[i] = round(4096 * log2(i)) (entries 0 and 1 are 0): a data table of the reference, regenerated from
// its definition (checked entry by entry against the reference's table when this was written)
const uint32_t* host_log2_4096() {
    static uint32_t tab[257];
    static bool init = false;
    if (!init) {
        tab[0] = 0;
        for (int k = 1; k <= 256; k++) tab[k] = (uint32_t)(4096.0 * std::log2((double)k) + 0.5);
        init = true;
    }
    return tab;
}

// internal/Global.go:175-191 Log2ScaledBy1024
KZ_D uint32_t log2_scaled_1024(uint32_t x, const uint32_t* T) {
    if (x < 256) return (T[x] + 2) >> 2;
    const uint32_t lg = 31u - (uint32_t)__clz((int)x);
    if ((x & (x - 1)) == 0) return lg << 10;
    return ((lg - 7) * 1024) + ((T[x >> (lg - 7)] + 2) >> 2);
}
// internal/Global.go:196-214 ComputeFirstOrderEntropy1024
KZ_D int entropy_1024(uint32_t block_len, const uint32_t* histo, const uint32_t* T) {
    if (block_len == 0) return 0;
    unsigned long long sum = 0;
    const uint32_t ll = log2_scaled_1024(block_len, T);
    for (int i = 0; i < 256; i++) {
        if (histo[i] == 0) continue;
        const uint32_t l = log2_scaled_1024(histo[i], T);
        sum += ((unsigned long long)histo[i] * (unsigned long long)(ll - l)) >> 3;
    }
    return (int)(sum / (unsigned long long)block_len);
}

// internal/Magic.go:73-112 GetMagicType reduced to what FSDCodec.Forward :170-187 asks: no magic, or BMP / RIFF / PBM / PGM / PPM
KZ_D bool magic_allows_fsd(const uint8_t* p, uint32_t n) {
    if (n < 4) return true;
    const uint32_t key = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    if ((key & ~0x0Fu) == 0xFFD8FFE0u) return false;                       // JPEG
    if ((key >> 8) == 0x425A68u || (key >> 8) == 0x494433u) return false;  // bzip2, ID3
    const uint32_t k32[18] = {0x47494638u, 0x25504446u, 0x504B0304u, 0x377ABCAFu, 0x89504E47u, 0x7F454C46u, 0xFEEDFACEu, 0xCEFAEDFEu, 0xFEEDFACFu,
                              0xCFFAEDFEu, 0x28B52FFDu, 0x81CFB2CEu, 0x4D534346u, 0x52494646u, 0x664C6143u, 0xFD377A58u, 0x4B414E5Au, 0x52617221u};
    for (int i = 0; i < 18; i++)
        if (key == k32[i]) return key == 0x52494646u;  // RIFF is a candidate, the others are not
    const uint32_t k16 = key >> 16;
    if (k16 == 0x1F8Bu || k16 == 0x4D5Au) return false;  // gzip, MZ
    return true;                                          // BMP, binary PBM / PGM / PPM, or no magic at all
}

KZ_D uint32_t fsd_dist(uint32_t idx) { return idx < 5 ? idx : (idx == 5 ? 8u : 16u); }  // distances {0, 1, 2, 3, 4, 8, 16}

__global__ void __launch_bounds__(256) fsd_sample_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, uint32_t* __restrict__ hist7) {
    __shared__ uint32_t h[7 * 256];
    const int b = blockIdx.y;
    const FsdBlock blk = blocks[b];
    if (blk.len < 1024) return;
    for (int i = threadIdx.x; i < 7 * 256; i += 256) h[i] = 0;
    __syncthreads();
    const uint32_t count10 = blk.len / 10, count5 = 2 * count10;
    const uint8_t* src = in + blk.src_off;
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < count10; k += gridDim.x * 256) {
        const uint32_t i = count10 + k;
#pragma unroll
        for (uint32_t sb = 0; sb < 3; sb++) {
            const uint8_t* p = src + 2 * sb * count5 + i;
            const uint32_t v = p[0];
            atomicAdd(&h[v], 1u);
            atomicAdd(&h[1 * 256 + (v ^ *(p - 1))], 1u);
            atomicAdd(&h[2 * 256 + (v ^ *(p - 2))], 1u);
            atomicAdd(&h[3 * 256 + (v ^ *(p - 3))], 1u);
            atomicAdd(&h[4 * 256 + (v ^ *(p - 4))], 1u);
            atomicAdd(&h[5 * 256 + (v ^ *(p - 8))], 1u);
            atomicAdd(&h[6 * 256 + (v ^ *(p - 16))], 1u);
        }
    }
    __syncthreads();
    uint32_t* g = hist7 + (size_t)b * 7 * 256;
    for (int i = threadIdx.x; i < 7 * 256; i += 256)
        if (h[i]) atomicAdd(&g[i], h[i]);
}

__global__ void __launch_bounds__(32) fsd_plan_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, int nblocks,
                                                       const uint32_t* __restrict__ hist7, const uint32_t* __restrict__ T, FsdInfo* __restrict__ infos,
                                                       uint8_t* __restrict__ out, FsdResult* __restrict__ res) {
    const int b = blockIdx.x;
    if (b >= nblocks) return;
    const int lane = threadIdx.x;
    const FsdBlock blk = blocks[b];
    FsdInfo& I = infos[b];
    const uint32_t count = blk.len;
    const uint8_t* src = in + blk.src_off;
    uint8_t* dst = out + blk.dst_off;
    const uint32_t count10 = count / 10, count5 = 2 * count10;
    uint32_t go = 0, dist = 0;
    int32_t dt_out = blk.data_type;
    int ent0 = 0;
    if (lane == 0) {
        go = count >= 1024 && blk.cap >= fsd_max_encoded_len(count);                                          // :145-153
        if (blk.data_type != DT_UNDEFINED && blk.data_type != DT_MULTIMEDIA && blk.data_type != DT_BIN) go = 0;  // :155-163
        if (go && !magic_allows_fsd(src, count)) go = 0;                                                        // :165-187
        if (go) {
            const uint32_t* h = hist7 + (size_t)b * 7 * 256;
            int ent[7];
            int min_idx = 0;
            for (int i = 0; i < 7; i++) {
                ent[i] = entropy_1024(3 * count10, h + i * 256, T);
                if (ent[i] < ent[min_idx]) min_idx = i;
            }
            ent0 = ent[0];
            if (ent[min_idx] >= ent[0]) {  // :236-243 the context gets whatever DetectSimpleType says, UNDEFINED included
                dt_out = detect_simple_type_dev(3 * count10, h);
                go = 0;
            } else {
                dt_out = DT_MULTIMEDIA;
                dist = fsd_dist((uint32_t)min_idx);
            }
        }
    }
    go = __shfl_sync(0xFFFFFFFFu, go, 0);
    dist = __shfl_sync(0xFFFFFFFFu, dist, 0);
    uint32_t mode = 0;
    if (go) {  // :253-268 xor coding when more than ~3 % of a sample are large deltas
        uint32_t large = 0;
        for (uint32_t i = 2 * count5 + lane; i < 3 * count5; i += 32) {
            const int delta = (int)src[i] - (int)src[i - dist];
            if (delta < -127 || delta > 127) large++;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) large += __shfl_xor_sync(0xFFFFFFFFu, large, d);
        mode = large > (count5 >> 5) ? 1u : 0u;
    }
    if (lane == 0) {
        I.go = go;
        I.mode = mode;
        I.dist = dist;
        I.ent0 = ent0;
        I.out_len = 0;
        if (go) {
            dst[0] = (uint8_t)mode;
            dst[1] = (uint8_t)dist;
            for (uint32_t i = 0; i < dist; i++) dst[2 + i] = src[i];
        }
        FsdResult r;
        r.status = 1;
        r.out_len = 0;
        r.data_type = dt_out;
        r.mode = mode | (dist << 8);
        res[b] = r;
    }
}

KZ_D uint32_t fsd_token_len(const uint8_t* src, uint32_t i, uint32_t dist) {
    const int d = (int)src[i] - (int)src[i - dist];
    return (d >= -127 && d <= 127) ? 1u : 2u;
}

// delta mode: bytes produced by every tile of 4096 positions (position p codes byte dist + p)
__global__ void __launch_bounds__(256) fsd_tiles_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, const FsdInfo* __restrict__ infos,
                                                         uint32_t* __restrict__ tile_sum, uint32_t ntiles) {
    const int b = blockIdx.y;
    const FsdInfo& I = infos[b];
    if (!I.go || I.mode != 0) return;
    const FsdBlock blk = blocks[b];
    const uint32_t npos = blk.len - I.dist;
    if ((uint64_t)blockIdx.x * FSD_TILE >= npos) return;
    __shared__ uint32_t s_sum[8];
    const uint8_t* src = in + blk.src_off;
    const uint32_t base = blockIdx.x * FSD_TILE + threadIdx.x * 16;
    uint32_t s = 0;
    for (uint32_t j = 0; j < 16; j++)
        if (base + j < npos) s += fsd_token_len(src, I.dist + base + j, I.dist);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_sum[w];
        tile_sum[(size_t)b * ntiles + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) fsd_scan_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, int nblocks, FsdInfo* __restrict__ infos,
                                                       const uint32_t* __restrict__ tile_sum, uint32_t* __restrict__ tile_off, uint32_t ntiles) {
    const int b = blockIdx.x;
    if (b >= nblocks || threadIdx.x != 0) return;
    FsdInfo& I = infos[b];
    if (!I.go) return;
    const FsdBlock blk = blocks[b];
    const uint32_t count = blk.len, dist = I.dist;
    if (I.mode == 1) {  // xor coding: one byte per byte, the loop has no bound test (:287-293)
        I.out_len = count + 2;
        return;
    }
    const uint32_t npos = count - dist;
    const uint32_t nt = (npos + FSD_TILE - 1) / FSD_TILE;
    uint64_t off = 2 + dist;
    for (uint32_t t = 0; t < nt; t++) {
        tile_off[(size_t)b * ntiles + t] = (uint32_t)off;
        off += tile_sum[(size_t)b * ntiles + t];
    }
    // :271-286 the loop stops at dstIdx >= dstEnd - 1; it completes iff the last token starts below that bound
    const uint32_t dst_end = fsd_max_encoded_len(count);
    const uint32_t last = fsd_token_len(in + blk.src_off, count - 1, dist);
    if (off - last >= (uint64_t)dst_end - 1) {
        I.go = 0;  // :295-297 "output buffer too small": skip
        return;
    }
    I.out_len = (uint32_t)off;
}

__global__ void __launch_bounds__(256) fsd_emit_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, const FsdInfo* __restrict__ infos,
                                                        const uint32_t* __restrict__ tile_off, uint32_t ntiles, uint8_t* __restrict__ out) {
    const int b = blockIdx.y;
    const FsdInfo& I = infos[b];
    if (!I.go) return;
    const FsdBlock blk = blocks[b];
    const uint32_t dist = I.dist, npos = blk.len - dist;
    if ((uint64_t)blockIdx.x * FSD_TILE >= npos) return;
    const uint8_t* src = in + blk.src_off;
    uint8_t* dst = out + blk.dst_off;
    const uint32_t base = blockIdx.x * FSD_TILE + threadIdx.x * 16;
    if (I.mode == 1) {
        for (uint32_t j = 0; j < 16; j++) {
            const uint32_t p = base + j;
            if (p < npos) dst[2 + dist + p] = src[dist + p] ^ src[p];
        }
        return;
    }
    __shared__ uint32_t s_sum[8];
    uint32_t lens = 0, s = 0;  // bit j of lens: token j is an escape
    for (uint32_t j = 0; j < 16; j++)
        if (base + j < npos) {
            const uint32_t l = fsd_token_len(src, dist + base + j, dist);
            lens |= (l - 1) << j;
            s += l;
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_sum[warp] = incl;
    __syncthreads();
    uint32_t pre = 0;
    for (int w = 0; w < warp; w++) pre += s_sum[w];
    uint32_t o = tile_off[(size_t)b * ntiles + blockIdx.x] + pre + incl - s;
    for (uint32_t j = 0; j < 16; j++) {
        const uint32_t p = base + j;
        if (p >= npos) break;
        const uint32_t cur = src[dist + p], prv = src[p];
        if ((lens >> j) & 1u) {
            dst[o] = 0xFF;  // _FSD_ESCAPE_TOKEN
            dst[o + 1] = (uint8_t)(cur ^ prv);
            o += 2;
        } else {
            const int d = (int)cur - (int)prv;  // _FSD_ZIGZAG1[127 + d]
            dst[o++] = (uint8_t)(d < 0 ? -2 * d - 1 : 2 * d);
        }
    }
}

// :296-310 histogram of two sub-ranges of the output
__global__ void __launch_bounds__(256) fsd_post_hist_kernel(const FsdBlock* __restrict__ blocks, const FsdInfo* __restrict__ infos, const uint8_t* __restrict__ out,
                                                             uint32_t* __restrict__ hpost) {
    __shared__ uint32_t h[256];
    const int b = blockIdx.y;
    const FsdInfo& I = infos[b];
    if (!I.go) return;
    h[threadIdx.x] = 0;
    __syncthreads();
    const FsdBlock blk = blocks[b];
    const uint32_t count10 = blk.len / 10, count5 = 2 * count10;
    const uint8_t* dst = out + blk.dst_off;
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < count10; k += gridDim.x * 256) {
        atomicAdd(&h[dst[count5 + k]], 1u);
        atomicAdd(&h[dst[3 * count5 + k]], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hpost[(size_t)b * 256 + threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(32) fsd_final_kernel(const FsdBlock* __restrict__ blocks, int nblocks, const FsdInfo* __restrict__ infos,
                                                        const uint32_t* __restrict__ hpost, const uint32_t* __restrict__ T, FsdResult* __restrict__ res) {
    const int b = blockIdx.x;
    if (b >= nblocks || threadIdx.x != 0) return;
    const FsdInfo& I = infos[b];
    if (!I.go) return;
    const uint32_t count10 = blocks[b].len / 10, count5 = 2 * count10;
    if (entropy_1024(count5, hpost + (size_t)b * 256, T) >= I.ent0) return;  // "no improvement": skip
    FsdResult r = res[b];
    r.status = 0;
    r.out_len = I.out_len;
    res[b] = r;
}

// ---- inverse: one warp per block, 32 source bytes per round (FSDCodec.go:317-401) ----
// out[o] = f(out[o - dist]) is `dist` interleaved chains; along a chain the delta tokens are a prefix sum (mod 256) and the rare escape
// tokens (0xFF + payload: out = payload ^ previous) cut it. A round takes 32 source bytes, one per lane:
//   * which 0xFF bytes are escape markers is a parity question inside runs of 0xFF (a marker's successor is a payload, whatever its value),
//     answered with one ballot and a clz per lane; the marker state is carried into the next round;
//   * every byte that is not a marker is one output token; the k-th token moves to lane k (through 64 bytes of shared memory);
//   * token values: inclusive shuffle scan with stride `dist` over the deltas (escape tokens count as 0) on top of the chain's last value
//     of earlier rounds (the last 16 outputs live in the registers of lanes 0..15), then one correction per escape token of the round, in
//     order: E = payload ^ value before it; every later token of its chain moves by E - tentative value;
//   * one coalesced store. XOR coding is the same scan with xor and no escapes.
// ~250 cycles per 32 bytes instead of ~140 per byte for the one-thread walk this replaces.
__global__ void __launch_bounds__(32) fsd_inverse_kernel(const uint8_t* __restrict__ in, const FsdBlock* __restrict__ blocks, int nblocks, uint8_t* __restrict__ out,
                                                          FsdResult* __restrict__ res) {
    __shared__ uint16_t s_tok[32];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const FsdBlock blk = blocks[b];
    FsdResult r;
    r.status = 0;
    r.out_len = 0;
    r.data_type = 0;
    r.mode = 0;
    const uint32_t src_end = blk.len, dst_end = blk.cap;
    const uint8_t* src = in + blk.src_off;
    uint8_t* dst = out + blk.dst_off;
    if (src_end == 0 || dst_end == 0) {
        if (lane == 0) res[b] = r;
        return;
    }
    bool bad = src_end < 2;
    uint32_t mode = 0, dist = 0;
    if (!bad) {
        mode = src[0];
        dist = src[1];
        if (dist < 1 || (dist > 4 && dist != 8 && dist != 16)) bad = true;          // :337-339
        if (!bad && (src_end - 2 < dist || dst_end < dist || mode > 1)) bad = true;  // :346-352, :392-394
    }
    if (bad) {
        r.status = -KZ_E_PROCESS_BLOCK;
        if (lane == 0) res[b] = r;
        return;
    }
    // the first `dist` bytes are copied; tail = the last 16 outputs so far, out[di - 16 + lane] in lanes 0..15
    uint32_t si = 2 + dist, di = dist;
    uint32_t tail = 0;
    if (lane < 16 && (uint32_t)lane + dist >= 16u) tail = src[2 + lane + dist - 16];
    if ((uint32_t)lane < dist) dst[lane] = src[2 + lane];
    uint32_t pending = 0;  // the previous round ended with an escape marker: this round's first byte is its payload
    const uint32_t below = (1u << lane) - 1u;
    uint32_t nxt = si + lane < src_end ? src[si + lane] : 0u;
    while (si < src_end && di < dst_end) {
        const uint32_t t = nxt;
        const uint32_t cnt = min(32u, src_end - si);
        const uint32_t valid = cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u);
        nxt = si + 32 + lane < src_end ? src[si + 32 + lane] : 0u;  // next round in flight
        uint32_t marker = 0;
        if (mode == 0) {
            const uint32_t ff = __ballot_sync(0xFFFFFFFFu, t == 0xFFu) & valid;
            const uint32_t nz = ~ff & below;                                   // bytes below this lane that are not 0xFF
            const uint32_t s = nz ? 32u - (uint32_t)__clz((int)nz) : 0u;        // first lane of the run of 0xFF this lane is in
            const uint32_t par = (s == 0 && pending) ? 1u : 0u;                 // a run that starts with a payload
            const bool is_marker = ((ff >> lane) & 1u) && ((((uint32_t)lane - s) & 1u) == par);
            marker = __ballot_sync(0xFFFFFFFFu, is_marker);
        }
        const uint32_t payload = ((marker << 1) | pending) & valid;  // bytes that follow a marker
        const uint32_t tokens = valid & ~marker;                     // bytes that produce an output byte
        const uint32_t m = (uint32_t)__popc(tokens);
        if (di + m > dst_end) {  // the reference stops at dst_end with source bytes left: an error (:398-400)
            bad = true;
            break;
        }
        // token k -> lane k (through shared memory: every token lane knows its output index by a popc)
        if ((tokens >> lane) & 1u) s_tok[__popc(tokens & below)] = (uint16_t)(t | (((payload >> lane) & 1u) << 8));
        __syncwarp();
        const uint32_t tw = (uint32_t)lane < m ? s_tok[lane] : 0u;
        __syncwarp();
        const uint32_t tk = tw & 0xFFu;
        const bool is_esc = (tw >> 8) != 0;
        uint32_t v;  // value to combine along the chain: the delta, or 0 for an escape token
        if (mode == 0) v = is_esc ? 0u : ((tk & 1u) ? (0u - ((tk + 1u) >> 1)) : (tk >> 1));  // _FSD_ZIGZAG2
        else v = tk;
        if ((uint32_t)lane >= m) v = 0;
        for (uint32_t d = dist; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if ((uint32_t)lane >= d) v = mode == 0 ? v + y : v ^ y;
        }
        // the chain's last value of earlier rounds: out[di + (lane % dist) - dist] = tail lane 16 + (lane % dist) - dist
        const uint32_t head = (uint32_t)lane % dist;
        const uint32_t base = __shfl_sync(0xFFFFFFFFu, tail, (int)(16u + head - dist));
        v = (mode == 0 ? v + base : v ^ base) & 0xFFu;
        if (mode == 0) {
            uint32_t esc = __ballot_sync(0xFFFFFFFFu, is_esc);
            while (esc) {  // escapes of the round in order: every earlier token already has its final value
                const int e = __ffs((int)esc) - 1;
                esc &= esc - 1;
                const uint32_t prev_in = __shfl_sync(0xFFFFFFFFu, v, e >= (int)dist ? e - (int)dist : 0);
                const uint32_t prev_old = __shfl_sync(0xFFFFFFFFu, tail, (int)(16u + (uint32_t)e - dist) & 31);
                const uint32_t prev = e >= (int)dist ? prev_in : prev_old;
                const uint32_t tent = __shfl_sync(0xFFFFFFFFu, v, e);
                const uint32_t pay = __shfl_sync(0xFFFFFFFFu, tk, e);
                const uint32_t delta = ((pay ^ prev) - tent) & 0xFFu;
                if (lane >= e && ((uint32_t)(lane - e) % dist) == 0) v = (v + delta) & 0xFFu;
            }
        }
        if ((uint32_t)lane < m) dst[di + lane] = (uint8_t)v;
        // new tail: out[di + m - 16 + j] for j = 0..15
        {
            const int j = lane & 15;
            const int pos = (int)m - 16 + j;  // index into this round's outputs (negative: still in the old tail)
            const uint32_t from_new = __shfl_sync(0xFFFFFFFFu, v, pos >= 0 ? pos : 0);
            const uint32_t from_old = __shfl_sync(0xFFFFFFFFu, tail, pos >= 0 ? 0 : (pos + 16) & 15);
            tail = pos >= 0 ? from_new : from_old;
        }
        pending = (marker >> 31) & 1u;
        if (cnt < 32) pending = (marker >> (cnt - 1)) & 1u;
        di += m;
        si += cnt;
    }
    if (!bad && pending) bad = true;       // a marker was the last source byte (:380-383)
    if (!bad && si != src_end) bad = true;  // :398-400
    r.out_len = di;
    if (bad) r.status = -KZ_E_PROCESS_BLOCK;
    if (lane == 0) res[b] = r;
}

}  // namespace

size_t fsd_workspace(uint32_t nblocks, uint32_t max_len) {
    const size_t ntiles = (max_len + FSD_TILE - 1) / FSD_TILE + 1;
    return align256(257 * 4) + align256((size_t)nblocks * 7 * 256 * 4) + align256((size_t)nblocks * 256 * 4) + align256((size_t)nblocks * sizeof(FsdInfo)) +
           2 * align256((size_t)nblocks * ntiles * 4) + 256;
}

cudaError_t fsd_forward_batch(const uint8_t* d_in, uint8_t* d_out, const FsdBlock* d_blocks, uint32_t nblocks, uint32_t max_len, uint8_t* ws, size_t ws_bytes,
                              FsdResult* d_res, cudaStream_t stream, uint64_t* launches) {
    if (nblocks == 0) return cudaSuccess;
    if (ws_bytes < fsd_workspace(nblocks, max_len)) return cudaErrorInvalidValue;
    const Ws w = carve(ws, nblocks, max_len);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(w.log2tab, host_log2_4096(), 257 * 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.hist7, 0, (size_t)nblocks * 7 * 256 * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.hpost, 0, (size_t)nblocks * 256 * 4, stream)) != cudaSuccess) return e;
    const uint32_t gx = std::max<uint32_t>(1, (max_len + FSD_TILE - 1) / FSD_TILE);
    fsd_sample_kernel<<<dim3(SAMPLE_SLICES, nblocks), 256, 0, stream>>>(d_in, d_blocks, w.hist7);
    fsd_plan_kernel<<<nblocks, 32, 0, stream>>>(d_in, d_blocks, (int)nblocks, w.hist7, w.log2tab, w.info, d_out, d_res);
    fsd_tiles_kernel<<<dim3(gx, nblocks), 256, 0, stream>>>(d_in, d_blocks, w.info, w.tile_sum, w.ntiles);
    fsd_scan_kernel<<<nblocks, 32, 0, stream>>>(d_in, d_blocks, (int)nblocks, w.info, w.tile_sum, w.tile_off, w.ntiles);
    fsd_emit_kernel<<<dim3(gx, nblocks), 256, 0, stream>>>(d_in, d_blocks, w.info, w.tile_off, w.ntiles, d_out);
    fsd_post_hist_kernel<<<dim3(SAMPLE_SLICES, nblocks), 256, 0, stream>>>(d_blocks, w.info, d_out, w.hpost);
    fsd_final_kernel<<<nblocks, 32, 0, stream>>>(d_blocks, (int)nblocks, w.info, w.hpost, w.log2tab, d_res);
    if (launches) *launches += 7;
    return cudaGetLastError();
}

cudaError_t fsd_inverse_batch(const uint8_t* d_in, uint8_t* d_out, const FsdBlock* d_blocks, uint32_t nblocks, FsdResult* d_res, cudaStream_t stream,
                              uint64_t* launches) {
    if (nblocks == 0) return cudaSuccess;
    fsd_inverse_kernel<<<nblocks, 32, 0, stream>>>(d_in, d_blocks, (int)nblocks, d_out, d_res);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace kz
