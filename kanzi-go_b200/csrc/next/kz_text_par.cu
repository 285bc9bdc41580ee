// NEXT ROUND — not part of the build (compile-checked only: nvcc -c). Kernels around the three phases of kz_text_par_core.cuh, meant to
// replace text_forward_walk_kernel of kz_text.cu once a GPU run has confirmed them against tests/test_gpu_text.py.
//
//   A1 text_nontext_tiles_kernel   per tile of 4096 positions: last non-letter position inside the tile
//   A2 text_tile_carry_kernel      per block (one warp): exclusive running maximum over the tiles -> last non-letter before each tile;
//                                  also turns the per-tile candidate / found counts into offsets (used twice)
//   A3 text_candidates_kernel      per tile, two modes: count the candidate words of the tile / write them at the tile's offset
//   B  text_dictionary_kernel      per block, one thread: dictionary_pass over the candidates (the only serial part)
//   C1 text_found_cost_kernel      per found word: bytes of (literal span + optional flip byte + word index)
//   C2 (prefix sum over the found words: cub::DeviceScan per block, or the tile-carry kernel again)
//   C3 text_emit_kernel            per found word: writes its span and index at its offset; one extra thread writes the leading spaces and
//                                  the tail span; the size checks of Forward :1338-1344 / :1382-1389 become comparisons on the prefix sums
#include "../kz_common.cuh"
#include "kz_text_par_core.cuh"

namespace kz {
namespace textpar {

using namespace textc;

static const int TILE = 4096;  // positions per tile = 256 threads x 16

struct ParBlock {
    uint64_t src_off, dst_off;
    uint32_t len;
    uint32_t mode;        // mode byte of the block (text_plan_kernel), 0x100 set when the block goes on
    uint32_t lead;        // number of leading spaces
    uint32_t pad;
};

// block-wide inclusive max-scan of one value per thread (256 threads)
KZ_D int cta_incl_max(int v, int* s_w /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v = max(v, t);
    }
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    int pre = -1;
    for (int w = 0; w < warp; w++) pre = max(pre, s_w[w]);
    __syncthreads();
    return max(v, pre);
}
KZ_D uint32_t cta_incl_add(uint32_t v, uint32_t* s_w /*[8]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
    for (int w = 0; w < 8; w++) {
        if (w < warp) pre += s_w[w];
        tot += s_w[w];
    }
    *total = tot;
    __syncthreads();
    return v + pre;
}

// ---- A1: last non-letter position of every tile (-1 if none)
__global__ void __launch_bounds__(256) text_nontext_tiles_kernel(const uint8_t* __restrict__ in, const ParBlock* __restrict__ blocks, int* __restrict__ tile_last,
                                                                  uint32_t ntiles) {
    __shared__ int s_w[8];
    const int b = blockIdx.y;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u) || (uint64_t)blockIdx.x * TILE >= blk.len) return;
    const uint8_t* src = in + blk.src_off;
    const uint32_t base = blockIdx.x * TILE + threadIdx.x * 16;
    int last = -1;
    for (uint32_t j = 0; j < 16; j++)
        if (base + j < blk.len && !is_text(src[base + j])) last = (int)(base + j);
    const int incl = cta_incl_max(last, s_w);
    if (threadIdx.x == 255) tile_last[(size_t)b * ntiles + blockIdx.x] = incl;
}

// ---- A2: per block, serial over tiles. mode 0: tile_last (inclusive, per tile) -> carry_in[t] = last non-letter before tile t.
//          mode 1: counts[t] -> offsets[t] (exclusive sum), total in totals[b]
__global__ void __launch_bounds__(32) text_tile_carry_kernel(const ParBlock* __restrict__ blocks, int nblocks, int mode, uint32_t tiles_per_unit, int* __restrict__ a,
                                                              uint32_t* __restrict__ totals, uint32_t ntiles) {
    const int b = blockIdx.x;
    if (b >= nblocks || threadIdx.x != 0) return;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u)) return;
    const uint32_t nt = (blk.len + tiles_per_unit - 1) / tiles_per_unit;
    int* t = a + (size_t)b * ntiles;
    if (mode == 0) {
        int run = -1;
        for (uint32_t i = 0; i < nt; i++) {
            const int v = t[i];
            t[i] = run;
            run = max(run, v);
        }
    } else {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < nt; i++) {
            const uint32_t v = (uint32_t)t[i];
            t[i] = (int)acc;
            acc += v;
        }
        totals[b] = acc;
    }
}

// ---- A3: candidates of a tile. write == 0: counts[tile] = number of candidates; write == 1: candidates written at offsets[tile] + rank
__global__ void __launch_bounds__(256) text_candidates_kernel(const uint8_t* __restrict__ in, const ParBlock* __restrict__ blocks, const int* __restrict__ carry_in,
                                                               int* __restrict__ counts_or_offsets, uint32_t ntiles, int write,
                                                               Candidate* __restrict__ cand_all, uint64_t cand_stride) {
    __shared__ int s_w[8];
    __shared__ uint32_t s_a[8];
    const int b = blockIdx.y;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u) || (uint64_t)blockIdx.x * TILE >= blk.len) return;
    const uint8_t* src = in + blk.src_off;
    const int count = (int)blk.len;
    const int base = (int)(blockIdx.x * TILE + threadIdx.x * 16);
    // last non-letter before my first position: inside my earlier threads of the tile, else the tile's carry
    int last = -1;
    for (int j = 0; j < 16; j++)
        if (base + j < count && !is_text(src[base + j])) last = base + j;
    const int incl = cta_incl_max(last, s_w);
    __shared__ int s_prev[256];
    s_prev[threadIdx.x] = incl;
    __syncthreads();
    const int before = threadIdx.x ? s_prev[threadIdx.x - 1] : -1;
    const int carry = carry_in[(size_t)b * ntiles + blockIdx.x];
    int prev = max(before, carry);
    Candidate c[16];
    uint32_t mine = 0, mask = 0;
    for (int j = 0; j < 16; j++) {
        const int p = base + j;
        if (p >= count) break;
        // positions inside the leading spaces are never visited by the walk; they cannot end a word anyway (p > prev + 2 fails)
        if (candidate_at(src, count, p, prev, &c[mine])) {
            mask |= 1u << j;
            mine++;
        }
        if (!is_text(src[p])) prev = p;
    }
    uint32_t total;
    const uint32_t incl_n = cta_incl_add(mine, s_a, &total);
    if (!write) {
        if (threadIdx.x == 0) counts_or_offsets[(size_t)b * ntiles + blockIdx.x] = (int)total;
        return;
    }
    Candidate* out = cand_all + (size_t)b * cand_stride + (uint32_t)counts_or_offsets[(size_t)b * ntiles + blockIdx.x] + (incl_n - mine);
    for (uint32_t k = 0; k < mine; k++) out[k] = c[k];
    (void)mask;
}

// ---- B: the dictionary pass (map / list initialised as for text_forward_walk_kernel)
__global__ void __launch_bounds__(32) text_dictionary_kernel(const uint8_t* __restrict__ in, const ParBlock* __restrict__ blocks, int nblocks, int static_n,
                                                              const uint8_t* __restrict__ swords, int32_t* __restrict__ map_all, uint32_t log, Entry* __restrict__ list_all,
                                                              const Candidate* __restrict__ cand_all, uint64_t cand_stride, const uint32_t* __restrict__ ncand,
                                                              Found* __restrict__ found_all, uint32_t* __restrict__ nfound) {
    const int b = blockIdx.x;
    if (b >= nblocks || threadIdx.x != 0) return;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u)) return;
    Dict D;
    D.map = map_all + ((size_t)b << log);
    D.hash_mask = (1u << log) - 1;
    D.list = list_all + (size_t)b * MAX_DICT_SIZE;
    D.dict_size = initial_dict_size((int)blk.len);
    D.static_size = static_n;
    D.swords = swords;
    for (int i = 0; i < static_n; i++) D.map[(uint32_t)D.list[i].hash & D.hash_mask] = i;
    nfound[b] = (uint32_t)dictionary_pass(in + blk.src_off, cand_all + (size_t)b * cand_stride, (int)ncand[b], D, found_all + (size_t)b * cand_stride);
}

// ---- C1: bytes every found word contributes: literal span in front of it + flip byte + index
__global__ void __launch_bounds__(256) text_found_cost_kernel(const uint8_t* __restrict__ in, const ParBlock* __restrict__ blocks, const Found* __restrict__ found_all,
                                                               uint64_t cand_stride, const uint32_t* __restrict__ nfound, uint32_t* __restrict__ cost_all) {
    const int b = blockIdx.y;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u)) return;
    const uint8_t* src = in + blk.src_off;
    const bool is_crlf = (blk.mode & MASK_CRLF) != 0;
    const Found* found = found_all + (size_t)b * cand_stride;
    const uint32_t nf = nfound[b];
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < nf; k += gridDim.x * 256) {
        const int emit_anchor = k ? found[k - 1].start + found[k - 1].length : (int)blk.lead;
        int from, to;
        literal_span(src, emit_anchor, found[k], &from, &to);
        uint32_t c = 0;
        for (int i = from; i < to; i++) c += (uint32_t)symbol_cost(src[i], is_crlf);
        cost_all[(size_t)b * cand_stride + k] = c + (uint32_t)found[k].flip + (uint32_t)index_cost(found[k].index);
    }
}

// ---- C3: emission. offsets = exclusive prefix sum of cost_all per block, starting at 1 + lead. status[b] = 1 when the reference skips.
__global__ void __launch_bounds__(256) text_emit_kernel(const uint8_t* __restrict__ in, const ParBlock* __restrict__ blocks, const Found* __restrict__ found_all,
                                                         uint64_t cand_stride, const uint32_t* __restrict__ nfound, const uint32_t* __restrict__ cost_all,
                                                         const uint32_t* __restrict__ offs_all, uint8_t* __restrict__ out, uint32_t* __restrict__ out_len,
                                                         int32_t* __restrict__ status) {
    const int b = blockIdx.y;
    const ParBlock blk = blocks[b];
    if (!(blk.mode & 0x100u)) return;
    const uint8_t* src = in + blk.src_off;
    uint8_t* dst = out + blk.dst_off;
    const bool is_crlf = (blk.mode & MASK_CRLF) != 0;
    const int count = (int)blk.len, dst_end = count, dst_end3 = count - 3;
    const Found* found = found_all + (size_t)b * cand_stride;
    const uint32_t* cost = cost_all + (size_t)b * cand_stride;
    const uint32_t* offs = offs_all + (size_t)b * cand_stride;
    const uint32_t nf = nfound[b];
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < nf; k += gridDim.x * 256) {
        const Found f = found[k];
        const int emit_anchor = k ? found[k - 1].start + found[k - 1].length : (int)blk.lead;
        int from, to;
        literal_span(src, emit_anchor, f, &from, &to);
        const uint32_t head = cost[k] - (uint32_t)f.flip - (uint32_t)index_cost(f.index);  // bytes of the literal span
        int o = (int)offs[k];
        if (o + (int)head >= dst_end3) {  // Forward :1338-1344
            atomicExch(&status[b], 1);
            continue;
        }
        o += emit_symbols(src + from, to - from, dst + o, dst_end - o, is_crlf);
        if (f.flip) dst[o++] = MASK_FLIP_CASE;
        emit_word_index2(dst + o, f.index);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // mode byte, leading spaces, tail span, final size test
        dst[0] = (uint8_t)blk.mode;
        for (uint32_t i = 0; i < blk.lead; i++) dst[1 + i] = ' ';
        const int emit_anchor = nf ? found[nf - 1].start + found[nf - 1].length : (int)blk.lead;
        int o = nf ? (int)(offs[nf - 1] + cost[nf - 1]) : 1 + (int)blk.lead;
        uint32_t tail = 0;
        for (int i = emit_anchor; i < count; i++) tail += (uint32_t)symbol_cost(src[i], is_crlf);
        if (o + (int)tail > dst_end) {  // :1382-1389
            atomicExch(&status[b], 1);
        } else {
            o += emit_symbols(src + emit_anchor, count - emit_anchor, dst + o, dst_end - o, is_crlf);
            out_len[b] = (uint32_t)o;
        }
    }
}

}  // namespace textpar
}  // namespace kz
