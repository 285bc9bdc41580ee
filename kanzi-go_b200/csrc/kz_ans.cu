// Static rANS (kanzi "ANS0", order 0) on sm_100a: statistics + header, interleaved-state encode, header walk, decode.
//
// Reference semantics reproduced bit-exactly (paths relative to the reference's v2/ directory):
//   entropy/ANSRangeCodec.go:274-311  Write (chunking, <=32 byte raw rule handled by the caller)
//   entropy/ANSRangeCodec.go:408-427  rebuildStatistics      -> ans0_stats_kernel (histogram)
//   entropy/EntropyUtils.go:123-260   NormalizeFrequencies   -> ans0_stats_kernel (warp-parallel restatement)
//   entropy/ANSRangeCodec.go:171-270  updateFrequencies / encodeHeader + EntropyUtils.go:38-67 EncodeAlphabet
//   entropy/ANSRangeCodec.go:446-468  encSymbol.reset        -> packed 8-byte table entries
//   entropy/ANSRangeCodec.go:313-405  encodeSymbol / encodeChunk -> ans0_encode_kernel
//   entropy/ANSRangeCodec.go:605-710  decodeHeader           -> ans_walk_kernel (skim) + ans0_decode_kernel (full)
//   entropy/ANSRangeCodec.go:846-957  decodeSymbol / decodeChunkV2 -> ans0_decode_kernel
//
// Parallel decomposition (the wire format pins it): the unit of independent work is one 16 KiB chunk whose four
// interleaved rANS states share one byte cursor. One THREAD owns one chunk (4-way ILP over the states, the cursor
// coupling is resolved with a prefix count of the four renormalisation flags), 32 chunks per warp, per-chunk tables
// in shared memory laid out bank-interleaved (lane l only ever touches bank l -> conflict free for arbitrary indices).
// Compressed bytes are read / written 16 bytes at a time per thread so that the L1 line throughput of the inherently
// uncoalesced per-chunk streams stays below the dependent-chain latency.
#include "kz_ans.cuh"

namespace kz {

// ------------------------------------------------------------------------------------------------------------
// statistics: one warp per chunk
// ------------------------------------------------------------------------------------------------------------
namespace {

KZ_D void smem_put_bits(uint32_t* words, uint32_t bitpos, uint32_t value, uint32_t n) {  // n in 1..32, MSB-first words
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

}  // namespace

// lr: log range (12 for ANS0). Shared by the order-1 path which calls it per context with lr = 11.
// freq8[j] holds the raw count of symbol (lane + 32 j); on return the normalised frequency. Returns alphabet size.
// Warp-parallel restatement of NormalizeFrequencies (EntropyUtils.go:123-260); all lanes must call.
static __device__ int warp_normalize(uint32_t (&f)[8], uint32_t total, uint32_t lr, int lane) {
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

// encSymbol.reset (ANSRangeCodec.go:446-468) packed into 8 bytes:
//   x = inv_freq (32 bit), y = freq(13) | bias(14) << 13 | (inv_shift-32)(4) << 27   [bias <= cum + 2^lr - 1 < 2^14 for lr <= 13]
static __device__ __forceinline__ uint2 make_enc_entry(uint32_t cum, uint32_t freq, uint32_t lr) {
    uint32_t fr = freq < (1u << lr) - 1u ? freq : (1u << lr) - 1u;
    uint32_t inv, sh, bias;
    if (fr < 2) {
        inv = 0xFFFFFFFFu;
        sh = 0;
        bias = cum + (1u << lr) - 1u;
    } else {
        uint32_t shift = 32u - (uint32_t)__clz((int)(fr - 1));  // smallest shift with fr <= 1<<shift
        inv = (uint32_t)((((uint64_t)1 << (shift + 31)) + (uint64_t)(fr - 1)) / (uint64_t)fr);
        sh = shift - 1;
        bias = cum;
    }
    return make_uint2(inv, fr | (bias << 13) | (sh << 27));
}

// grid: ceil(nchunks / STATS_WARPS) CTAs of STATS_WARPS warps
__global__ void __launch_bounds__(STATS_WARPS * 32) ans0_stats_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                                       uint2* __restrict__ enc_tab, uint8_t* __restrict__ hdr,
                                                                       uint32_t* __restrict__ hdr_bits, uint32_t* __restrict__ asz_out) {
    __shared__ uint32_t s_hist[STATS_WARPS][8][256];  // 8 replicas per warp to spread same-address atomics
    __shared__ uint32_t s_hdr[STATS_WARPS][HDR_STRIDE / 4];
    __shared__ uint16_t s_rank[STATS_WARPS][256];  // freq-1 by alphabet rank
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * STATS_WARPS + warp;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    const uint32_t len = ck.len;
    uint32_t* hist = &s_hist[warp][0][0];
    for (int i = lane; i < 8 * 256; i += 32) hist[i] = 0;
    for (int i = lane; i < HDR_STRIDE / 4; i += 32) s_hdr[warp][i] = 0;
    __syncwarp();
    // histogram (internal/Global.go:220-251); chunk starts are 16-byte aligned
    uint32_t* my = &s_hist[warp][lane & 7][0];
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
    uint32_t f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t s = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) s += s_hist[warp][r][lane + 32 * j];
        f[j] = s;
    }
    const uint32_t lr = 12;
    const int asz = warp_normalize(f, len, lr, lane);
    // cumulative frequencies in symbol order + table entries (updateFrequencies :187-203)
    uint32_t carry = 0;
    int rank_base = 0;
    uint32_t last_sym = 0;
    uint2* tab = enc_tab + (size_t)c * 256;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t incl = warp_incl_scan(f[j], lane);
        uint32_t cum = carry + incl - f[j];
        uint32_t b = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        int rank = rank_base + __popc(b & ((1u << lane) - 1u));
        if (f[j] != 0) {
            tab[lane + 32 * j] = make_enc_entry(cum, f[j], lr);
            s_rank[warp][rank] = (uint16_t)(f[j] - 1);
        } else {
            tab[lane + 32 * j] = make_uint2(0, 0);
        }
        if (b) last_sym = 32u * j + (31u - (uint32_t)__clz((int)b));
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        rank_base += __popc(b);
    }
    __syncwarp();
    // header: 3 bits lr-8, alphabet, frequency groups (encodeHeader :216-270, EncodeAlphabet EntropyUtils.go:38-67)
    uint32_t* hw = s_hdr[warp];
    uint32_t pos = 0;
    if (lane == 0) smem_put_bits(hw, 0, lr - 8, 3);
    pos = 3;
    if (asz == 256) {
        pos += 2;  // '0','0'
    } else if (asz == 0) {
        if (lane == 0) smem_put_bits(hw, pos + 1, 1, 1);  // '0','1'
        pos += 2;
    } else {
        const uint32_t last_mask = last_sym >> 3;
        if (lane == 0) {
            smem_put_bits(hw, pos, 1, 1);
            smem_put_bits(hw, pos + 1, last_mask, 5);
        }
        pos += 6;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (f[j] != 0) {
                uint32_t s = (uint32_t)lane + 32u * j;
                smem_put_bits(hw, pos + 8 * (s >> 3) + (7 - (s & 7)), 1, 1);
            }
        }
        pos += 8 * (last_mask + 1);
    }
    if (asz > 1) {
        const int gs = asz < 64 ? 6 : 8;
        uint32_t llr = 3;
        while ((1u << llr) <= lr) llr++;
        const int ngroups = (asz - 1 + gs - 1) / gs;  // <= 32
        uint32_t log_max = 0, glen = 0;
        int first = 1 + lane * gs, count = 0;
        if (lane < ngroups) {
            count = asz - first < gs ? asz - first : gs;
            uint32_t mx = 0;
            for (int k = 0; k < count; k++) mx = max(mx, (uint32_t)s_rank[warp][first + k]);
            log_max = mx ? 32u - (uint32_t)__clz((int)mx) : 0u;
            glen = llr + (uint32_t)count * log_max;
        }
        uint32_t gincl = warp_incl_scan(glen, lane);
        uint32_t goff = pos + gincl - glen;
        if (lane < ngroups) {
            smem_put_bits(hw, goff, log_max, llr);
            if (log_max) {
                uint32_t p = goff + llr;
                for (int k = 0; k < count; k++, p += log_max) smem_put_bits(hw, p, s_rank[warp][first + k], log_max);
            }
        }
        pos += __shfl_sync(0xFFFFFFFFu, gincl, 31);
    }
    __syncwarp();
    uint32_t* gh = reinterpret_cast<uint32_t*>(hdr + (size_t)c * HDR_STRIDE);
    for (int i = lane; i < HDR_STRIDE / 4; i += 32) gh[i] = bswap32(hw[i]);
    if (lane == 0) {
        hdr_bits[c] = pos;
        asz_out[c] = (uint32_t)asz;
    }
}

// ------------------------------------------------------------------------------------------------------------
// encode: one thread per chunk, one warp (32 chunks) per CTA
// ------------------------------------------------------------------------------------------------------------
struct EncState {
    uint32_t a0, a1, a2, a3;  // 16 bytes of pending output; a0 = lowest address
    uint32_t cnt;             // pending 16-bit words (0..7)
    uint8_t* ptr;             // next 16-byte store goes to ptr-16
};

KZ_D void enc_symbol(uint32_t& st, uint2 e, uint32_t lr, EncState& o) {
    const uint32_t freq = e.y & 0x1FFF;
    const uint32_t bias = (e.y >> 13) & 0x3FFF;
    const uint32_t sh = e.y >> 27;
    const uint32_t x_max = freq << (31 - lr);  // ((ANS_TOP >> lr) << 16) * freq
    if (st >= x_max) {
        // emit st & 0xFFFF: low byte at buffer[n], high byte at buffer[n-1] (:322-326) -> ascending memory = [hi, lo]
        o.a3 = __byte_perm(o.a2, o.a3, 0x5432);
        o.a2 = __byte_perm(o.a1, o.a2, 0x5432);
        o.a1 = __byte_perm(o.a0, o.a1, 0x5432);
        o.a0 = __byte_perm(st, o.a0, 0x5401);
        st >>= 16;
        if (++o.cnt == 8) {
            o.ptr -= 16;
            *reinterpret_cast<uint4*>(o.ptr) = make_uint4(o.a0, o.a1, o.a2, o.a3);
            o.cnt = 0;
        }
    }
    const uint32_t q = __umulhi(st, e.x) >> sh;  // (st * invFreq) >> invShift, invShift = 32 + sh
    st = st + bias + q * ((1u << lr) - freq);
}

__global__ void __launch_bounds__(32) ans0_encode_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                          const uint2* __restrict__ enc_tab, uint8_t* __restrict__ hdr,
                                                          const uint32_t* __restrict__ hdr_bits, const uint32_t* __restrict__ asz_in,
                                                          uint8_t* __restrict__ pay, uint64_t* __restrict__ seg_src, uint64_t* __restrict__ seg_bits,
                                                          const uint32_t* __restrict__ seg_index) {
    extern __shared__ uint2 s_tab[];  // [256][32] : entry of symbol s for lane l at s*32 + l
    const int lane = threadIdx.x;
    const int base = blockIdx.x * 32;
    // cooperative, coalesced table load
    for (int cc = 0; cc < 32; cc++) {
        int c = base + cc;
        if (c >= nchunks) break;
        const uint2* t = enc_tab + (size_t)c * 256;
#pragma unroll
        for (int j = 0; j < 8; j++) s_tab[(lane + 32 * j) * 32 + cc] = __ldg(&t[lane + 32 * j]);
    }
    __syncwarp();
    const int c = base + lane;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint32_t asz = asz_in[c];
    const uint32_t hb = hdr_bits[c];
    const uint32_t si = seg_index[c];
    uint8_t* my_hdr = hdr + (size_t)c * HDR_STRIDE;
    seg_src[si] = (uint64_t)(uintptr_t)my_hdr;
    if (asz <= 1) {  // Write :303: chunk not encoded, header only
        seg_bits[si] = hb;
        seg_src[si + 1] = 0;
        seg_bits[si + 1] = 0;
        return;
    }
    const uint8_t* src = in + ck.src_off;
    const uint32_t len = ck.len;
    const uint32_t end4 = len & ~3u;
    const uint32_t lr = 12;
    uint8_t* e_words = pay + (size_t)c * PAY_STRIDE + PAY_WORDS_END;
    // tail bytes go last in the stream (:339-342)
    for (uint32_t i = end4; i < len; i++) e_words[i - end4] = src[i];
    EncState o;
    o.a0 = o.a1 = o.a2 = o.a3 = 0;
    o.cnt = 0;
    o.ptr = e_words;
    uint32_t st0 = ANS_TOP, st1 = ANS_TOP, st2 = ANS_TOP, st3 = ANS_TOP;
    const uint2* tab = s_tab + lane;
    uint32_t g = end4 >> 2;  // groups of 4 bytes still to encode, processed from the last to the first
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    // :347-352: st0 <- block[i], st1 <- block[i-1], st2 <- block[i-2], st3 <- block[i-3] with i = 4g-1
    while (g & 3) {
        g--;
        uint32_t v = __ldg(&src32[g]);
        enc_symbol(st0, tab[(v >> 24) * 32], lr, o);
        enc_symbol(st1, tab[((v >> 16) & 0xFF) * 32], lr, o);
        enc_symbol(st2, tab[((v >> 8) & 0xFF) * 32], lr, o);
        enc_symbol(st3, tab[(v & 0xFF) * 32], lr, o);
    }
    const uint4* src128 = reinterpret_cast<const uint4*>(src);
    uint32_t q = g >> 2;
    while (q) {
        q--;
        const uint4 v4 = __ldg(&src128[q]);
        const uint32_t w[4] = {v4.w, v4.z, v4.y, v4.x};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v = w[k];
            enc_symbol(st0, tab[(v >> 24) * 32], lr, o);
            enc_symbol(st1, tab[((v >> 16) & 0xFF) * 32], lr, o);
            enc_symbol(st2, tab[((v >> 8) & 0xFF) * 32], lr, o);
            enc_symbol(st3, tab[(v & 0xFF) * 32], lr, o);
        }
    }
    // flush the pending words: they occupy the low 2*cnt bytes of a0..a3 and belong right below o.ptr
    {
        const uint32_t a[4] = {o.a0, o.a1, o.a2, o.a3};
        const uint32_t nb = 2 * o.cnt;
        uint8_t* p = o.ptr - nb;
        for (uint32_t i = 0; i < nb; i++) p[i] = (uint8_t)(a[i >> 2] >> (8 * (i & 3)));
        o.ptr = p;
    }
    const uint32_t sz = (uint32_t)(e_words - o.ptr) + (len - end4);
    // varint(sz) + 4 x 32-bit states appended to the chunk header (:392-399)
    BitWriter bw(my_hdr, hb);
    bw.write_varint(sz);
    bw.write(st0, 32);
    bw.write(st1, 32);
    bw.write(st2, 32);
    bw.write(st3, 32);
    seg_bits[si] = bw.pos;
    seg_src[si + 1] = (uint64_t)(uintptr_t)o.ptr;
    seg_bits[si + 1] = 8ull * sz;
}

// ------------------------------------------------------------------------------------------------------------
// decode side: header walk (one thread per block) and chunk decode (one thread per chunk)
// ------------------------------------------------------------------------------------------------------------
// Skims one alphabet + frequency table (decodeHeader :626-680 without storing anything). Returns alphabet size,
// or -1 on malformed input.
KZ_D int skim_one_table(BitReader& br, uint32_t lr, uint32_t llr) {
    int asz;
    if (br.read(1) == 0) {
        asz = br.read(1) ? 0 : 256;
    } else {
        uint32_t last = br.read(5);
        asz = 0;
        for (uint32_t i = 0; i <= last; i++) asz += __popc(br.read(8));
    }
    if (asz <= 1) return asz;
    const int gs = asz < 64 ? 6 : 8;
    for (int i = 1; i < asz; i += gs) {
        uint32_t log_max = br.read(llr);
        if ((1u << log_max) > (1u << lr)) return -1;
        int n = asz - i < gs ? asz - i : gs;
        br.skip((uint64_t)n * log_max);
    }
    return asz;
}

__global__ void ans_walk_kernel(const uint32_t* __restrict__ words, const DecBlock* __restrict__ blocks, int nblocks, int order,
                                uint64_t* __restrict__ chunk_hdr_bit, uint64_t* __restrict__ blk_end, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const DecBlock blk = blocks[b];
    if (blk.nchunks == 0) return;
    const uint32_t chunk_size = order ? (ANS0_CHUNK << 8) : ANS0_CHUNK;
    BitReader br(words, blk.data_bit, blk.end_bit);
    uint32_t remaining = blk.pre_len;
    int32_t st = 0;
    for (uint32_t k = 0; k < blk.nchunks; k++) {
        chunk_hdr_bit[blk.chunk_base + k] = br.pos;
        if (st) continue;  // descriptors of the remaining chunks point at the failure position; decode is skipped
        const uint32_t clen = remaining < chunk_size ? remaining : chunk_size;
        remaining -= clen;
        const uint32_t lr = 8 + br.read(3);
        uint32_t llr = 3;
        while ((1u << llr) <= lr) llr++;
        int total = 0;
        const int dim = order ? 256 : 1;
        for (int ctx = 0; ctx < dim; ctx++) {
            int a = skim_one_table(br, lr, llr);
            if (a < 0) {
                st = -KZ_E_PROCESS_BLOCK;
                break;
            }
            total += a;
        }
        if (st == 0 && total == 0) st = -KZ_E_PROCESS_BLOCK;
        if (st == 0 && lr > (order ? 11u : 12u)) st = -KZ_E_INVALID_CODEC;  // tables sized for the encoder's fixed ranges
        if (st == 0 && (order == 1 || total > 1)) {
            uint32_t sz = br.read_varint();
            if (sz >= (uint32_t)ANS_MAX_CHUNK) st = -KZ_E_PROCESS_BLOCK;
            br.skip(128 + 8ull * sz);
        }
        if (br.overrun && st == 0) st = -KZ_E_PROCESS_BLOCK;
    }
    blk_end[b] = br.pos;
    if (st) status[b] = st;
}

// big-endian halves of a 16-byte block
KZ_D void be_halves(const uint4 v, uint64_t& h0, uint64_t& h1) {
    h0 = ((uint64_t)bswap32(v.x) << 32) | bswap32(v.y);
    h1 = ((uint64_t)bswap32(v.z) << 32) | bswap32(v.w);
}

__global__ void __launch_bounds__(32) ans0_decode_kernel(const uint32_t* __restrict__ words, uint64_t words_len /*in 32-bit words*/,
                                                          const DecChunk* __restrict__ chunks, int nchunks,
                                                          const uint64_t* __restrict__ chunk_hdr_bit, uint8_t* __restrict__ out,
                                                          int32_t* __restrict__ status) {
    extern __shared__ uint32_t s_dec[];  // [DEC_BM_WORDS][32] bitmap+rank words, then [256][32] symbol entries
    const int lane = threadIdx.x;
    const int c = blockIdx.x * 32 + lane;
    if (c >= nchunks) return;
    const DecChunk ck = chunks[c];
    if (status[ck.block] != 0) return;  // walk failed for this block
    uint32_t* bm = s_dec + lane;                        // word w at bm[w*32]
    uint32_t* symtab = s_dec + DEC_BM_WORDS * 32 + lane;  // entry i at symtab[i*32]
    BitReader br(words, chunk_hdr_bit[c], ck.end_bit);
    // ---- full header parse (decodeHeader :605-710)
    const uint32_t lr = 8 + br.read(3);
    const uint32_t scale = 1u << lr;
    uint32_t llr = 3;
    while ((1u << llr) <= lr) llr++;
    int asz = 0;
    if (br.read(1) == 0) {
        if (br.read(1) == 0) {
            asz = 256;
            for (int i = 0; i < 256; i++) symtab[i * 32] = (uint32_t)i;
        }
    } else {
        uint32_t last = br.read(5);
        for (uint32_t i = 0; i <= last; i++) {
            uint32_t m = br.read(8);
            while (m) {
                int j = __ffs((int)m) - 1;
                m &= m - 1;
                symtab[asz * 32] = 8 * i + (uint32_t)j;
                asz++;
            }
        }
    }
    uint8_t* dst = out + ck.out_off;
    if (asz == 0 || lr > 12) {
        atomicCAS(&status[ck.block], 0, asz == 0 ? -KZ_E_PROCESS_BLOCK : -KZ_E_INVALID_CODEC);
        return;
    }
    if (asz == 1) {  // Read :737-741
        const uint8_t v = (uint8_t)symtab[0];
        for (uint32_t i = 0; i < ck.out_len; i++) dst[i] = v;
        return;
    }
    {
        const int gs = asz < 64 ? 6 : 8;
        uint32_t sum = 0;
        bool bad = false;
        for (int i = 1; i < asz; i += gs) {
            uint32_t log_max = br.read(llr);
            if ((1u << log_max) > scale) bad = true;
            int endj = i + gs < asz ? i + gs : asz;
            for (int j = i; j < endj; j++) {
                uint32_t freq = 1;
                if (log_max > 0) {
                    freq = 1 + br.read(log_max);
                    if (freq >= scale) bad = true;
                }
                symtab[j * 32] |= freq << 8;
                sum += freq;
            }
        }
        if (scale <= sum) bad = true;
        if (bad || br.overrun) {
            atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
            return;
        }
        symtab[0] |= (scale - sum) << 8;
    }
    // ---- tables: 24 start-of-symbol bits + 8-bit (rank-1) per word; entries sym | freq << 8 | cum << 20
    const uint32_t nwords = (scale + 23) / 24;
    for (uint32_t w = 0; w < nwords; w++) bm[w * 32] = 0;
    {
        uint32_t cum = 0;
        for (int i = 0; i < asz; i++) {
            uint32_t e = symtab[i * 32];
            uint32_t f = e >> 8;
            uint32_t w = cum / 24, j = cum - 24 * w;
            bm[w * 32] |= 1u << j;
            uint32_t fc = f < scale - 1 ? f : scale - 1;  // decSymbol.reset :973-978
            symtab[i * 32] = (e & 0xFF) | (fc << 8) | (cum << 20);
            cum += f;
        }
        uint32_t running = 0;
        for (uint32_t w = 0; w < nwords; w++) {
            uint32_t b = bm[w * 32];
            bm[w * 32] = b | (((running - 1u) & 0xFFu) << 24);
            running += __popc(b);
        }
    }
    // ---- chunk payload (decodeChunkV2 :860-957)
    const uint32_t sz = br.read_varint();
    uint32_t st0 = br.read(32), st1 = br.read(32), st2 = br.read(32), st3 = br.read(32);
    if (sz >= (uint32_t)ANS_MAX_CHUNK || br.overrun || br.pos + 8ull * sz > ck.end_bit) {
        atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
        return;
    }
    const uint64_t pb = br.pos;                 // first payload bit
    const uint64_t pe = pb + 8ull * sz;         // one past the last payload bit
    const uint4* blocks128 = reinterpret_cast<const uint4*>(words);
    const uint64_t nblk128 = words_len >> 2;
    uint64_t jn = pb >> 7;
    const uint64_t jlast = sz ? ((pe - 1) >> 7) : 0;
    auto load_blk = [&](uint64_t j) -> uint4 {
        if (sz == 0 || j > jlast || j >= nblk128) return make_uint4(0, 0, 0, 0);  // zero guard (:888-893)
        return __ldg(&blocks128[j]);
    };
    uint64_t q0, q1, c0, c1;
    uint32_t cnt;
    {
        uint64_t h0, h1;
        be_halves(load_blk(jn++), h0, h1);
        const uint32_t off = (uint32_t)(pb & 127);
        if (off >= 64) {
            q0 = h1 << (off - 64);
            q1 = 0;
        } else {
            q0 = off ? ((h0 << off) | (h1 >> (64 - off))) : h0;
            q1 = h1 << off;
        }
        cnt = 128 - off;
    }
    be_halves(load_blk(jn++), c0, c1);
    uint4 nxt = load_blk(jn++);
    uint32_t pfh = 0;
    const uint32_t mask = scale - 1;
    const uint32_t end4 = ck.out_len & ~3u;

    // one interleaved step: decodes 4 symbols, returns them packed as block[i] | block[i+1] << 8 | ...
    auto step = [&]() -> uint32_t {
        if (cnt < 64) {  // top up the window with the next 64 bits
            const uint64_t h = pfh ? c1 : c0;
            q0 |= h >> cnt;
            q1 = cnt ? (h << (64 - cnt)) : 0;
            cnt += 64;
            if (++pfh == 2) {
                be_halves(nxt, c0, c1);
                nxt = load_blk(jn++);
                pfh = 0;
            }
        }
        uint32_t sl3 = st3 & mask, sl2 = st2 & mask, sl1 = st1 & mask, sl0 = st0 & mask;
        uint32_t w3 = sl3 / 24, w2 = sl2 / 24, w1 = sl1 / 24, w0 = sl0 / 24;
        uint32_t b3 = bm[w3 * 32], b2 = bm[w2 * 32], b1 = bm[w1 * 32], b0 = bm[w0 * 32];
        uint32_t i3 = ((b3 >> 24) + __popc(b3 & ((2u << (sl3 - 24 * w3)) - 1u) & 0xFFFFFFu)) & 0xFF;
        uint32_t i2 = ((b2 >> 24) + __popc(b2 & ((2u << (sl2 - 24 * w2)) - 1u) & 0xFFFFFFu)) & 0xFF;
        uint32_t i1 = ((b1 >> 24) + __popc(b1 & ((2u << (sl1 - 24 * w1)) - 1u) & 0xFFFFFFu)) & 0xFF;
        uint32_t i0 = ((b0 >> 24) + __popc(b0 & ((2u << (sl0 - 24 * w0)) - 1u) & 0xFFFFFFu)) & 0xFF;
        uint32_t e3 = symtab[i3 * 32], e2 = symtab[i2 * 32], e1 = symtab[i1 * 32], e0 = symtab[i0 * 32];
        // D(x) = (s, freq * (x >> lr) + (x & mask) - cum)   (:849)
        st3 = ((e3 >> 8) & 0xFFF) * (st3 >> lr) + sl3 - (e3 >> 20);
        st2 = ((e2 >> 8) & 0xFFF) * (st2 >> lr) + sl2 - (e2 >> 20);
        st1 = ((e1 >> 8) & 0xFFF) * (st1 >> lr) + sl1 - (e1 >> 20);
        st0 = ((e0 >> 8) & 0xFFF) * (st0 >> lr) + sl0 - (e0 >> 20);
        // renormalisation: states below ANS_TOP pull 16 bits each from the shared cursor, in the order 3,2,1,0
        const uint32_t n3 = st3 < (uint32_t)ANS_TOP, n2 = st2 < (uint32_t)ANS_TOP, n1 = st1 < (uint32_t)ANS_TOP, n0 = st0 < (uint32_t)ANS_TOP;
        const uint32_t p2 = n3, p1 = n3 + n2, p0 = p1 + n1, tot = p0 + n0;
        const uint32_t qlo = (uint32_t)q0, qhi = (uint32_t)(q0 >> 32);
        if (n3) st3 = (st3 << 16) | (qhi >> 16);
        if (n2) st2 = (st2 << 16) | (__byte_perm(qlo, qhi, 0x7676 - 0x2222 * p2) & 0xFFFF);
        if (n1) st1 = (st1 << 16) | (__byte_perm(qlo, qhi, 0x7676 - 0x2222 * p1) & 0xFFFF);
        if (n0) st0 = (st0 << 16) | (__byte_perm(qlo, qhi, 0x7676 - 0x2222 * p0) & 0xFFFF);
        const uint32_t sh = 16 * tot;
        if (sh == 64) {
            q0 = q1;
            q1 = 0;
        } else if (sh) {
            q0 = (q0 << sh) | (q1 >> (64 - sh));
            q1 <<= sh;
        }
        cnt -= sh;
        return (e3 & 0xFF) | ((e2 & 0xFF) << 8) | ((e1 & 0xFF) << 16) | (e0 << 24);
    };

    uint32_t i = 0;
    for (; i + 16 <= end4; i += 16) {
        uint4 o;
        o.x = step();
        o.y = step();
        o.z = step();
        o.w = step();
        *reinterpret_cast<uint4*>(dst + i) = o;
    }
    for (; i < end4; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = step();
    // tail bytes follow the consumed words (:951-954); cursor = pe - cnt-adjusted position
    {
        // bits consumed so far = (bits loaded into the window) - cnt; recompute the cursor from what is left:
        // window holds cnt bits, c-halves hold (2 - pfh) * 64 bits, nxt block and beyond start at 128 * (jn - 1)
        const uint64_t loaded_end = (jn - 1) << 7;  // first bit of `nxt`
        uint64_t cursor = loaded_end - (uint64_t)(2 - pfh) * 64 - cnt;
        for (uint32_t k = end4; k < ck.out_len; k++, cursor += 8) dst[k] = cursor + 8 <= pe ? (uint8_t)bits_at(words, cursor, 8) : 0;
    }
}

}  // namespace kz
