// Static rANS order 1 (kanzi "ANS1") on sm_100a — first correct CUDA path.
//
// Reference semantics reproduced bit-exactly (paths relative to the reference's v2/ directory):
//   entropy/ANSRangeCodec.go:98-110   order 1: chunk = 16384 << 8 = 4 MiB, log range = 11
//   entropy/ANSRangeCodec.go:408-427  rebuildStatistics: four order-1 histograms over the quarters (context 0 at each
//                                     quarter start, internal/Global.go:252-303), the len & 3 tail bytes are not counted
//   entropy/ANSRangeCodec.go:171-270  updateFrequencies / encodeHeader: 3 bits lr-8, then 256 per-context tables
//   entropy/ANSRangeCodec.go:353-388  encodeChunk order 1: state k owns quarter k, walks it backwards, symbol table
//                                     index (context << 8) | symbol, first byte of a quarter coded with context 0
//   entropy/ANSRangeCodec.go:918-949  decodeChunkV2 order 1
//
// The format leaves ONE chunk per <= 4 MiB with four serial chains of up to 1 Mi steps: the GPU is latency bound here
// by construction (SURVEY §7 "hard parts"); parallelism only comes from the number of chunks in flight.
// Layout: statistics and tables live in HBM/L2 (256 contexts x 256 symbols do not fit shared memory); the four states
// of a chunk sit in four lanes of a warp and share the byte cursor through one ballot per step, like the order-0 path.
#include "kz_ans1.cuh"

#include "kz_warp.cuh"

namespace kz {

namespace {
const uint32_t LR1 = 11;
}

// grid (nchunks, HIST_CTAS): order-1 histogram with totals: hist[c][257 * ctx + sym], hist[c][257 * ctx + 256] = total
__global__ void __launch_bounds__(256) ans1_hist_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                         uint32_t* __restrict__ hist) {
    const int c = blockIdx.x;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    uint32_t* h = hist + (size_t)c * ANS1_HIST_WORDS;
    const uint32_t len = ck.len;
    const uint32_t quarter = len >> 2;
    const uint32_t counted = quarter ? 4 * quarter : len;  // the len & 3 tail is not part of the statistics when quarter > 0
    const uint32_t tid = blockIdx.y * blockDim.x + threadIdx.x, nthreads = gridDim.y * blockDim.x;
    for (uint32_t i = tid; i < counted; i += nthreads) {
        const bool first = quarter ? (i % quarter == 0) : (i == 0);
        const uint32_t ctx = first ? 0u : src[i - 1];
        atomicAdd(&h[257u * ctx + src[i]], 1u);
        atomicAdd(&h[257u * ctx + 256u], 1u);
    }
}

// one warp per (chunk, context): normalise, table entries, header bits of that context
__global__ void __launch_bounds__(128) ans1_stats_kernel(const uint32_t* __restrict__ hist, int nchunks, uint32_t* __restrict__ enc_tab,
                                                          uint8_t* __restrict__ ctx_hdr, uint32_t* __restrict__ ctx_bits) {
    __shared__ uint32_t s_hdr[4][ANS1_CTX_HDR_STRIDE / 4];
    __shared__ uint16_t s_rank[4][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int id = blockIdx.x * 4 + warp;  // chunk * 256 + ctx
    if (id >= nchunks * 256) return;
    const int c = id >> 8, ctx = id & 255;
    const uint32_t* h = hist + (size_t)c * ANS1_HIST_WORDS + 257u * ctx;
    uint32_t* hw = s_hdr[warp];
    for (int i = lane; i < ANS1_CTX_HDR_STRIDE / 4; i += 32) hw[i] = 0;
    uint32_t f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = h[lane + 32 * j];
    const uint32_t total = h[256];
    __syncwarp();
    const int asz = warp_normalize(f, total, LR1, lane);
    uint32_t carry = 0;
    int rank_base = 0;
    uint32_t last_sym = 0;
    uint32_t* tab = enc_tab + ((size_t)c << 16) + ((size_t)ctx << 8);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t incl = warp_incl_scan(f[j], lane);
        const uint32_t cum = carry + incl - f[j];
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        const int rank = rank_base + __popc(b & ((1u << lane) - 1u));
        uint32_t e = 0;
        if (f[j] != 0) {
            const uint32_t fr = f[j] < (1u << LR1) - 1u ? f[j] : (1u << LR1) - 1u;  // encSymbol.reset :446-468
            const uint32_t bias = fr < 2 ? cum + (1u << LR1) - 1u : cum;
            const uint32_t sh = 31u - (uint32_t)__clz((int)((fr - 1u) | 1u));  // invShift - 32
            e = fr | (sh << 13) | (bias << 17);
            s_rank[warp][rank] = (uint16_t)(f[j] - 1);
        }
        tab[lane + 32 * j] = e;
        if (b) last_sym = 32u * j + (31u - (uint32_t)__clz((int)b));
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        rank_base += __popc(b);
    }
    __syncwarp();
    // per-context header (encodeHeader :216-270): alphabet + frequency groups; the 3 bits of lr are emitted once per chunk by the host
    uint32_t pos = 0;
    if (asz == 256) {
        pos = 2;
    } else if (asz == 0) {
        if (lane == 0) smem_put_bits(hw, 1, 1, 1);
        pos = 2;
    } else {
        const uint32_t last_mask = last_sym >> 3;
        if (lane == 0) {
            smem_put_bits(hw, 0, 1, 1);
            smem_put_bits(hw, 1, last_mask, 5);
        }
        pos = 6;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (f[j] != 0) {
                const uint32_t s = (uint32_t)lane + 32u * j;
                smem_put_bits(hw, pos + 8 * (s >> 3) + (7 - (s & 7)), 1, 1);
            }
        pos += 8 * (last_mask + 1);
    }
    if (asz > 1) {
        const int gs = asz < 64 ? 6 : 8;
        const uint32_t llr = 4;  // smallest llr with 1 << llr > 11
        const int ngroups = (asz - 1 + gs - 1) / gs;
        uint32_t log_max = 0, glen = 0;
        const int first = 1 + lane * gs;
        int count = 0;
        if (lane < ngroups) {
            count = asz - first < gs ? asz - first : gs;
            uint32_t mx = 0;
            for (int k = 0; k < count; k++) mx = max(mx, (uint32_t)s_rank[warp][first + k]);
            log_max = mx ? 32u - (uint32_t)__clz((int)mx) : 0u;
            glen = llr + (uint32_t)count * log_max;
        }
        // asz < 64 uses groups of 6: up to 11 groups; asz >= 64 groups of 8: up to 32 groups -> one group per lane
        const uint32_t gincl = warp_incl_scan(glen, lane);
        const uint32_t goff = pos + gincl - glen;
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
    uint32_t* gh = reinterpret_cast<uint32_t*>(ctx_hdr + (size_t)id * ANS1_CTX_HDR_STRIDE);
    for (int i = lane; i < ANS1_CTX_HDR_STRIDE / 4; i += 32) gh[i] = bswap32(hw[i]);
    if (lane == 0) ctx_bits[id] = pos;
}

// one warp per chunk, lanes 0..3 = the four states (quarters)
__global__ void __launch_bounds__(32) ans1_encode_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                          const uint32_t* __restrict__ enc_tab, const uint32_t* __restrict__ rcp_g,
                                                          uint8_t* __restrict__ ctx_hdr, const uint32_t* __restrict__ ctx_bits, uint8_t* __restrict__ tail_hdr,
                                                          uint8_t* __restrict__ pay, uint64_t pay_stride, uint64_t* __restrict__ seg_src,
                                                          uint64_t* __restrict__ seg_bits, const uint32_t* __restrict__ seg_index, int32_t* __restrict__ status) {
    __shared__ uint32_t s_rcp[2048];
    __shared__ __align__(16) uint8_t s_ring[256];
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= nchunks) return;
    for (int i = lane; i < 2048; i += 32) s_rcp[i] = __ldg(rcp_g + i);
    __syncwarp();
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    const uint32_t len = ck.len;
    const uint32_t end4 = len & ~3u;
    const uint32_t quarter = end4 >> 2;
    const uint32_t si = seg_index[c];
    const uint32_t* tab = enc_tab + ((size_t)c << 16);
    uint8_t* e_words = pay + (size_t)c * pay_stride + (pay_stride - 32);  // 16-byte aligned end of the word region
    // segments 0..255: context headers
    for (int k = lane; k < 256; k += 32) {
        seg_src[si + k] = (uint64_t)(uintptr_t)(ctx_hdr + ((size_t)c * 256 + k) * ANS1_CTX_HDR_STRIDE);
        seg_bits[si + k] = ctx_bits[c * 256 + k];
    }
    if (len > 1 && quarter == 0) {  // :353-362 indexes block[-1]: the reference panics -> ERR_PROCESS_BLOCK for this block
        if (lane == 0) {
            atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
            seg_src[si + 256] = 0;
            seg_bits[si + 256] = 0;
            seg_src[si + 257] = 0;
            seg_bits[si + 257] = 0;
        }
        return;
    }
    if (lane == 0)
        for (uint32_t i = end4; i < len; i++) e_words[i - end4] = src[i];  // tail bytes (:339-342)
    const int k = lane & 3;
    const bool worker = lane < 4;
    const uint32_t grp_mask = 0xFu;
    const uint32_t lo_mask = grp_mask & ((1u << lane) - 1u);
    uint32_t st = ANS_TOP, emitted = 0, flushed = 0;
    const uint8_t* q = src + (size_t)k * quarter;
    const uint32_t ring = smem_u32(s_ring);
    // One rANS step of the four states. The table entry e (freq | (invShift-32) << 13 | bias << 17) and the reciprocal
    // inv do not depend on the state, so the callers fetch them ahead of time: the dependent chain of a step is
    // compare / select / multiply-high / multiply-add only.
    auto core = [&](uint32_t e, uint32_t inv, bool active) {
        const uint32_t freq = e & 0x1FFFu;
        const uint32_t sh = (e >> 13) & 15u;
        const bool emit = active && st >= (freq << (31u - LR1));
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, emit);
        const uint32_t u = emitted + __popc(bal & lo_mask);
        const uint32_t addr = ring + ((0u - 2u * (u + 1u)) & 255u);
        const uint32_t w16 = __byte_perm(st, 0, 0x4401);
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u16 [%0], %1; }" ::"r"(addr), "r"(w16), "r"((uint32_t)emit) : "memory");
        const uint32_t x = emit ? (st >> 16) : st;
        const uint32_t qq = __umulhi(x, inv) >> sh;
        if (active) st = x + (e >> 17) + qq * ((1u << LR1) - freq);
        emitted += __popc(bal & grp_mask);
    };
    // write out every completed 64-byte block of the ring (uniform condition); a batch of ANS1_ENC_BATCH steps adds at
    // most 8 bytes per step, so residual (< 64) + batch (<= 64) stays well inside the 256-byte ring
    auto flush = [&]() {
        __syncwarp();
        while (2u * emitted >= 64u * (flushed + 1u)) {
            if (worker) {
                const uint32_t ro = (0u - 64u * (flushed + 1u)) & 255u;
                const uint4 v = *reinterpret_cast<const uint4*>(s_ring + ro + 16 * k);
                *reinterpret_cast<uint4*>(e_words - 64 * (size_t)(flushed + 1) + 16 * k) = v;
            }
            flushed++;
        }
        __syncwarp();
    };
    constexpr int B = 8;  // steps per batch
    // table entries of steps i, i-1, .., i-B+1 (step j encodes with tab[(q[j] << 8) | q[j+1]]); needs i-B+1 >= 0
    auto load_batch = [&](int64_t i, uint32_t (&E)[B]) {
        uint32_t b[B + 1];
#pragma unroll
        for (int j = 0; j <= B; j++) b[j] = worker ? (uint32_t)q[i + 1 - j] : 0u;
#pragma unroll
        for (int j = 0; j < B; j++) E[j] = worker ? __ldg(tab + ((b[j + 1] << 8) | b[j])) : 0u;
    };
    if (quarter > 0) {
        // positions quarter-1 .. 1 with their true context, then position 0 with context 0 (:364-387)
        int64_t i = (int64_t)quarter - 2;
        uint32_t En[B];
        if (i >= B - 1) load_batch(i, En);
        while (i >= B - 1) {
            uint32_t E[B], I[B];
#pragma unroll
            for (int j = 0; j < B; j++) E[j] = En[j];
            if (i - B >= B - 1) load_batch(i - B, En);  // next batch in flight (L2 / HBM latency) while this one is encoded
#pragma unroll
            for (int j = 0; j < B; j++) I[j] = s_rcp[E[j] & 2047u];
#pragma unroll
            for (int j = 0; j < B; j++) core(E[j], I[j], worker);
            flush();
            i -= B;
        }
        for (; i >= 0; i--) {
            const uint32_t e = worker ? __ldg(tab + (((uint32_t)q[i] << 8) | (uint32_t)q[i + 1])) : 0u;
            core(e, s_rcp[e & 2047u], worker);
            flush();
        }
        const uint32_t e = worker ? __ldg(tab + (uint32_t)q[0]) : 0u;
        core(e, s_rcp[e & 2047u], worker);
        flush();
    }
    // epilogue
    const uint32_t total = 2u * emitted;
    if (worker)
        for (uint32_t o = 64u * flushed + k; o < total; o += 4) e_words[-(int64_t)o - 1] = s_ring[(0u - (o + 1u)) & 255u];
    const uint32_t s0 = __shfl_sync(0xFFFFFFFFu, st, 0), s1 = __shfl_sync(0xFFFFFFFFu, st, 1);
    const uint32_t s2 = __shfl_sync(0xFFFFFFFFu, st, 2), s3 = __shfl_sync(0xFFFFFFFFu, st, 3);
    if (lane == 0) {
        const uint32_t sz = total + (len - end4);
        uint8_t* th = tail_hdr + (size_t)c * 32;
        for (int i = 0; i < 32; i++) th[i] = 0;
        BitWriter bw(th, 0);
        bw.write_varint(sz);
        bw.write(s0, 32);
        bw.write(s1, 32);
        bw.write(s2, 32);
        bw.write(s3, 32);
        seg_src[si + 256] = (uint64_t)(uintptr_t)th;
        seg_bits[si + 256] = bw.pos;
        seg_src[si + 257] = (uint64_t)(uintptr_t)(e_words - total);
        seg_bits[si + 257] = 8ull * sz;
    }
}

// ---------------- decode ----------------
// one warp per chunk: lane 0 parses the 256 tables and builds f2s / symbol entries in HBM, then lanes 0..3 decode
__global__ void __launch_bounds__(32) ans1_decode_kernel(const uint32_t* __restrict__ words, uint64_t words_len, const DecChunk* __restrict__ chunks,
                                                          int nchunks, const uint64_t* __restrict__ chunk_hdr_bit, uint8_t* __restrict__ f2s_all,
                                                          uint32_t* __restrict__ sym_all, uint8_t* __restrict__ out, int32_t* __restrict__ status) {
    __shared__ __align__(16) uint8_t s_ring[256];
    __shared__ uint32_t s_info[16];
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= nchunks) return;
    const DecChunk ck = chunks[c];
    if (status[ck.block] != 0) return;
    uint8_t* f2s = f2s_all + ((size_t)c << (8 + LR1));   // [256][2048]
    uint32_t* symt = sym_all + ((size_t)c << 16);        // [256][256] : freq | cum << 16
    uint8_t* dst = out + ck.out_off;
    int mode = 0;
    if (lane == 0) {
        BitReader br(words, chunk_hdr_bit[c], ck.end_bit);
        const uint32_t lr = 8 + br.read(3);
        bool bad = lr != LR1;  // tables are sized for the encoder's fixed order-1 range (walk kernel reports other ranges)
        const uint32_t scale = 1u << LR1;
        int total_syms = 0;
        for (int ctx = 0; ctx < 256 && !bad; ctx++) {
            uint32_t* st_ = symt + (ctx << 8);
            uint8_t alpha[256];
            int asz = 0;
            if (br.read(1) == 0) {
                if (br.read(1) == 0) {
                    asz = 256;
                    for (int i = 0; i < 256; i++) alpha[i] = (uint8_t)i;
                }
            } else {
                const uint32_t last = br.read(5);
                for (uint32_t i = 0; i <= last; i++) {
                    uint32_t m = br.read(8);
                    while (m) {
                        const int j = __ffs((int)m) - 1;
                        m &= m - 1;
                        alpha[asz++] = (uint8_t)(8 * i + j);
                    }
                }
            }
            if (asz == 0) continue;
            total_syms += asz;
            uint32_t fr[256];
            uint32_t sum = 0;
            const int gs = asz < 64 ? 6 : 8;
            for (int i = 1; i < asz; i += gs) {
                const uint32_t log_max = br.read(4);
                if ((1u << log_max) > scale) bad = true;
                const int endj = i + gs < asz ? i + gs : asz;
                for (int j = i; j < endj; j++) {
                    uint32_t freq = 1;
                    if (log_max > 0) {
                        freq = 1 + br.read(log_max);
                        if (freq >= scale) bad = true;
                    }
                    fr[j] = freq;
                    sum += freq;
                }
            }
            if (scale <= sum && asz > 1) bad = true;
            if (bad) break;
            fr[0] = scale - sum;
            uint32_t cum = 0;
            uint8_t* fs = f2s + ((size_t)ctx << LR1);
            for (int i = 0; i < asz; i++) {
                const uint32_t f = fr[i];
                for (uint32_t j = 0; j < f && cum + j < scale; j++) fs[cum + j] = alpha[i];
                const uint32_t fc = f < scale - 1 ? f : scale - 1;  // decSymbol.reset :973-978
                st_[alpha[i]] = fc | (cum << 16);
                cum += f;
            }
        }
        if (!bad && total_syms == 0) bad = true;
        uint32_t sz = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        if (!bad) {
            sz = br.read_varint();
            s0 = br.read(32);
            s1 = br.read(32);
            s2 = br.read(32);
            s3 = br.read(32);
            if (sz >= (uint32_t)ANS_MAX_CHUNK || br.overrun || br.pos + 8ull * sz > ck.end_bit) bad = true;
        }
        if (bad || br.overrun) {
            atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
        } else {
            mode = 1;
            s_info[0] = s0;
            s_info[1] = s1;
            s_info[2] = s2;
            s_info[3] = s3;
            s_info[4] = (uint32_t)br.pos;
            s_info[5] = (uint32_t)(br.pos >> 32);
            s_info[6] = sz;
        }
    }
    __threadfence();
    __syncwarp();
    mode = __shfl_sync(0xFFFFFFFFu, mode, 0);
    if (!mode) return;
    const uint64_t pb = ((uint64_t)s_info[5] << 32) | s_info[4];
    const uint64_t pe = pb + 8ull * s_info[6];
    const int k = lane & 3;
    const bool worker = lane < 4;
    uint32_t st = s_info[k];
    __syncwarp();
    const uint32_t end4 = ck.out_len & ~3u;
    const uint32_t quarter = end4 >> 2;
    // payload ring (same scheme as the order-0 decoder): raw bytes, 64-byte rounds by lanes 0..3
    const uint64_t base_b = (pb >> 3) & ~15ull;
    const uint32_t rel = (uint32_t)(pb - 8 * base_b);
    const uint32_t rsh = 16u - (rel & 15u);
    const uint64_t end_b = (pe + 7) >> 3;
    const uint64_t lim_b = words_len * 4;
    uint32_t fill = 0, cur2 = 2u * (rel >> 4);
    const uint32_t ring = smem_u32(s_ring);
    auto issue = [&](bool want) {
        const uint64_t sb = base_b + fill + 16u * k;
        const bool ok = want && worker && sb < end_b && sb + 16 <= lim_b;
        const uint8_t* s = reinterpret_cast<const uint8_t*>(words) + (ok ? sb : 0);
        const uint32_t src_size = ok ? 16u : 0u;
        if (want && worker) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ring + ((fill + 16u * k) & 255u)), "l"(s), "r"(src_size) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int r = 0; r < 4; r++) {
        issue(true);
        fill += 64;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t grp_mask = 0xFu;
    const uint32_t hi_mask = grp_mask & ~((2u << lane) - 1u);
    uint32_t prv = 0;
    uint8_t* o = dst + (size_t)k * quarter;
    bool pending = false;
    for (uint32_t i = 0; i < quarter; i++) {
        uint32_t sym = 0, nst = 0;
        bool need = false;
        if (worker) {
            const uint32_t slot = st & ((1u << LR1) - 1u);
            sym = f2s[((size_t)prv << LR1) + slot];
            const uint32_t e = symt[(prv << 8) + sym];
            nst = (e & 0xFFFFu) * (st >> LR1) + slot - (e >> 16);
            need = nst < (uint32_t)ANS_TOP;
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, need);
        const uint32_t t2 = cur2 + 2u * __popc(bal & hi_mask);
        uint32_t va, vb;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(va) : "r"(ring + (t2 & 254u)));
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(vb) : "r"(ring + ((t2 + 2) & 254u)));
        const uint32_t x = __byte_perm(va, vb, 0x0145) >> rsh;
        if (worker) {
            st = need ? __byte_perm(x, nst, 0x5410) : nst;
            o[i] = (uint8_t)sym;
            prv = sym;
        }
        cur2 += 2u * __popc(bal & grp_mask);
        if ((i & 3) == 3) {  // ring maintenance every 4 steps
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            fill += pending ? 64u : 0u;
            __syncwarp();
            pending = (int32_t)(fill - cur2) <= (256 - 64 - 32);
            issue(pending);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        uint64_t pos = 8 * (base_b + cur2) + (rel & 15u);
        for (uint32_t j = end4; j < ck.out_len; j++, pos += 8) dst[j] = pos + 8 <= pe ? (uint8_t)bits_at(words, pos, 8) : 0;
    }
}

}  // namespace kz
