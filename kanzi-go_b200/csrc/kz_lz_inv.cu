// Data-parallel LZ / LZX inverse on sm_100a. Reference: v2/transform/LZCodec.go:621-778 (inverseV6), readLengthLZ :215-232.
//
// The reference decoder is one loop per block: token -> literal run -> match copy, every match reading bytes the loop wrote earlier.
// Here the two halves are separated:
//
//   parse    one warp per block, 32 tokens per round (one per lane). Everything a token needs is a prefix over the tokens before it —
//            literal stream position (known literal lengths by a warp scan, the 1 / 3 / 4 byte length extensions of literal runs >= 7 chased
//            in order: the only dependent loads of a round), distance stream position (scan of 0..3 bytes per token), match length
//            extensions (chased like the literal ones, in the same loop), repeat distances (scan of the (repd0, repd1) update functions:
//            each component of the new pair is a constant or one of the old pair, closed under composition), output position (scan of
//            literal + match lengths). Output: one record per token (literal source, literal length, match length, distance) and its
//            output offset. Every sanity check of the reference is evaluated; a block that trips one (or that is not shaped like an
//            encoder's stream: early end, missing final literals) is handed to the serial kernel of kz_lz.cu, which reproduces the
//            reference's error behaviour.
//   expand   one lane per token: parent[o] for every output byte o — a literal byte of the stream (flag bit + source offset) or the
//            output byte `dist` behind it.
//   jump     pointer jumping over parent[] (in place, up to 8 hops per byte and pass) until every byte points at a literal: the copy
//            order of the serial decoder is replaced by log(depth) passes over the block; an overlapping match (dist < length) is just a
//            deeper chain.
//   gather   out[o] = stream[parent[o]].
#include <vector>

#include "kz_lz.cuh"

namespace kz {

namespace {

const uint32_t LIT_FLAG = 0x80000000u;
const int64_t LZ_MAX_DISTANCE1 = (1 << 16) - 2, LZ_MAX_DISTANCE2 = (1 << 24) - 2;  // LZCodec.go:34-35
const int EXPAND_SHORT = 48;  // bytes a lane writes on its own; longer literal runs / matches are spread over the warp

struct IBlock {  // device descriptor
    uint64_t src_off, dst_off;
    uint64_t rec_off;     // first record of the block
    uint64_t parent_off;  // first entry of the block in parent[]
    uint32_t len, cap;
    uint32_t active, pad;
};
struct IResult {
    uint32_t status;  // 0: parsed, 1: not handled here (serial kernel), 2: inactive
    uint32_t out_len, n_tokens, pad;
};

__device__ __forceinline__ uint32_t ld_u32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// repeat-distance state update of one token: new (r0, r1) components are a constant (>= 0), OLD0 (-1) or OLD1 (-2)
struct RepFn {
    int32_t a, b;
};
__device__ __forceinline__ int32_t rep_pick(int32_t g, const RepFn& f) { return g >= 0 ? g : (g == -1 ? f.a : f.b); }
__device__ __forceinline__ RepFn rep_then(const RepFn& f, const RepFn& g) {  // first f, then g
    RepFn r;
    r.a = rep_pick(g.a, f);
    r.b = rep_pick(g.b, f);
    return r;
}

__global__ void __launch_bounds__(32) lzi_parse_kernel(const uint8_t* __restrict__ in, const IBlock* __restrict__ blocks, int nblocks, uint4* __restrict__ rec_all,
                                                        uint32_t* __restrict__ dst_all, IResult* __restrict__ res) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const IBlock B = blocks[b];
    IResult R;
    R.status = 2, R.out_len = 0, R.n_tokens = 0, R.pad = 0;
    if (!B.active) {
        if (lane == 0) res[b] = R;
        return;
    }
    R.status = 1;
    const uint8_t* src = in + B.src_off;
    const int64_t count = B.len, cap = B.cap;
    if (count < 13) {
        if (lane == 0) res[b] = R;
        return;
    }
    const int64_t tk0 = ld_u32le(src), tk_n = ld_u32le(src + 4), m_n = ld_u32le(src + 8);
    const int64_t m0 = tk0 + tk_n, mlen0 = m0 + m_n;
    if (tk0 > count || m0 > count || mlen0 > count || tk0 < 13 || tk_n == 0) {
        if (lane == 0) res[b] = R;
        return;
    }
    const int64_t src_end = tk0 - 13, dst_end = cap - 16;
    const int64_t max_dist = (src[12] & 1) ? LZ_MAX_DISTANCE2 : LZ_MAX_DISTANCE1;
    const int32_t min_match = ((src[12] >> 1) & 7) + 2;
    uint4* rec = rec_all + B.rec_off;
    uint32_t* dsts = dst_all + B.rec_off;
    int64_t lit_pos = 13, m_pos = m0, mlen_pos = mlen0, dst_pos = 0;
    int32_t repd0 = (int32_t)count, repd1 = (int32_t)count;
    bool bad = false;
    uint32_t tok_next = lane < tk_n ? src[tk0 + lane] : 0u;
    for (int64_t t0 = 0; t0 < tk_n && !bad; t0 += 32) {
        const uint32_t n = (uint32_t)min((int64_t)32, tk_n - t0);
        const bool live = (uint32_t)lane < n;
        const uint32_t tok = tok_next;
        tok_next = t0 + 32 + lane < tk_n ? src[tk0 + t0 + 32 + lane] : 0u;
        {  // the three byte streams are consumed front to back: pull the next lines into L1 so the chased loads below hit
            const int64_t pf = lane < 4 ? lit_pos + 128 * (lane + 1) : (lane < 6 ? m_pos + 128 * (lane - 3) : mlen_pos + 128 * (lane - 5));
            if (lane < 8 && pf < count) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + pf));
        }
        const bool last = live && t0 + lane == tk_n - 1;  // the final token: literals only
        const uint32_t lit_code = live ? tok >> 5 : 0u;
        const bool lit_ext = live && tok >= 0xE0u;
        const uint32_t f = tok & 0x18u;
        const bool has_match = live && !last;
        const bool rep = f == 0;
        const uint32_t ml_code = rep ? (tok & 3u) : (tok & 7u);
        const bool ml_ext = has_match && ml_code == (rep ? 3u : 7u);
        const uint32_t nb = (has_match && !rep) ? (f == 0x18u ? 3u : (f == 0x10u ? 2u : 1u)) : 0u;
        // exclusive scans: known literal bytes, distance bytes
        uint32_t g = lit_code, mo = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t yg = __shfl_up_sync(0xFFFFFFFFu, g, d), ym = __shfl_up_sync(0xFFFFFFFFu, mo, d);
            if (lane >= d) g += yg, mo += ym;
        }
        const uint32_t g_tot = __shfl_sync(0xFFFFFFFFu, g, 31), m_tot = __shfl_sync(0xFFFFFFFFu, mo, 31);
        g -= lit_code;
        mo -= nb;
        // distance bytes (independent loads, issued before the chases)
        int32_t dist_raw = 0;
        if (nb) {
            const int64_t p = m_pos + mo;
            if (p + nb > count) bad = true;
            else {
                dist_raw = src[p];
                if (nb >= 2) dist_raw = (dist_raw << 8) | src[p + 1];
                if (nb == 3) dist_raw = (dist_raw << 8) | src[p + 2];
            }
        }
        // the length extensions, both streams in one loop: literal runs >= 7 (inside the literal stream) and long matches
        uint32_t lmask = __ballot_sync(0xFFFFFFFFu, lit_ext), mmask = __ballot_sync(0xFFFFFFFFu, ml_ext);
        uint32_t extra_before = 0;  // extension bytes + extension values of the extended literal runs in lower lanes
        uint32_t my_lit_v = 0, my_lit_nb = 0, my_ml_v = 0;
        uint32_t extra = 0;
        while (lmask | mmask) {
            if (lmask) {
                const int e = __ffs((int)lmask) - 1;
                lmask &= lmask - 1;
                const uint32_t ge = __shfl_sync(0xFFFFFFFFu, g, e);
                const int64_t p = lit_pos + ge + extra;
                uint32_t v = 0, nbytes = 1;
                if (p >= count) bad = true;
                else {
                    v = src[p];
                    if (v == 254) {
                        if (p + 3 <= count) v += ((uint32_t)src[p + 1] << 8) + src[p + 2], nbytes = 3;
                        else bad = true;
                    } else if (v == 255) {
                        if (p + 4 <= count) v += ((uint32_t)src[p + 1] << 16) + ((uint32_t)src[p + 2] << 8) + src[p + 3], nbytes = 4;
                        else bad = true;
                    }
                }
                if (lane == e) my_lit_v = v, my_lit_nb = nbytes;
                extra += nbytes + v;
                if (lane > e) extra_before = extra;
            }
            if (mmask) {
                const int e = __ffs((int)mmask) - 1;
                mmask &= mmask - 1;
                const int64_t p = mlen_pos;
                uint32_t v = 0, nbytes = 1;
                if (p >= count) bad = true;
                else {
                    v = src[p];
                    if (v == 254) {
                        if (p + 3 <= count) v += ((uint32_t)src[p + 1] << 8) + src[p + 2], nbytes = 3;
                        else bad = true;
                    } else if (v == 255) {
                        if (p + 4 <= count) v += ((uint32_t)src[p + 1] << 16) + ((uint32_t)src[p + 2] << 8) + src[p + 3], nbytes = 4;
                        else bad = true;
                    }
                }
                if (lane == e) my_ml_v = v;
                mlen_pos += nbytes;
            }
            if (__any_sync(0xFFFFFFFFu, bad)) break;
        }
        bad = __any_sync(0xFFFFFFFFu, bad);
        if (bad) break;
        const uint32_t lit_len = lit_ext ? 7u + my_lit_v : lit_code;
        const int64_t lit_src = lit_pos + g + extra_before + my_lit_nb;
        const int64_t lit_after = lit_src + lit_len;  // srcIdx behind this token's literals
        const uint32_t m_len = has_match ? ml_code + (uint32_t)min_match + my_ml_v : 0u;
        // repeat distances: scan of the state updates
        RepFn fn;
        fn.a = -1, fn.b = -2;
        if (has_match) {
            if (!rep) fn.a = dist_raw, fn.b = -1;
            else if ((tok & 4u) == 0) fn.a = -1, fn.b = -1;
            else fn.a = -2, fn.b = -1;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            RepFn y;
            y.a = __shfl_up_sync(0xFFFFFFFFu, fn.a, d);
            y.b = __shfl_up_sync(0xFFFFFFFFu, fn.b, d);
            if (lane >= d) fn = rep_then(y, fn);
        }
        RepFn in0;
        in0.a = repd0, in0.b = repd1;
        const int32_t after0 = rep_pick(fn.a, in0), after1 = rep_pick(fn.b, in0);
        int32_t before0 = __shfl_up_sync(0xFFFFFFFFu, after0, 1), before1 = __shfl_up_sync(0xFFFFFFFFu, after1, 1);
        if (lane == 0) before0 = repd0, before1 = repd1;
        const int32_t dist = !has_match ? 0 : (!rep ? dist_raw : ((tok & 4u) == 0 ? before0 : before1));
        // output offsets
        uint32_t o = lit_len + m_len;
        uint64_t o64 = (uint64_t)lit_len + m_len;
        if (o64 > 0x7FFFFFFFull) bad = true;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, o, d);
            if (lane >= d) o += y;
        }
        const uint64_t o_tot = __shfl_sync(0xFFFFFFFFu, o, 31);
        const int64_t dst = dst_pos + (o - (lit_len + m_len));
        // the reference's checks (:676-681 literals, :737-739 match) and the shape of an encoder's stream
        if (live) {
            if (lit_after > count || dst + lit_len > cap) bad = true;
            if (tok >= 32 && !last && lit_after >= src_end) bad = true;  // the reference would stop here
            if (last && (tok < 32 || lit_after != tk0)) bad = true;       // ... and must stop at the last token, on the last literal
            if (has_match) {
                const int64_t m_at = dst + lit_len;
                if (dist <= 0 || dist > max_dist || m_at - dist < 0 || m_at + m_len > dst_end) bad = true;
            }
            if (dst + lit_len + m_len > 0x7FFFFFF0ll || lit_after > 0x7FFFFFF0ll) bad = true;
        }
        bad = __any_sync(0xFFFFFFFFu, bad);
        if (bad) break;
        if (live) {
            rec[t0 + lane] = make_uint4((uint32_t)lit_src, lit_len, m_len, (uint32_t)dist);
            dsts[t0 + lane] = (uint32_t)dst;
        }
        // carry
        lit_pos += g_tot + extra;
        m_pos += m_tot;
        dst_pos += o_tot;
        repd0 = __shfl_sync(0xFFFFFFFFu, after0, 31);
        repd1 = __shfl_sync(0xFFFFFFFFu, after1, 31);
    }
    if (lane == 0) {
        R.status = bad ? 1u : 0u;
        R.out_len = bad ? 0u : (uint32_t)dst_pos;
        R.n_tokens = (uint32_t)tk_n;
        res[b] = R;
    }
}

// one lane per token: parents of its literal bytes and of its match bytes
__global__ void __launch_bounds__(256) lzi_expand_kernel(const IBlock* __restrict__ blocks, const IResult* __restrict__ res, const uint4* __restrict__ rec_all,
                                                          const uint32_t* __restrict__ dst_all, uint32_t* __restrict__ parent_all) {
    const int b = blockIdx.y;
    const IResult R = res[b];
    if (R.status != 0) return;
    const IBlock& B = blocks[b];
    const uint4* rec = rec_all + B.rec_off;
    const uint32_t* dsts = dst_all + B.rec_off;
    uint32_t* parent = parent_all + B.parent_off;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * 256) >> 5;
    for (uint32_t t0 = ((blockIdx.x * 256 + threadIdx.x) >> 5) * 32; t0 < R.n_tokens; t0 += warps * 32) {
        const uint32_t t = t0 + lane;
        uint4 r = make_uint4(0, 0, 0, 0);
        uint32_t d = 0;
        if (t < R.n_tokens) r = rec[t], d = dsts[t];
        const uint32_t lit_src = r.x, lit_len = r.y, m_len = r.z, dist = r.w;
        if (lit_len <= EXPAND_SHORT)
            for (uint32_t k = 0; k < lit_len; k++) parent[d + k] = LIT_FLAG | (lit_src + k);
        const uint32_t m_at = d + lit_len;
        if (m_len <= EXPAND_SHORT)
            for (uint32_t k = 0; k < m_len; k++) parent[m_at + k] = m_at + k - dist;
        uint32_t longs = __ballot_sync(0xFFFFFFFFu, lit_len > EXPAND_SHORT);
        while (longs) {
            const int e = __ffs((int)longs) - 1;
            longs &= longs - 1;
            const uint32_t s = __shfl_sync(0xFFFFFFFFu, lit_src, e), n = __shfl_sync(0xFFFFFFFFu, lit_len, e), at = __shfl_sync(0xFFFFFFFFu, d, e);
            for (uint32_t k = lane; k < n; k += 32) parent[at + k] = LIT_FLAG | (s + k);
        }
        longs = __ballot_sync(0xFFFFFFFFu, m_len > EXPAND_SHORT);
        while (longs) {
            const int e = __ffs((int)longs) - 1;
            longs &= longs - 1;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, m_len, e), at = __shfl_sync(0xFFFFFFFFu, m_at, e), ds = __shfl_sync(0xFFFFFFFFu, dist, e);
            for (uint32_t k = lane; k < n; k += 32) parent[at + k] = at + k - ds;
        }
    }
}
// pointer jumping in place; unresolved[b] != 0 afterwards: another pass is needed
__global__ void __launch_bounds__(256) lzi_jump_kernel(const IBlock* __restrict__ blocks, const IResult* __restrict__ res, const uint32_t* __restrict__ todo,
                                                        uint32_t* __restrict__ parent_all, uint32_t* __restrict__ unresolved) {
    const int b = blockIdx.y;
    const IResult R = res[b];
    if (R.status != 0 || !todo[b]) return;
    uint32_t* parent = parent_all + blocks[b].parent_off;
    bool open = false;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < R.out_len; i += gridDim.x * 256) {
        uint32_t p = parent[i];
        if (p & LIT_FLAG) continue;
#pragma unroll 1
        for (int hop = 0; hop < 8 && !(p & LIT_FLAG); hop++) p = __ldcg(parent + p);
        parent[i] = p;
        open = open || !(p & LIT_FLAG);
    }
    if (__syncthreads_or(open ? 1 : 0) && threadIdx.x == 0) unresolved[b] = 1;
}
__global__ void __launch_bounds__(256) lzi_gather_kernel(const uint8_t* __restrict__ in, const IBlock* __restrict__ blocks, const IResult* __restrict__ res,
                                                          const uint32_t* __restrict__ parent_all, uint8_t* __restrict__ out) {
    const int b = blockIdx.y;
    const IResult R = res[b];
    if (R.status != 0) return;
    const IBlock& B = blocks[b];
    const uint8_t* src = in + B.src_off;
    uint8_t* dst = out + B.dst_off;
    const uint32_t* parent = parent_all + B.parent_off;
    // four output bytes per thread: one 32-bit store
    const uint32_t n4 = R.out_len >> 2;
    for (uint32_t q = blockIdx.x * 256 + threadIdx.x; q < n4; q += gridDim.x * 256) {
        const uint4 p = reinterpret_cast<const uint4*>(parent)[q];
        const uint32_t v = (uint32_t)src[p.x & ~LIT_FLAG] | ((uint32_t)src[p.y & ~LIT_FLAG] << 8) | ((uint32_t)src[p.z & ~LIT_FLAG] << 16) | ((uint32_t)src[p.w & ~LIT_FLAG] << 24);
        reinterpret_cast<uint32_t*>(dst)[q] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (R.out_len & 3u)) {
        const uint32_t i = (n4 << 2) + threadIdx.x;
        dst[i] = src[parent[i] & ~LIT_FLAG];
    }
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

// LZ / LZX inverse of a batch. lb[b].len == 0: skipped (status 0, length 0). out_len[b] / status[b] as lz_inverse_kernel leaves them (status 0 or
// -KZ_E_PROCESS_BLOCK). d_lb: the same descriptors on the device (for the serial kernel). Synchronises the stream.
cudaError_t lz_inverse_parallel(const uint8_t* d_in, uint8_t* d_out, const std::vector<LzBlock>& lb, LzWorkspace& W, cudaStream_t stream, std::vector<uint32_t>& out_len,
                                std::vector<int32_t>& status, LzHook* hook, uint64_t* launches) {
    const uint32_t nblocks = (uint32_t)lb.size();
    out_len.assign(nblocks, 0);
    status.assign(nblocks, 0);
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e;
    // the three section sizes at the head of every stream: the token count sizes the record arrays
    std::vector<uint8_t> heads((size_t)nblocks * 16, 0);
    {
        uint32_t any = 0;
        for (uint32_t b = 0; b < nblocks; b++)
            if (lb[b].len >= 13) {
                if ((e = cudaMemcpyAsync(heads.data() + (size_t)b * 16, d_in + lb[b].src_off, 12, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
                any++;
            }
        if (any && (e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    }
    std::vector<IBlock> ib(nblocks);
    uint64_t recs = 0, parents = 0;
    uint32_t max_tokens = 1, max_cap = 1;
    for (uint32_t b = 0; b < nblocks; b++) {
        IBlock& B = ib[b];
        B.src_off = lb[b].src_off, B.dst_off = lb[b].dst_off, B.len = lb[b].len, B.cap = lb[b].cap;
        B.active = lb[b].len != 0 ? 1u : 0u;
        B.pad = 0;
        uint32_t tk_n = 0;
        memcpy(&tk_n, heads.data() + (size_t)b * 16 + 4, 4);
        if (tk_n > lb[b].len) tk_n = 0;  // the kernel rejects the block
        B.rec_off = recs;
        recs += ((uint64_t)tk_n + 63) & ~63ull;
        B.parent_off = parents;
        parents += ((uint64_t)lb[b].cap + 63) & ~63ull;
        if (B.active) max_tokens = std::max(max_tokens, tk_n), max_cap = std::max(max_cap, lb[b].cap);
    }
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off = align256(off + bytes);
        return at;
    };
    const size_t o_ib = take(nblocks * sizeof(IBlock)), o_res = take(nblocks * sizeof(IResult)), o_flag = take((size_t)nblocks * 4), o_todo = take((size_t)nblocks * 4);
    const size_t o_lb = take(nblocks * sizeof(LzBlock)), o_sl = take((size_t)nblocks * 4), o_ss = take((size_t)nblocks * 4);
    const size_t o_rec = take(recs * sizeof(uint4)), o_dst = take(recs * 4), o_par = take(parents * 4);
    uint8_t* ws = W.ensure(off + 4096);
    if (!ws) return cudaErrorMemoryAllocation;
    IBlock* d_ib = (IBlock*)(ws + o_ib);
    IResult* d_res = (IResult*)(ws + o_res);
    uint32_t *d_flag = (uint32_t*)(ws + o_flag), *d_todo = (uint32_t*)(ws + o_todo);
    uint4* d_rec = (uint4*)(ws + o_rec);
    uint32_t *d_dst = (uint32_t*)(ws + o_dst), *d_par = (uint32_t*)(ws + o_par);
    if ((e = cudaMemcpyAsync(d_ib, ib.data(), nblocks * sizeof(IBlock), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    {
        LzHookScope hs(hook, "lzi_parse");
        lzi_parse_kernel<<<nblocks, 32, 0, stream>>>(d_in, d_ib, (int)nblocks, d_rec, d_dst, d_res);
    }
    const uint32_t gx_t = std::min<uint32_t>(std::max<uint32_t>((max_tokens + 255) / 256, 1), 512);
    {
        LzHookScope hs(hook, "lzi_expand");
        lzi_expand_kernel<<<dim3(gx_t, nblocks), 256, 0, stream>>>(d_ib, d_res, d_rec, d_dst, d_par);
    }
    if (launches) *launches += 2;
    std::vector<IResult> h_res(nblocks);
    if ((e = cudaMemcpyAsync(h_res.data(), d_res, nblocks * sizeof(IResult), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    std::vector<uint32_t> todo(nblocks, 0), flags(nblocks, 0);
    bool any_todo = false;
    for (uint32_t b = 0; b < nblocks; b++) todo[b] = h_res[b].status == 0 ? 1u : 0u, any_todo = any_todo || todo[b];
    const uint32_t gx_o = std::min<uint32_t>(std::max<uint32_t>((max_cap + 1023) / 1024, 1), 1024);
    for (int pass = 0; any_todo && pass < 64; pass++) {
        if ((e = cudaMemcpyAsync(d_todo, todo.data(), (size_t)nblocks * 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(d_flag, 0, (size_t)nblocks * 4, stream)) != cudaSuccess) return e;
        {
            LzHookScope hs(hook, "lzi_jump");
            lzi_jump_kernel<<<dim3(gx_o, nblocks), 256, 0, stream>>>(d_ib, d_res, d_todo, d_par, d_flag);
        }
        if (launches) *launches += 1;
        if ((e = cudaMemcpyAsync(flags.data(), d_flag, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        any_todo = false;
        for (uint32_t b = 0; b < nblocks; b++) todo[b] = flags[b], any_todo = any_todo || flags[b];
    }
    // a chain that did not resolve in 64 passes cannot come from a stream the parse accepted (every pass shortens it 8-fold); treat as serial
    for (uint32_t b = 0; b < nblocks; b++)
        if (todo[b]) h_res[b].status = 1;
    if (any_todo && (e = cudaMemcpyAsync(d_res, h_res.data(), nblocks * sizeof(IResult), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    {
        LzHookScope hs(hook, "lzi_gather");
        lzi_gather_kernel<<<dim3(gx_o, nblocks), 256, 0, stream>>>(d_in, d_ib, d_res, d_par, d_out);
    }
    if (launches) *launches += 1;
    // blocks the parse did not take: the serial kernel (error behaviour of the reference included)
    std::vector<LzBlock> sb;
    std::vector<uint32_t> sidx;
    for (uint32_t b = 0; b < nblocks; b++) {
        if (h_res[b].status == 0) out_len[b] = h_res[b].out_len;
        else if (h_res[b].status == 1) sb.push_back(lb[b]), sidx.push_back(b);
    }
    if (!sb.empty()) {
        LzBlock* d_sb = (LzBlock*)(ws + o_lb);
        uint32_t* d_sl = (uint32_t*)(ws + o_sl);
        int32_t* d_ss = (int32_t*)(ws + o_ss);
        if ((e = cudaMemcpyAsync(d_sb, sb.data(), sb.size() * sizeof(LzBlock), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        {
            LzHookScope hs(hook, "lz_inverse_serial");
            lz_inverse_kernel<<<(uint32_t)sb.size(), 32, 0, stream>>>(d_in, d_sb, (int)sb.size(), d_out, d_sl, d_ss);
        }
        if (launches) *launches += 1;
        std::vector<uint32_t> hl(sb.size());
        std::vector<int32_t> hs2(sb.size());
        if ((e = cudaMemcpyAsync(hl.data(), d_sl, sb.size() * 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(hs2.data(), d_ss, sb.size() * 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        for (size_t k = 0; k < sb.size(); k++) out_len[sidx[k]] = hl[k], status[sidx[k]] = hs2[k];
    }
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace kz
