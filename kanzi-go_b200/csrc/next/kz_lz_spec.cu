// NEXT ROUND — not part of the build (compile-checked only: nvcc -c). Warp-speculative LZ / LZX parse: the CUDA form of the CPU model
// tests/host/lz_spec_host.cpp (which tests/test_lz_spec_host.py checks byte for byte against the oracle). Same interface and scratch
// layout as lz_parse_kernel of kz_lz.cu, so lz_gather_kernel and the host code stay as they are.
//
// All 32 lanes carry the scalar parse state redundantly (warp-uniform values); per batch lane j probes position P_j of the literal run:
//   P_0 = src_idx, P_{j+1} = P_j + 1 + ((src_inc + j) >> 6)          (what the serial parse visits if lanes 0..j all miss)
//   ref0_j = hash table slot of P_j, or P_i of the latest earlier lane i with the same slot (__match_any_sync): the serial view
//   hit_j  = repeat-distance match at P_j + 1, or hash match at P_j                     (lane 0 uses the live repd_idx, the others 0)
// The first hit wins (__ballot_sync / __ffs); lanes up to the winner commit their inserts — per slot the highest such lane writes, which is
// what the serial order leaves there; the match itself is processed warp-uniformly with lane 0 doing the stores, and the inserts of the
// positions inside the match run 32 at a time with atomicMax (positions only grow, so the maximum is the last writer of the serial order).
#include "../kz_lz.cuh"

namespace kz {
namespace lzspec {

namespace {
const uint64_t LZ_HASH_SEED = 0x1E35A7BDull;
const int LZ_MAX_DISTANCE1 = (1 << 16) - 2, LZ_MAX_DISTANCE2 = (1 << 24) - 2, LZ_MAX_MATCH = 65535 + 254 + 4, LZ_MIN_BLOCK = 24;

KZ_D uint64_t ld64(const uint64_t* __restrict__ w, int p) {
    const int i = p >> 3, s = (p & 7) << 3;
    const uint64_t a = w[i];
    if (s == 0) return a;
    return (a >> s) | (w[i + 1] << (64 - s));
}
KZ_D uint32_t ld32(const uint64_t* __restrict__ w, int p) { return (uint32_t)ld64(w, p); }
KZ_D uint32_t lz_hash(uint64_t v, bool extra) {
    return extra ? (uint32_t)(((v << 24) * LZ_HASH_SEED) >> (64 - 19)) : (uint32_t)(((v << 24) * LZ_HASH_SEED) >> (64 - 16));
}
KZ_D int emit_length(uint8_t* block, int length) {
    if (length < 254) {
        block[0] = (uint8_t)length;
        return 1;
    }
    if (length < 65536 + 254) {
        length -= 254;
        block[0] = 254;
        block[1] = (uint8_t)(length >> 8);
        block[2] = (uint8_t)length;
        return 3;
    }
    length -= 255;
    block[0] = 255;
    block[1] = (uint8_t)(length >> 16);
    block[2] = (uint8_t)(length >> 8);
    block[3] = (uint8_t)length;
    return 4;
}
KZ_D int length_bytes(int length) { return length < 254 ? 1 : (length < 65536 + 254 ? 3 : 4); }
KZ_D int find_match(const uint64_t* __restrict__ w, int src_idx, int ref, int max_match) {
    int best = 0;
    while (best + 8 <= max_match) {
        const uint64_t diff = ld64(w, src_idx + best) ^ ld64(w, ref + best);
        if (diff != 0) {
            best += (__ffsll((long long)diff) - 1) >> 3;
            break;
        }
        best += 8;
    }
    return best;
}
}  // namespace

__global__ void __launch_bounds__(32) lz_parse_spec_kernel(const uint8_t* __restrict__ in, const LzBlock* __restrict__ blocks, int nblocks, int extra,
                                                            int32_t* hashes_all, uint8_t* __restrict__ scratch_all, uint8_t* __restrict__ out,
                                                            LzResult* __restrict__ res) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const LzBlock blk = blocks[b];
    LzResult r;
    r.status = 1;
    r.out_len = 0;
    r.n_runs = 0;
    r.lit_end = 0;
    r.tk_n = r.m_n = r.mlen_n = 0;
    const int count = (int)blk.len;
    if (count < LZ_MIN_BLOCK || blk.data_type == 9 /*DT_SMALL_ALPHABET*/) {
        if (lane == 0) res[b] = r;
        return;
    }
    const uint8_t* src = in + blk.src_off;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(src);
    uint8_t* dst = out + blk.dst_off;
    int32_t* hashes = hashes_all + ((size_t)b << (extra ? 19 : 16));
    uint8_t* scratch = scratch_all + blk.scratch_off;
    uint8_t* tk_buf = scratch;
    uint8_t* m_buf = tk_buf + (count / 4 + 64);
    uint8_t* mlen_buf = m_buf + (3 * (count / 4) + 64);
    LzRun* runs = reinterpret_cast<LzRun*>(scratch + (((size_t)(count / 4 + 64) + (3 * (size_t)(count / 4) + 64) + ((size_t)count + 64) + 15) & ~(size_t)15));
    const int src_end = count - 16 - 2;
    int max_dist = LZ_MAX_DISTANCE2;
    uint8_t flags = 1;
    if (src_end < 4 * LZ_MAX_DISTANCE1) {
        max_dist = LZ_MAX_DISTANCE1;
        flags = 0;
    }
    const int min_match = blk.data_type == 6 /*DT_DNA*/ ? 6 : 4;
    flags |= (uint8_t)(((min_match - 2) & 7) << 1);
    if (lane == 0) dst[12] = flags;
    const bool ex = extra != 0;
    // warp-uniform scalar state
    int src_idx = 0, dst_idx = 13, anchor = 0, mlen_idx = 0, m_idx = 0, tk_idx = 0, n_runs = 0;
    int repd0 = count, repd1 = count, repd_idx = 0, src_inc = 0;
    bool too_many = false;
    while (src_idx < src_end) {
        // ---- batch: my position
        int pj = src_idx;
        for (int i = 0; i < lane; i++) pj += 1 + ((src_inc + i) >> 6);
        const bool live = pj < src_end;
        const uint32_t live_mask = __ballot_sync(0xFFFFFFFFu, live);  // a prefix of the lanes (positions grow with the lane)
        const int nl = __popc(live_mask);
        uint64_t p = 0;
        uint32_t h = 0xFFFFFFFFu - (uint32_t)lane;  // dead lanes: distinct dummy slots
        int ref0 = 0;
        if (live) {
            p = ld64(w, pj);
            h = lz_hash(p, ex);
            ref0 = hashes[h];
        }
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, h) & live_mask;
        const uint32_t earlier = same & ((1u << lane) - 1u);
        const int from = earlier ? 31 - __clz((int)earlier) : lane;
        const int fwd = __shfl_sync(0xFFFFFFFFu, pj, from);
        if (earlier) ref0 = fwd;
        // ---- probe
        int best_len = 0, ref = 0;
        bool hit = false, rep = false;
        if (live) {
            const int rj = lane == 0 ? repd_idx : 0;
            const int pj1 = pj + 1;
            const int max_match = min(src_end - pj1, LZ_MAX_MATCH);
            const int min_ref = max(pj - max_dist, 0);
            ref = pj1 - (rj ? repd1 : repd0);
            if (ref > min_ref && (uint32_t)(p >> 8) == ld32(w, ref)) {
                best_len = find_match(w, pj1, ref, max_match);
            } else {
                ref = pj1 - (rj ? repd0 : repd1);
                if (ref > min_ref && (uint32_t)(p >> 8) == ld32(w, ref)) best_len = find_match(w, pj1, ref, max_match);
            }
            if (best_len >= min_match) {
                hit = true;
                rep = true;
            } else {
                ref = ref0;
                if (ref > min_ref && (uint32_t)p == ld32(w, ref)) {
                    best_len = find_match(w, pj, ref, min(src_end - pj, LZ_MAX_MATCH));
                    hit = best_len >= min_match;
                }
            }
        }
        const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
        const int wl = hits ? __ffs((int)hits) - 1 : -1;
        // ---- commit the inserts of the lanes up to the winner: per slot, the highest such lane writes
        const uint32_t commit_mask = wl < 0 ? live_mask : (live_mask & ((2u << wl) - 1u));
        if ((commit_mask >> lane) & 1u) {
            const uint32_t later = same & commit_mask & ~((2u << lane) - 1u);
            if (!later) hashes[h] = pj;
        }
        __syncwarp();
        if (wl < 0) {  // every live lane missed
            const int last = __shfl_sync(0xFFFFFFFFu, pj, nl - 1);
            src_idx = last + 1 + ((src_inc + nl - 1) >> 6);
            src_inc += nl;
            repd_idx = 0;
            continue;
        }
        src_inc += wl;
        if (wl > 0) repd_idx = 0;
        src_idx = __shfl_sync(0xFFFFFFFFu, pj, wl);
        best_len = __shfl_sync(0xFFFFFFFFu, best_len, wl);
        ref = __shfl_sync(0xFFFFFFFFu, ref, wl);
        rep = __shfl_sync(0xFFFFFFFFu, (int)rep, wl) != 0;
        // ---- the match, warp-uniform; lane 0 stores
        const int src_idx1 = src_idx + 1;
        const int max_match = min(src_end - src_idx1, LZ_MAX_MATCH);
        const int min_ref = max(src_idx - max_dist, 0);
        if (!rep) {
            if (ref != src_idx - repd0 && ref != src_idx - repd1) {
                const uint32_t h1 = lz_hash(ld64(w, src_idx1), ex);
                const int ref1 = hashes[h1];
                __syncwarp();
                if (lane == 0) hashes[h1] = src_idx1;
                if (ref1 > min_ref + 1 && ld32(w, src_idx1 + best_len - 3) == ld32(w, ref1 + best_len - 3)) {
                    const int bl1 = find_match(w, src_idx1, ref1, max_match);
                    if (bl1 >= best_len) {
                        ref = ref1;
                        best_len = bl1;
                        src_idx = src_idx1;
                    }
                }
                if (ex) {
                    const int src_idx2 = src_idx1 + 1;
                    const uint32_t h2 = lz_hash(ld64(w, src_idx2), ex);
                    __syncwarp();
                    const int ref2 = hashes[h2];  // after the insert of h1 (the two slots may coincide)
                    __syncwarp();
                    if (lane == 0) hashes[h2] = src_idx2;
                    if (ref2 > min_ref + 2 && ld32(w, src_idx2 + best_len - 3) == ld32(w, ref2 + best_len - 3)) {
                        const int bl2 = find_match(w, src_idx2, ref2, min(src_end - src_idx2, LZ_MAX_MATCH));
                        if (bl2 >= best_len) {
                            ref = ref2;
                            best_len = bl2;
                            src_idx = src_idx2;
                        }
                    }
                }
                __syncwarp();
            }
            while (src_idx > anchor && ref > min_ref && src[src_idx - 1] == src[ref - 1]) {
                best_len++;
                ref--;
                src_idx--;
            }
            if (best_len > LZ_MAX_MATCH) {
                src_idx += best_len - LZ_MAX_MATCH;
                ref += best_len - LZ_MAX_MATCH;
                best_len = LZ_MAX_MATCH;
            }
        } else {
            if (src[src_idx] == src[ref - 1] && best_len < LZ_MAX_MATCH) {
                best_len++;
                ref--;
            } else {
                src_idx++;
                if (lane == 0) hashes[lz_hash(ld64(w, src_idx), ex)] = src_idx;
                __syncwarp();
            }
        }
        src_inc = 0;
        const int dist = src_idx - ref;
        const int m_len = best_len - min_match;
        int token, m_len_th;
        if (dist == repd0) {
            token = 0x00;
            m_len_th = 3;
        } else if (dist == repd1) {
            token = 0x04;
            m_len_th = 3;
        } else {
            m_len_th = 7;
            if (dist >= 256) {
                if (dist >= 65536) {
                    if (lane == 0) {
                        m_buf[m_idx] = (uint8_t)(dist >> 16);
                        m_buf[m_idx + 1] = (uint8_t)(dist >> 8);
                    }
                    m_idx += 2;
                    token = 0x18;
                } else {
                    if (lane == 0) m_buf[m_idx] = (uint8_t)(dist >> 8);
                    m_idx++;
                    token = 0x10;
                }
            } else {
                token = 0x08;
            }
            if (lane == 0) m_buf[m_idx] = (uint8_t)dist;
            m_idx++;
        }
        if (m_len >= m_len_th) {
            token += m_len_th;
            if (lane == 0) emit_length(mlen_buf + mlen_idx, m_len - m_len_th);
            mlen_idx += length_bytes(m_len - m_len_th);
        } else {
            token += m_len;
        }
        repd1 = repd0;
        repd0 = dist;
        repd_idx = 1;
        const int lit_len = src_idx - anchor;
        if (lit_len == 0) {
            if (lane == 0) tk_buf[tk_idx] = (uint8_t)token;
            tk_idx++;
        } else {
            if (lit_len >= 7) {
                if (lit_len >= (1 << 24)) {
                    too_many = true;
                    break;
                }
                if (lane == 0) {
                    tk_buf[tk_idx] = (uint8_t)((7 << 5) | token);
                    emit_length(dst + dst_idx, lit_len - 7);
                }
                tk_idx++;
                dst_idx += length_bytes(lit_len - 7);
            } else {
                if (lane == 0) tk_buf[tk_idx] = (uint8_t)((lit_len << 5) | token);
                tk_idx++;
            }
            if (lane == 0) {
                runs[n_runs].src = (uint32_t)anchor;
                runs[n_runs].dst = (uint32_t)dst_idx;
                runs[n_runs].len = (uint32_t)lit_len;
            }
            n_runs++;
            dst_idx += lit_len;
        }
        anchor = src_idx + best_len;
        // positions inside the match: 32 at a time; the serial order leaves the last (largest) position of a slot
        for (int q = src_idx + 1 + lane; q < anchor; q += 32) atomicMax(&hashes[lz_hash(ld64(w, q), ex)], q);
        __syncwarp();
        src_idx = anchor;
    }
    if (!too_many && lane == 0) {
        const int lit_len = count - anchor;
        if (dst_idx + lit_len + tk_idx + m_idx < count) {
            if (lit_len >= 7) {
                tk_buf[tk_idx++] = (uint8_t)(7 << 5);
                dst_idx += emit_length(dst + dst_idx, lit_len - 7);
            } else {
                tk_buf[tk_idx++] = (uint8_t)(lit_len << 5);
            }
            runs[n_runs].src = (uint32_t)anchor;
            runs[n_runs].dst = (uint32_t)dst_idx;
            runs[n_runs].len = (uint32_t)lit_len;
            n_runs++;
            dst_idx += lit_len;
            const uint32_t a = (uint32_t)dst_idx, bb = (uint32_t)tk_idx, c = (uint32_t)m_idx;
            for (int i = 0; i < 4; i++) {
                dst[i] = (uint8_t)(a >> (8 * i));
                dst[4 + i] = (uint8_t)(bb >> (8 * i));
                dst[8 + i] = (uint8_t)(c >> (8 * i));
            }
            const int total = dst_idx + tk_idx + m_idx + mlen_idx;
            if (total <= count - count / 100) {
                r.status = 0;
                r.out_len = (uint32_t)total;
            }
            r.n_runs = (uint32_t)n_runs;
            r.lit_end = (uint32_t)dst_idx;
            r.tk_n = (uint32_t)tk_idx;
            r.m_n = (uint32_t)m_idx;
            r.mlen_n = (uint32_t)mlen_idx;
        }
    }
    if (lane == 0) res[b] = r;
}

}  // namespace lzspec
}  // namespace kz
