// Data-parallel LZ / LZX forward parse, host/device shared core (kz_lz_par.cu holds the kernels, tests/host/lz_par_host.cpp runs the
// same code on the CPU against the oracle).
//
// Reference: v2/transform/LZCodec.go:249-591 (LZXCodec.Forward), hash :238-244, findMatchLZX :593-608. The reference parse is a greedy
// state machine over a hash table of "last position with this hash". Two observations turn it into parallel work:
//
//  1. As long as the skip acceleration (srcInc >> 6, :358) has not fired, EVERY position below the cursor has been inserted into the table
//     in ascending order (literal probes :326, lazy probes :368 / :385, the rep branch :411-413 and the in-match loops :497-531 together
//     leave no gap), so the table entry a probe at position i reads is a pure function of the data:
//         prev[i] = max { j < i : hash(j) == hash(i) }
//     prev[], the common prefix length T[i] of (i, prev[i]) and the common suffix length B[i] before them are computed for all positions
//     at once (a stable sort by hash, then streaming compares). What findMatchLZX returns is a function of T and the position:
//         find_match(i, prev[i], maxm) = T < (maxm & ~7) ? T : (maxm & ~7)
//     and the lazy checks at +1 / +2 (:371-381, :388-398: four bytes at offset bestLen - 3, then length >= bestLen) are
//         T1 >= bestLen + 1  and  (maxm1 & ~7) >= bestLen.
//     Only the two repeat-distance probes (:339-347) read the data at parse time.
//  2. The parser state after a match is (anchor, repd[0], repd[1]) — srcInc = 0, repdIdx = 1 — and the table is not part of it (1.), so a
//     parse started anywhere with a guessed state is identical to the true parse from the first moment both stand behind a match with
//     the same anchor and the same last two distances. A block is cut into segments; every segment is parsed speculatively from its
//     first byte by its own warp (all segments of all blocks at once); a stitch pass then walks the block once: it continues the true
//     parse into a segment only until it meets that segment's speculative log (a few matches), and takes the rest of the log as is.
//
//  3. The skip acceleration (64 probes in a row without a match: at the start of every block, in incompressible regions) leaves positions
//     out of the table, and a backward-extended match puts some of them back later (:403-407, :497-531). Which positions, and from when on,
//     is a function of the parse itself: vf[j] = 0 (inserted when the cursor passed), NEVER, or the end of the match that inserted j
//     late (usable for probes at or behind that point). So the parse is the fixed point of
//         vf -> candidates filtered by vf (first usable position of the prev[] chain, T / B recomputed) -> parse -> vf'
//     and that fixed point is unique (induction over the position of the first difference). Round 0 starts from vf = 0; every round is
//     the same data-parallel pipeline (filter, speculative segments, stitch, vf' per position with a closed form for the probe pattern
//     of a literal run: group g of 64 probes has stride g + 1 and starts 32 g (g + 1) bytes behind the run's first byte).
//     Blocks converge after two or three rounds in practice; one that has not after the last round goes to the serial parse of kz_lz.cu.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define KZL_HD __host__ __device__ __forceinline__
#else
#define KZL_HD inline
#endif

namespace kz {
namespace lzp {

const int MAX_DISTANCE1 = (1 << 16) - 2;
const int MAX_DISTANCE2 = (1 << 24) - 2;
const int MAX_MATCH = 65535 + 254 + 4;
const int MIN_BLOCK = 24;
const uint32_t NONE = 0xFFFFFFFFu;
const int T_CAP = 255;  // stored prefix / suffix lengths saturate here; the parse extends them when a decision needs more
const int B_CAP = 255;
const uint64_t HASH_SEED = 0x1E35A7BDull;

struct Rec {  // per position: nearest earlier position with the same hash (all positions inserted), T | B << 8
    uint32_t prev, info;
};
struct Match {
    uint32_t start, len, dist, probe;  // probe = position of the probe that found the match
};
struct State {
    int32_t src_idx, anchor, repd0, repd1, repd_idx, src_inc;
};
struct Params {
    const uint8_t* src;
    int32_t count, src_end, max_dist, min_match;
    int32_t extra;
    const Rec* rec;  // candidates of this round (filtered by the previous round's vf)
    int32_t seg_size;
};
const uint32_t VF_NEVER = 0xFFFFFFFFu;
enum { R_MATCH = 0, R_END = 1 };
enum { END_POST_MATCH = 0, END_BLOCK = 1 };
struct SegDesc {  // result of the speculative parse of one segment
    uint32_t n;    // log entries
    uint32_t end;  // END_*
    State fin;     // state where the parse stopped
};
struct Piece {  // the final match list of a block is a sequence of pieces
    uint32_t spec;  // 0: entries [begin, end) of the block's fix-up log; 1: entries [begin, end) of segment `seg`'s speculative log
    uint32_t seg, begin, end;
};

KZL_HD uint64_t ld64(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
    const uint64_t* w = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(7));
    const int s = (int)(reinterpret_cast<uintptr_t>(p) & 7) << 3;
    const uint64_t a = w[0];
    if (s == 0) return a;
    return (a >> s) | (w[1] << (64 - s));
#else
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
#endif
}
KZL_HD uint32_t ld32(const uint8_t* p) { return (uint32_t)ld64(p); }
KZL_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
KZL_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
KZL_HD uint32_t lz_hash(uint64_t v, bool extra) { return (uint32_t)(((v << 24) * HASH_SEED) >> (64 - (extra ? 19 : 16))); }
KZL_HD int imin(int a, int b) { return a < b ? a : b; }
KZL_HD int imax(int a, int b) { return a > b ? a : b; }

// findMatchLZX :593-608 on the data itself (the repeat-distance probes)
KZL_HD int find_match(const uint8_t* src, int a, int b, int maxm) {
    int best = 0;
    while (best + 8 <= maxm) {
        const uint64_t diff = ld64(src + a + best) ^ ld64(src + b + best);
        if (diff != 0) {
            best += ctz64(diff) >> 3;
            break;
        }
        best += 8;
    }
    return best;
}

// common prefix of src[a..] and src[b..] (b < a), counted from offset `from` (the bytes below it are known to be equal), at most `limit`,
// never past the end of the block
KZL_HD int common_prefix(const uint8_t* src, int count, int a, int b, int from, int limit) {
    limit = imin(limit, count - a);
    int t = from;
    while (t < limit) {
        const uint64_t diff = ld64(src + a + t) ^ ld64(src + b + t);
        if (diff != 0) {
            t += ctz64(diff) >> 3;
            break;
        }
        t += 8;
    }
    return imin(t, limit);
}
// number of equal bytes going backwards from (a - 1, b - 1), at most `limit` and at most b
KZL_HD int common_suffix(const uint8_t* src, int a, int b, int from, int limit) {
    limit = imin(limit, b);
    int t = from;
    while (t < limit && src[a - 1 - t] == src[b - 1 - t]) t++;
    return t;
}

#if defined(__CUDACC__)
// Warp-uniform versions (every lane passes the same arguments): 32 x 8 bytes per step forwards, 32 bytes per step backwards. A match of
// MAX_MATCH bytes costs 257 steps instead of 8 K dependent loads; the speculative parse meets such matches in every segment they cover.
__device__ __forceinline__ int common_prefix_w(const uint8_t* src, int count, int a, int b, int from, int limit) {
    limit = imin(limit, count - a);
    const int lane = threadIdx.x & 31;
    int t = from;
    while (t < limit) {
        const int o = t + 8 * lane;
        uint64_t diff = 0;
        if (o < limit) diff = ld64(src + a + o) ^ ld64(src + b + o);
        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, diff != 0);
        if (mm) {
            const int l = __ffs((int)mm) - 1;
            const uint64_t d = __shfl_sync(0xFFFFFFFFu, diff, l);
            t += 8 * l + (ctz64(d) >> 3);
            break;
        }
        t += 256;
    }
    return imin(t, limit);
}
__device__ __forceinline__ int common_suffix_w(const uint8_t* src, int a, int b, int from, int limit) {
    limit = imin(limit, b);
    const int lane = threadIdx.x & 31;
    int t = from;
    while (t < limit) {
        const int o = t + lane;
        const bool ne = o < limit && src[a - 1 - o] != src[b - 1 - o];
        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, ne);
        if (mm) {
            t += __ffs((int)mm) - 1;
            break;
        }
        t += 32;
    }
    return imin(t, limit);
}
// find_match continued from offset `from` (a multiple of 8 below which the two strings are equal)
__device__ __forceinline__ int find_match_w(const uint8_t* src, int a, int b, int from, int maxm) {
    const int lane = threadIdx.x & 31;
    int best = from;
    while (best + 8 <= maxm) {
        const int o = best + 8 * lane;
        uint64_t diff = 0;
        if (o + 8 <= maxm) diff = ld64(src + a + o) ^ ld64(src + b + o);
        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, diff != 0);
        if (mm) {
            const int l = __ffs((int)mm) - 1;
            const uint64_t d = __shfl_sync(0xFFFFFFFFu, diff, l);
            return best + 8 * l + (ctz64(d) >> 3);
        }
        best += 256;
    }
    return imin(best, maxm & ~7);
}
#endif

// ---- per position precomputation (position i, candidate c = prev[i]) -------------------------------------------------------------
KZL_HD uint32_t make_info(const uint8_t* src, int count, int i, uint32_t c) {
    if (c == NONE) return 0;
    const int t = common_prefix(src, count, i, (int)c, 0, T_CAP);
    const int b = common_suffix(src, i, (int)c, 0, B_CAP);
    return (uint32_t)t | ((uint32_t)b << 8);
}

// the table entry a probe at position i reads: cand (0 = nothing usable), its prefix length t and suffix length b (saturated)
KZL_HD void lookup(const Params& P, int i, int& cand, int& t, int& b) {
    const Rec r = P.rec[i];
    t = (int)(r.info & 0xFF);
    b = (int)((r.info >> 8) & 0xFF);
    cand = r.prev == NONE ? 0 : (int)r.prev;  // position 0 is never accepted (ref > minRef >= 0), like an empty table slot
}
// true prefix length as far as `limit` matters (the stored one saturates at T_CAP)
// W: called with the same arguments by all lanes of a warp (the uniform second stage of the warp-cooperative parse): the long compares are shared
template <bool W = false>
KZL_HD int full_t(const Params& P, int i, int cand, int t, int limit) {
    if (t < T_CAP || limit <= T_CAP) return t;
#if defined(__CUDA_ARCH__)
    if (W) return common_prefix_w(P.src, P.count, i, cand, T_CAP & ~7, limit + 8);
#endif
    return common_prefix(P.src, P.count, i, cand, T_CAP & ~7, limit + 8);
}
template <bool W = false>
KZL_HD int full_b(const Params& P, int i, int cand, int b, int limit) {
    if (b < B_CAP || limit <= B_CAP) return b;
#if defined(__CUDA_ARCH__)
    if (W) return common_suffix_w(P.src, i, cand, B_CAP, limit);
#endif
    return common_suffix(P.src, i, cand, B_CAP, limit);
}
// what findMatchLZX(src, i, cand, maxm) returns, from the prefix length
KZL_HD int match_len(const Params& P, int i, int cand, int t, int maxm) {
    const int m8 = maxm & ~7;
    const int tt = full_t(P, i, cand, t, m8);
    return tt < m8 ? tt : m8;
}

// candidate of position i under vf: the first position of the prev[] chain that is in the table when i is probed
KZL_HD Rec filter_candidate(const uint8_t* src, int count, const Rec* rec0, const uint32_t* vf, int i) {
    Rec r = rec0[i];
    uint32_t c = r.prev;
    bool moved = false;
    while (c != NONE) {
        const uint32_t v = vf[c];
        if (v == 0 || (uint32_t)i >= v) break;  // in the table
        c = rec0[c].prev;
        moved = true;
    }
    if (moved) {
        r.prev = c;
        r.info = make_info(src, count, i, c);
    }
    return r;
}
// is the byte at offset d of a literal run probed (true) or jumped over? The j-th probe of a run is followed by a stride of 1 + (j >> 6).
KZL_HD bool run_offset_probed(uint32_t d) {
    uint32_t g = 0;
    if (d >= 64) {  // largest g with 32 g (g + 1) <= d
        g = (uint32_t)((sqrt(1.0 + (double)d / 8.0) - 1.0) * 0.5);
        while (32ull * (g + 1) * (g + 2) <= d) g++;
        while (g > 0 && 32ull * g * (g + 1) > d) g--;
    }
    return (d - 32u * g * (g + 1)) % (g + 1) == 0;
}
// vf of position j given the block's final match list ml[0..nm) (ascending): see the header comment
// t = number of matches that end at or before j
KZL_HD uint32_t matches_ended(const Match* ml, uint32_t nm, int j) {
    uint32_t lo = 0, hi = nm;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((int)(ml[mid].start + ml[mid].len) <= j) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
KZL_HD uint32_t derive_vf_at(const Match* ml, uint32_t nm, int src_end, int j, uint32_t t) {
    const int a = t ? (int)(ml[t - 1].start + ml[t - 1].len) : 0;
    if (t == nm) {
        if (j >= src_end) return 0;
        return run_offset_probed((uint32_t)(j - a)) ? 0u : VF_NEVER;
    }
    const Match m = ml[t];
    if (j >= (int)m.probe) return 0;
    if (run_offset_probed((uint32_t)(j - a))) return 0;
    return j > (int)m.start ? m.start + m.len : VF_NEVER;
}
KZL_HD uint32_t derive_vf(const Match* ml, uint32_t nm, int src_end, int j) { return derive_vf_at(ml, nm, src_end, j, matches_ended(ml, nm, j)); }

// One probe of the parse at position src_idx, first stage (:339-366): the two repeat-distance checks (repd_first before repd_second), then the
// table entry. Returns true when a match of at least min_match was found: kind 0 = repeat distance, 1 = table (b = stored suffix length).
// A pure function of the position, the two distances and the candidates: the warp-cooperative parse evaluates 32 probes at once with it.
KZL_HD bool probe_first(const Params& P, int src_idx, int repd_first, int repd_second, int& best_len, int& ref, int& kind, int& b) {
    const uint8_t* src = P.src;
    const int src_end = P.src_end, min_match = P.min_match;
    const int src_idx1 = src_idx + 1;
    const int max_match = imin(src_end - src_idx1, MAX_MATCH);
    const int min_ref = imax(src_idx - P.max_dist, 0);
    const uint64_t p = ld64(src + src_idx);
    best_len = 0;
    kind = 0;
    b = 0;
    ref = src_idx1 - repd_first;
    if (ref > min_ref && (uint32_t)(p >> 8) == ld32(src + ref)) {
        best_len = find_match(src, src_idx1, ref, max_match);
    } else {
        ref = src_idx1 - repd_second;
        if (ref > min_ref && (uint32_t)(p >> 8) == ld32(src + ref)) best_len = find_match(src, src_idx1, ref, max_match);
    }
    if (best_len >= min_match) return true;
    int cand, t;
    lookup(P, src_idx, cand, t, b);
    if (cand > min_ref && t >= 4) {
        best_len = match_len(P, src_idx, cand, t, imin(src_end - src_idx, MAX_MATCH));
        if (best_len >= min_match) {
            ref = cand;
            kind = 1;
            return true;
        }
    }
    return false;
}
// Second stage of a probe that found a match (:368-414): lazy checks at +1 / +2 and backward extension for a table match, the one byte
// backward step of a repeat match. `st` is the state at the probe (st.src_idx = the probe's position); on return it stands behind the match.
KZL_HD void unpack_rec(const Rec& r, int& cand, int& t, int& b) {
    t = (int)(r.info & 0xFF);
    b = (int)((r.info >> 8) & 0xFF);
    cand = r.prev == NONE ? 0 : (int)r.prev;
}
// rec1 / rec2: the table entries of the two positions behind the probe (the lazy checks), loaded by the caller together with everything else
template <bool W = false>
KZL_HD void probe_finish(const Params& P, State& st, int best_len, int ref, int kind, int b, const Rec& rec1, const Rec& rec2, Match& m) {
    const uint8_t* src = P.src;
    const int src_end = P.src_end;
    int src_idx = st.src_idx;
    const int src_idx1 = src_idx + 1;
    const int max_match = imin(src_end - src_idx1, MAX_MATCH);
    const int min_ref = imax(src_idx - P.max_dist, 0);
    if (kind == 1) {
        int bsel = b, isel = src_idx;  // suffix length / position of the selected candidate pair
        if (ref != src_idx - st.repd0 && ref != src_idx - st.repd1) {
            int c1, t1, b1;
            unpack_rec(rec1, c1, t1, b1);
            if (c1 > min_ref + 1) {
                const int m8 = max_match & ~7;
                if (m8 >= best_len && full_t<W>(P, src_idx1, c1, t1, best_len + 1) >= best_len + 1) {
                    const int tt = full_t<W>(P, src_idx1, c1, t1, m8);
                    best_len = tt < m8 ? tt : m8;
                    ref = c1;
                    src_idx = src_idx1;
                    bsel = b1;
                    isel = src_idx1;
                }
            }
            if (P.extra) {
                const int src_idx2 = src_idx1 + 1;
                int c2, t2, b2;
                unpack_rec(rec2, c2, t2, b2);
                if (c2 > min_ref + 2) {
                    const int m8 = imin(src_end - src_idx2, MAX_MATCH) & ~7;
                    if (m8 >= best_len && full_t<W>(P, src_idx2, c2, t2, best_len + 1) >= best_len + 1) {
                        const int tt = full_t<W>(P, src_idx2, c2, t2, m8);
                        best_len = tt < m8 ? tt : m8;
                        ref = c2;
                        src_idx = src_idx2;
                        bsel = b2;
                        isel = src_idx2;
                    }
                }
            }
        }
        {  // extend backwards (:403-407)
            const int room = imin(src_idx - st.anchor, ref - min_ref);
            if (room > 0 && bsel > 0) {
                const int ext = imin(full_b<W>(P, isel, ref, bsel, room), room);
                best_len += ext;
                ref -= ext;
                src_idx -= ext;
            }
        }
        if (best_len > MAX_MATCH) {
            src_idx += best_len - MAX_MATCH;
            ref += best_len - MAX_MATCH;
            best_len = MAX_MATCH;
        }
    } else {
        if (src[src_idx] == src[ref - 1] && best_len < MAX_MATCH) {
            best_len++;
            ref--;
        } else {
            src_idx++;
        }
    }
    const int dist = src_idx - ref;
    m.start = (uint32_t)src_idx;
    m.len = (uint32_t)best_len;
    m.dist = (uint32_t)dist;
    m.probe = (uint32_t)st.src_idx;
    st.repd1 = st.repd0;
    st.repd0 = dist;
    st.repd_idx = 1;
    st.src_inc = 0;
    st.anchor = src_idx + best_len;
    st.src_idx = st.anchor;
}
// Runs the parse from `st` to the next match. R_MATCH: m is the match, st the state behind it. R_END: the cursor reached srcEnd.
KZL_HD int parse_one(const Params& P, State& st, Match& m) {
    while (st.src_idx < P.src_end) {
        int best_len, ref, kind, b;
        if (!probe_first(P, st.src_idx, st.repd_idx ? st.repd1 : st.repd0, st.repd_idx ? st.repd0 : st.repd1, best_len, ref, kind, b)) {
            st.src_idx = st.src_idx + 1 + (st.src_inc >> 6);
            st.src_inc++;
            st.repd_idx = 0;
            continue;
        }
        probe_finish(P, st, best_len, ref, kind, b, P.rec[st.src_idx + 1], P.rec[st.src_idx + 2], m);
        return R_MATCH;
    }
    return R_END;
}

KZL_HD State initial_state(int count, int start) {
    State st;
    st.src_idx = start;
    st.anchor = start;
    st.repd0 = count;
    st.repd1 = count;
    st.repd_idx = 0;
    st.src_inc = 0;
    return st;
}

// Speculative parse of the segment [seg_start, seg_end): from a guessed state to the first match that ends at or behind seg_end.
KZL_HD void spec_parse_segment(const Params& P, int seg_start, int seg_end, Match* log, SegDesc& d) {
    State st = initial_state(P.count, seg_start);
    uint32_t n = 0;
    uint32_t end = END_BLOCK;
    for (;;) {
        Match m;
        const int r = parse_one(P, st, m);
        if (r == R_END) {
            end = END_BLOCK;
            break;
        }
        log[n++] = m;
        if (st.anchor >= seg_end) {
            end = END_POST_MATCH;
            break;
        }
    }
    d.n = n;
    d.end = end;
    d.fin = st;
}

// Stitch pass over one block. seg_log(k) = speculative log of segment k, desc[k] its result. Writes the fix-up log and the piece list
// (at most 2 * nsegs + 2 pieces). Returns the number of pieces; *fix_n = fix-up entries; *final_anchor = anchor behind the last match.
template <class SegLog>
KZL_HD uint32_t stitch_block(const Params& P, int nsegs, const SegDesc* desc, SegLog seg_log, Match* fix, Piece* pieces, uint32_t* fix_n, int32_t* final_anchor) {
    const int seg_size = P.seg_size;
    State st = initial_state(P.count, 0);
    uint32_t np = 0, fn = 0, fbegin = 0;
    bool first = true;  // block start: segment 0 was parsed from exactly this state
    int cur_seg = -1;
    uint32_t cursor = 0;
    while (st.src_idx < P.src_end) {
        // st stands behind a match (or at the block start): can it join the speculative log of the segment it is in?
        const int k = imin(st.src_idx / seg_size, nsegs - 1);
        if (k != cur_seg) {
            cur_seg = k;
            cursor = 0;
        }
        const SegDesc& d = desc[k];
        const Match* lg = seg_log(k);
        int j = -2;  // -2: no; -1: the whole log (block start); >= 0: the entries behind j
        if (first && k == 0) {
            j = -1;
        } else {
            while (cursor < d.n && (int)(lg[cursor].start + lg[cursor].len) < st.anchor) cursor++;
            if (cursor >= 1 && cursor < d.n && (int)(lg[cursor].start + lg[cursor].len) == st.anchor && (int)lg[cursor].dist == st.repd0 &&
                (int)lg[cursor - 1].dist == st.repd1)
                j = (int)cursor;
        }
        first = false;
        if (j != -2) {
            if (fn > fbegin) {
                pieces[np].spec = 0, pieces[np].seg = 0, pieces[np].begin = fbegin, pieces[np].end = fn;
                np++;
                fbegin = fn;
            }
            if ((int)d.n > j + 1) {
                pieces[np].spec = 1, pieces[np].seg = (uint32_t)k, pieces[np].begin = (uint32_t)(j + 1), pieces[np].end = d.n;
                np++;
            }
            st = d.fin;  // the state the segment's speculative parse stopped in
            if (d.end == END_BLOCK) break;
            continue;
        }
        Match m;
        if (parse_one(P, st, m) == R_END) break;
        fix[fn++] = m;
    }
    if (fn > fbegin) {
        pieces[np].spec = 0, pieces[np].seg = 0, pieces[np].begin = fbegin, pieces[np].end = fn;
        np++;
    }
    *fix_n = fn;
    *final_anchor = st.anchor;
    return np;
}

// ---- parallel stitch: one junction per segment -----------------------------------------------------------------------------------------
// The serial stitch above only ever does one thing between two joins: starting from the state a segment's speculative log stopped in, it
// parses on until it meets the log of the segment the cursor is in. That piece of work depends on nothing but the segment it starts from, so
// every segment's junction is parsed at once (junction_parse, one lane each) and what remains serial is a walk over the junction records
// (walk_junctions: a few loads per live segment). A junction that starts from a false state (its segment is never reached by the true parse:
// a long match jumps over it) may wander; its fix-up log is capped, and a capped junction that turns out to be live sends the block to the
// serial stitch.
enum { J_JOINED = 0, J_TERMINAL = 1, J_OVERFLOW = 2 };
struct Junction {
    uint32_t status;  // J_*
    uint32_t fix_n;   // fix-up matches parsed by this junction
    uint32_t seg;     // J_JOINED: the segment whose log is joined ...
    uint32_t idx;     // ... behind entry idx
    int32_t anchor;   // J_TERMINAL: anchor behind the last match of the block
};
template <class SegLog>
KZL_HD void junction_parse(const Params& P, int k0, int nsegs, const SegDesc* desc, SegLog seg_log, Match* fix, uint32_t fix_cap, Junction& J) {
    const int seg_size = P.seg_size;
    State st = desc[k0].fin;
    J.status = J_TERMINAL, J.fix_n = 0, J.seg = 0, J.idx = 0, J.anchor = st.anchor;
    if (desc[k0].end == END_BLOCK) return;  // the walk stops at this segment: no junction
    uint32_t fn = 0;
    int cur_seg = -1;
    uint32_t cursor = 0;
    while (st.src_idx < P.src_end) {
        const int k = imin(st.src_idx / seg_size, nsegs - 1);
        if (k != cur_seg) {
            cur_seg = k;
            cursor = 0;
        }
        const SegDesc& d = desc[k];
        const Match* lg = seg_log(k);
        while (cursor < d.n && (int)(lg[cursor].start + lg[cursor].len) < st.anchor) cursor++;
        if (cursor >= 1 && cursor < d.n && (int)(lg[cursor].start + lg[cursor].len) == st.anchor && (int)lg[cursor].dist == st.repd0 &&
            (int)lg[cursor - 1].dist == st.repd1) {
            J.status = J_JOINED, J.fix_n = fn, J.seg = (uint32_t)k, J.idx = cursor;
            return;
        }
        if (fn >= fix_cap) {
            J.status = J_OVERFLOW, J.fix_n = fn;
            return;
        }
        Match m;
        if (parse_one(P, st, m) == R_END) break;
        fix[fn++] = m;
    }
    J.status = J_TERMINAL, J.fix_n = fn, J.anchor = st.anchor;
}
// Walk over the junction records of one block. Pieces: spec 1 = entries [begin, end) of segment `seg`'s speculative log, spec 2 = entries
// [begin, end) of junction `seg`'s fix-up log. Returns the number of pieces (at most 2 * nsegs), or 0xFFFFFFFF when a live junction overflowed.
KZL_HD uint32_t walk_junctions(int nsegs, const SegDesc* desc, const Junction* junc, Piece* pieces, uint32_t* nmatch, int32_t* final_anchor) {
    uint32_t np = 0, total = 0;
    int k = 0;
    uint32_t begin = 0;
    for (;;) {
        const uint32_t n = desc[k].n;
        if (n > begin) {
            pieces[np].spec = 1, pieces[np].seg = (uint32_t)k, pieces[np].begin = begin, pieces[np].end = n;
            np++;
            total += n - begin;
        }
        if (desc[k].end == END_BLOCK) {
            *final_anchor = desc[k].fin.anchor;
            break;
        }
        const Junction J = junc[k];
        if (J.status == J_OVERFLOW) return 0xFFFFFFFFu;
        if (J.fix_n) {
            pieces[np].spec = 2, pieces[np].seg = (uint32_t)k, pieces[np].begin = 0, pieces[np].end = J.fix_n;
            np++;
            total += J.fix_n;
        }
        if (J.status == J_TERMINAL) {
            *final_anchor = J.anchor;
            break;
        }
        k = (int)J.seg;
        begin = J.idx + 1;
    }
    *nmatch = total;
    return np;
}

// ---- emission: sizes and bytes of one match of the final list (LZCodec.go:416-495) ---------------------------------------------------
KZL_HD int length_bytes(int length) { return length < 254 ? 1 : (length < 65536 + 254 ? 3 : 4); }
KZL_HD int emit_length(uint8_t* block, int length) {
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
struct Sizes {
    uint32_t m_bytes, mlen_bytes, lit_bytes;  // distance bytes, match length bytes, literal length bytes + literals
};
// prev_anchor = end of the previous match (0 for the first), repd0 / repd1 = distances of the previous two matches (count when absent)
KZL_HD void match_sizes(const Match& m, int prev_anchor, int repd0, int repd1, int min_match, Sizes& s, bool& too_many) {
    const int dist = (int)m.dist, m_len = (int)m.len - min_match;
    int th;
    if (dist == repd0 || dist == repd1) {
        th = 3;
        s.m_bytes = 0;
    } else {
        th = 7;
        s.m_bytes = dist >= 65536 ? 3 : (dist >= 256 ? 2 : 1);
    }
    s.mlen_bytes = m_len >= th ? (uint32_t)length_bytes(m_len - th) : 0;
    const int lit_len = (int)m.start - prev_anchor;
    too_many = lit_len >= (1 << 24);
    s.lit_bytes = (uint32_t)lit_len + (lit_len >= 7 ? (uint32_t)length_bytes(lit_len - 7) : 0u);
}
// writes the token, the distance bytes, the match length bytes and the literal LENGTH bytes (the literals themselves are copied by the
// caller: src[prev_anchor, start) -> lit_dst + (literal length bytes))
KZL_HD void match_emit(const Match& m, int prev_anchor, int repd0, int repd1, int min_match, uint8_t* tk, uint8_t* mb, uint8_t* mlenb, uint8_t* lit_dst) {
    const int dist = (int)m.dist, m_len = (int)m.len - min_match;
    int token, th;
    if (dist == repd0) {
        token = 0x00;
        th = 3;
    } else if (dist == repd1) {
        token = 0x04;
        th = 3;
    } else {
        th = 7;
        if (dist >= 256) {
            if (dist >= 65536) {
                mb[0] = (uint8_t)(dist >> 16);
                mb[1] = (uint8_t)(dist >> 8);
                mb += 2;
                token = 0x18;
            } else {
                mb[0] = (uint8_t)(dist >> 8);
                mb++;
                token = 0x10;
            }
        } else {
            token = 0x08;
        }
        mb[0] = (uint8_t)dist;
    }
    if (m_len >= th) {
        token += th;
        emit_length(mlenb, m_len - th);
    } else {
        token += m_len;
    }
    const int lit_len = (int)m.start - prev_anchor;
    if (lit_len >= 7) {
        tk[0] = (uint8_t)((7 << 5) | token);
        emit_length(lit_dst, lit_len - 7);
    } else {
        tk[0] = (uint8_t)((lit_len << 5) | token);
    }
}

}  // namespace lzp
}  // namespace kz
