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
// When the skip acceleration fires (64 probes in a row without a match — at the start of every block, and in incompressible regions)
// positions are left out of the table. Speculative parses stop there; the true parse (stitch pass) goes on in "skip mode": a bitmap holds
// the skipped positions and a candidate that points to one is replaced by the next inserted position of the prev[] chain. A skipped
// position j falsifies exactly one stored candidate, the one of next[j] (the next position with j's hash): that position is marked
// "tainted", and a speculative log is only taken up to the first match whose literal run or match bytes contain a tainted position; the
// true parse steps over it and merges again behind it.
#pragma once
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
    uint32_t start, len, dist, pad;
};
struct State {
    int32_t src_idx, anchor, repd0, repd1, repd_idx, src_inc;
};
struct Params {
    const uint8_t* src;
    int32_t count, src_end, max_dist, min_match;
    int32_t extra;
    const Rec* rec;
    const uint32_t* next;  // next position with the same hash (NONE at the end of a chain)
    uint32_t* skipmap;     // one bit per position, all zero until the skip acceleration fires
    uint32_t* taintmap;    // one bit per position whose stored candidate is a skipped position
    uint32_t* seg_taint;   // [0] = number of tainted positions of the block (merging stops when it exceeds taint_cap)
    int32_t seg_size;
    uint32_t taint_cap;
};
// capacity of the piece list of a block: every speculative piece is a whole segment log or ends at a tainted position
KZL_HD uint32_t piece_capacity(uint32_t nsegs, uint32_t taint_cap) { return 2u * (nsegs + taint_cap) + 16u; }
enum { R_MATCH = 0, R_END = 1, R_SKIP = 2 };
enum { END_POST_MATCH = 0, END_BLOCK = 1, END_STOP_SKIP = 2 };
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

// ---- per position precomputation (position i, candidate c = prev[i]) -------------------------------------------------------------
KZL_HD uint32_t make_info(const uint8_t* src, int count, int i, uint32_t c) {
    if (c == NONE) return 0;
    const int t = common_prefix(src, count, i, (int)c, 0, T_CAP);
    const int b = common_suffix(src, i, (int)c, 0, B_CAP);
    return (uint32_t)t | ((uint32_t)b << 8);
}

KZL_HD bool skipped(const Params& P, int pos) { return (P.skipmap[pos >> 5] >> (pos & 31)) & 1u; }

// the table entry a probe at position i reads: cand (NONE / 0 = nothing usable), its prefix length t and suffix length b (saturated)
KZL_HD void lookup(const Params& P, int i, bool skipmode, int& cand, int& t, int& b) {
    const Rec r = P.rec[i];
    uint32_t c = r.prev;
    t = (int)(r.info & 0xFF);
    b = (int)((r.info >> 8) & 0xFF);
    if (skipmode && c != NONE && skipped(P, (int)c)) {
        do c = P.rec[c].prev;
        while (c != NONE && skipped(P, (int)c));
        if (c != NONE) {
            const uint32_t info = make_info(P.src, P.count, i, c);
            t = (int)(info & 0xFF);
            b = (int)((info >> 8) & 0xFF);
        }
    }
    cand = c == NONE ? 0 : (int)c;  // position 0 is never accepted (ref > minRef >= 0), like an empty table slot
}
// true prefix length as far as `limit` matters (the stored one saturates at T_CAP)
KZL_HD int full_t(const Params& P, int i, int cand, int t, int limit) {
    if (t < T_CAP || limit <= T_CAP) return t;
    return common_prefix(P.src, P.count, i, cand, T_CAP & ~7, limit + 8);
}
KZL_HD int full_b(const Params& P, int i, int cand, int b, int limit) {
    if (b < B_CAP || limit <= B_CAP) return b;
    return common_suffix(P.src, i, cand, B_CAP, limit);
}
// what findMatchLZX(src, i, cand, maxm) returns, from the prefix length
KZL_HD int match_len(const Params& P, int i, int cand, int t, int maxm) {
    const int m8 = maxm & ~7;
    const int tt = full_t(P, i, cand, t, m8);
    return tt < m8 ? tt : m8;
}

KZL_HD void mark_skipped(const Params& P, int from, int to) {  // [from, to)
    for (int p = from; p < to; p++) {
        P.skipmap[p >> 5] |= 1u << (p & 31);
        const uint32_t t = P.next[p];
        if (t != NONE && !((P.taintmap[t >> 5] >> (t & 31)) & 1u)) {
            P.taintmap[t >> 5] |= 1u << (t & 31);
            P.seg_taint[0]++;
        }
    }
}
// first tainted position in [from, to), or -1
KZL_HD int next_tainted(const Params& P, int from, int to) {
    if (from >= to) return -1;
    int w = from >> 5;
    const int wl = (to - 1) >> 5;
    uint32_t bits = P.taintmap[w] & (0xFFFFFFFFu << (from & 31));
    for (;;) {
        if (bits) {
            const int pos = (w << 5) + ctz64((uint64_t)bits);
            return pos < to ? pos : -1;
        }
        if (++w > wl) return -1;
        bits = P.taintmap[w];
    }
}
KZL_HD void clear_skipped(const Params& P, int from, int to) {  // [from, to): these positions are inserted by the in-match loop
    if (from >= to) return;
    int p = from;
    while (p < to && (p & 31)) {
        P.skipmap[p >> 5] &= ~(1u << (p & 31));
        p++;
    }
    while (p + 32 <= to) {
        P.skipmap[p >> 5] = 0;
        p += 32;
    }
    while (p < to) {
        P.skipmap[p >> 5] &= ~(1u << (p & 31));
        p++;
    }
}

// Runs the parse from `st` to the next match. R_MATCH: m is the match, st the state behind it. R_END: the cursor reached srcEnd (st.src_idx
// >= src_end). R_SKIP (only when !skipmode): the probe at st.src_idx misses and would be followed by a stride > 1 — st is the state BEFORE
// that probe, so that the caller can repeat it in skip mode.
KZL_HD int parse_one(const Params& P, State& st, bool skipmode, Match& m) {
    const uint8_t* src = P.src;
    const int src_end = P.src_end, min_match = P.min_match;
    while (st.src_idx < src_end) {
        int src_idx = st.src_idx;
        const int src_idx1 = src_idx + 1;
        const int max_match = imin(src_end - src_idx1, MAX_MATCH);
        const int min_ref = imax(src_idx - P.max_dist, 0);
        const uint64_t p = ld64(src + src_idx);
        int best_len = 0;
        int ref = src_idx1 - (st.repd_idx ? st.repd1 : st.repd0);
        if (ref > min_ref && (uint32_t)(p >> 8) == ld32(src + ref)) {
            best_len = find_match(src, src_idx1, ref, max_match);
        } else {
            ref = src_idx1 - (st.repd_idx ? st.repd0 : st.repd1);
            if (ref > min_ref && (uint32_t)(p >> 8) == ld32(src + ref)) best_len = find_match(src, src_idx1, ref, max_match);
        }
        if (best_len < min_match) {
            int cand, t, b;
            lookup(P, src_idx, skipmode, cand, t, b);
            bool found = false;
            if (cand > min_ref && t >= 4) {
                best_len = match_len(P, src_idx, cand, t, imin(src_end - src_idx, MAX_MATCH));
                found = best_len >= min_match;
            }
            if (!found) {
                const int stride = st.src_inc >> 6;
                if (stride > 0) {
                    if (!skipmode) return R_SKIP;
                    mark_skipped(P, src_idx1, imin(src_idx1 + stride, P.count));
                }
                st.src_idx = src_idx1 + stride;
                st.src_inc++;
                st.repd_idx = 0;
                continue;
            }
            ref = cand;
            int bsel = b, isel = src_idx;  // suffix length / position of the selected candidate pair
            if (ref != src_idx - st.repd0 && ref != src_idx - st.repd1) {
                int c1, t1, b1;
                lookup(P, src_idx1, skipmode, c1, t1, b1);
                if (c1 > min_ref + 1) {
                    const int m8 = max_match & ~7;
                    if (m8 >= best_len && full_t(P, src_idx1, c1, t1, best_len + 1) >= best_len + 1) {
                        const int tt = full_t(P, src_idx1, c1, t1, m8);
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
                    lookup(P, src_idx2, skipmode, c2, t2, b2);
                    if (c2 > min_ref + 2) {
                        const int m8 = imin(src_end - src_idx2, MAX_MATCH) & ~7;
                        if (m8 >= best_len && full_t(P, src_idx2, c2, t2, best_len + 1) >= best_len + 1) {
                            const int tt = full_t(P, src_idx2, c2, t2, m8);
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
                    const int ext = imin(full_b(P, isel, ref, bsel, room), room);
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
        m.pad = 0;
        st.repd1 = st.repd0;
        st.repd0 = dist;
        st.repd_idx = 1;
        st.src_inc = 0;
        st.anchor = src_idx + best_len;
        if (skipmode) clear_skipped(P, src_idx + 1, st.anchor);
        st.src_idx = st.anchor;
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
        const int r = parse_one(P, st, false, m);
        if (r == R_END) {
            end = END_BLOCK;
            break;
        }
        if (r == R_SKIP) {
            end = END_STOP_SKIP;
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

// Stitch pass over one block. seg_log(k) = speculative log of segment k, desc[k] its result. Writes the fix-up log and the piece list.
// Returns the number of pieces; *fix_n = fix-up entries; *final_anchor = anchor behind the last match (start of the final literals).
template <class SegLog>
KZL_HD uint32_t stitch_block(const Params& P, int nsegs, const SegDesc* desc, SegLog seg_log, Match* fix, Piece* pieces, uint32_t* fix_n, int32_t* final_anchor) {
    const int seg_size = P.seg_size;
    State st = initial_state(P.count, 0);
    bool skipmode = false;
    uint32_t np = 0, fn = 0, fbegin = 0;
    bool post = true;    // st stands behind a match (or at the block start): a merge is possible
    bool first = true;   // block start: segment 0 was parsed from exactly this state
    int cur_seg = -1;
    uint32_t cursor = 0;
    while (st.src_idx < P.src_end) {
        if (post && P.seg_taint[0] <= P.taint_cap) {
            const int k = imin(st.src_idx / seg_size, nsegs - 1);
            if (k != cur_seg) {
                cur_seg = k;
                cursor = 0;
            }
            const SegDesc& d = desc[k];
            const Match* lg = seg_log(k);
            int j = -2;  // -2: no merge; -1: the whole log (block start); >= 0: entries behind j
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
                // how much of the log behind j is free of tainted positions (their stored candidates are wrong)?
                int e = (int)d.n;  // first entry that is NOT taken
                bool whole = true;
                if (skipmode) {
                    const int cover_end = (int)d.n > j + 1 ? (int)(lg[d.n - 1].start + lg[d.n - 1].len) : st.anchor;
                    // the speculative parse may also have probed up to 66 positions behind its last match (END_STOP_SKIP / END_BLOCK)
                    const int probe_end = d.end == END_POST_MATCH ? cover_end : imin(imax(cover_end, d.fin.src_idx) + 3, P.count);
                    const int t = next_tainted(P, st.anchor, probe_end);
                    if (t >= 0) {
                        whole = false;
                        e = j + 1;
                        while (e < (int)d.n && (int)(lg[e].start + lg[e].len) <= t) e++;
                        // entry e (or the tail of the parse behind the last entry) looked at position t
                    }
                }
                if (whole || e > j + 1) {
                    if (fn > fbegin) {
                        pieces[np].spec = 0, pieces[np].seg = 0, pieces[np].begin = fbegin, pieces[np].end = fn;
                        np++;
                        fbegin = fn;
                    }
                    if (e > j + 1) {
                        pieces[np].spec = 1, pieces[np].seg = (uint32_t)k, pieces[np].begin = (uint32_t)(j + 1), pieces[np].end = (uint32_t)e;
                        np++;
                    }
                    if (whole) {  // the state behind the segment's log = the state its speculative parse stopped in
                        st = d.fin;
                        post = d.end == END_POST_MATCH;
                        if (d.end == END_BLOCK) break;
                        continue;
                    }
                    // state behind entry e - 1
                    st.repd1 = e - 2 >= 0 ? (int)lg[e - 2].dist : P.count;
                    st.repd0 = (int)lg[e - 1].dist;
                    st.anchor = (int)(lg[e - 1].start + lg[e - 1].len);
                    st.src_idx = st.anchor;
                    st.repd_idx = 1;
                    st.src_inc = 0;
                    cursor = (uint32_t)(e - 1);
                }
                // fall through: one step of the true parse over the tainted position
            }
        }
        Match m;
        const int r = parse_one(P, st, skipmode, m);
        if (r == R_SKIP) {
            skipmode = true;
            post = false;
            continue;
        }
        if (r == R_END) break;
        fix[fn++] = m;
        post = true;
    }
    if (fn > fbegin) {
        pieces[np].spec = 0, pieces[np].seg = 0, pieces[np].begin = fbegin, pieces[np].end = fn;
        np++;
    }
    *fix_n = fn;
    *final_anchor = st.anchor;
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
