// CPU harness of the data-parallel LZ / LZX parse: runs kanzi-go_b200/csrc/kz_lz_par_core.cuh (the code the kernels of kz_lz_par.cu run) pass by
// pass on the host, segments one after another, so that tests/test_lz_par_host.py can compare its output with the oracle byte for byte.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kanzi-go_b200/csrc/kz_lz_par_core.cuh"

using namespace kz::lzp;

static uint32_t g_jcap = 64;  // fix-up matches per junction (JCAP of kz_lz_par.cu)
extern "C" {
void lz_set_jcap(int v) { g_jcap = uint32_t(v); }
// returns the encoded length, -1 when the transform declines. stats: [0] matches, [1] fix-up matches (true parse work of the stitch passes, all
// rounds), [2] pieces, [3] rounds to the fixed point, [4] longest live junction, [5] rounds that needed the serial stitch
int64_t lz_forward_par(int extra_i, int data_type, const uint8_t* src_in, int64_t n, uint8_t* dst, int64_t cap, int seg_size, int64_t* stats) {
    const int max_rounds = 64;
    if (n == 0) return 0;
    const int count = int(n);
    if (cap < (n <= 1024 ? n + 16 : n + n / 64)) return -1;
    if (count < MIN_BLOCK) return -1;
    if (data_type == 9) return -1;
    std::vector<uint8_t> padded(size_t(n) + 64, 0);
    memcpy(padded.data(), src_in, size_t(n));
    const uint8_t* src = padded.data();
    Params P;
    P.src = src;
    P.count = count;
    P.src_end = count - 16 - 2;
    P.max_dist = MAX_DISTANCE2;
    uint8_t flags = 1;
    if (P.src_end < 4 * MAX_DISTANCE1) {
        P.max_dist = MAX_DISTANCE1;
        flags = 0;
    }
    P.min_match = data_type == 6 ? 6 : 4;
    flags |= uint8_t(((P.min_match - 2) & 7) << 1);
    P.extra = extra_i;
    // pass 0: prev[] (what the stable sort by hash produces), T / B
    const int npos = std::max(0, std::min(count - 8, P.src_end + 2));
    std::vector<Rec> rec0(size_t(count) + 8), rec(size_t(count) + 8);
    {
        std::vector<uint32_t> last(size_t(1) << (extra_i ? 19 : 16), NONE);
        for (int i = 0; i < count; i++) rec0[i].prev = NONE, rec0[i].info = 0;
        for (int i = 0; i < npos; i++) {
            const uint32_t h = lz_hash(ld64(src + i), extra_i != 0);
            rec0[i].prev = last[h];
            last[h] = uint32_t(i);
            rec0[i].info = make_info(src, count, i, rec0[i].prev);
        }
    }
    P.seg_size = seg_size;
    const int nsegs = std::max(1, (P.src_end + seg_size - 1) / seg_size);
    const int seg_cap = std::min(seg_size, count) / 4 + 8;  // a match is at least 4 bytes long
    std::vector<Match> logs(size_t(nsegs) * seg_cap);
    std::vector<SegDesc> desc(nsegs);
    std::vector<Match> fix(size_t(count) / 4 + 16);
    std::vector<Piece> pieces(size_t(2) * nsegs + 4);
    std::vector<uint32_t> vf(size_t(count) + 8, 0), vf2(size_t(count) + 8, 0);
    std::vector<Match> ml;
    std::vector<Junction> junc;
    std::vector<Match> jfix;
    uint32_t fix_n = 0, np = 0;
    int32_t final_anchor = 0;
    int rounds = 0;
    int64_t fix_total = 0, max_junction = 0, serial_stitches = 0;
    const uint32_t JCAP = g_jcap;
    for (;; rounds++) {
        if (rounds >= max_rounds) return -4;
        // filter the candidates by the previous round's vf (round 0: vf = 0, nothing moves); changed[k]: a candidate of segment k differs from
        // what the previous round's parse saw
        std::vector<uint8_t> changed(nsegs, rounds == 0 ? 1 : 0), dirty(nsegs, 1);
        for (int i = 0; i < count; i++) {
            const Rec r_new = rounds == 0 ? rec0[i] : filter_candidate(src, count, rec0.data(), vf.data(), i);
            if (rounds > 0 && (r_new.prev != rec[i].prev || r_new.info != rec[i].info)) changed[std::min(i / seg_size, nsegs - 1)] = 1;
            rec[i] = r_new;
        }
        P.rec = rec.data();
        // dirty[k] (kz_lz_par.cu: lzp_dirty_kernel): something changed in the segments the previous parse of segment k read
        if (rounds > 0) {
            for (int k = 0; k < nsegs; k++) {
                int k_end = nsegs - 1;
                if (desc[k].end != END_BLOCK) k_end = std::min(nsegs - 1, (std::max(desc[k].fin.anchor, 1) - 1) / seg_size);
                uint8_t any = 0;
                for (int m = k; m <= k_end && !any; m++) any = changed[m];
                dirty[k] = any;
            }
        }
        // speculative parse of every dirty segment (clean ones keep their log)
        for (int k = 0; k < nsegs; k++) {
            if (!dirty[k]) continue;
            const int s0 = k * seg_size;
            const int s1 = k == nsegs - 1 ? 0x7FFFFFFF : s0 + seg_size;
            spec_parse_segment(P, s0, s1, logs.data() + size_t(k) * seg_cap, desc[k]);
            if (int(desc[k].n) > seg_cap) return -2;
        }
        // stitch: one junction per segment (parallel on the GPU), then the walk over the junction records; a live junction that hit its
        // cap sends the block to the serial stitch. A junction is redone when a segment between its own and the one it joined is dirty.
        Match* lp = logs.data();
        auto seg_log = [lp, seg_cap](int k) { return (const Match*)(lp + size_t(k) * seg_cap); };
        if (junc.empty()) junc.resize(nsegs), jfix.resize(size_t(nsegs) * JCAP);
        for (int k = 0; k < nsegs; k++) {
            bool redo = rounds == 0 || junc[k].status != J_JOINED;
            if (!redo) {
                const int k_end = std::min<int>(int(junc[k].seg), nsegs - 1);
                for (int m = k; m <= k_end && !redo; m++) redo = dirty[m];
            }
            if (redo) junction_parse(P, k, nsegs, desc.data(), seg_log, jfix.data() + size_t(k) * JCAP, JCAP, junc[k]);
        }
        uint32_t nmatch = 0;
        np = walk_junctions(nsegs, desc.data(), junc.data(), pieces.data(), &nmatch, &final_anchor);
        const bool serial_stitch = np == 0xFFFFFFFFu;
        if (serial_stitch) {
            serial_stitches++;
            np = stitch_block(P, nsegs, desc.data(), seg_log, fix.data(), pieces.data(), &fix_n, &final_anchor);
        } else {
            fix_n = 0;
            for (uint32_t p = 0; p < np; p++)
                if (pieces[p].spec == 2) fix_n += pieces[p].end - pieces[p].begin, max_junction = std::max<int64_t>(max_junction, pieces[p].end - pieces[p].begin);
        }
        if (np > pieces.size()) return -5;
        fix_total += fix_n;
        // flatten
        ml.clear();
        for (uint32_t p = 0; p < np; p++) {
            const Match* base = pieces[p].spec == 1 ? logs.data() + size_t(pieces[p].seg) * seg_cap : pieces[p].spec == 2 ? jfix.data() + size_t(pieces[p].seg) * JCAP : fix.data();
            for (uint32_t e = pieces[p].begin; e < pieces[p].end; e++) ml.push_back(base[e]);
        }
        if (!serial_stitch && ml.size() != nmatch) return -6;
        // the table membership this parse implies
        bool same = true;
        for (int j = 0; j < count; j++) {
            vf2[j] = derive_vf(ml.data(), uint32_t(ml.size()), P.src_end, j);
            same = same && vf2[j] == vf[j];
        }
        if (same) break;
        vf.swap(vf2);
    }
    if (stats) {
        stats[0] = int64_t(ml.size());
        stats[1] = fix_total;
        stats[2] = np;
        stats[3] = rounds + 1;
        stats[4] = max_junction;
        stats[5] = serial_stitches;
    }
    std::vector<uint8_t> tk(ml.size() + 8), mb(3 * ml.size() + 8), mlenb(4 * ml.size() + 8);
    size_t tk_idx = 0, m_idx = 0, mlen_idx = 0, dst_idx = 13;
    int prev_anchor = 0, r0 = count, r1 = count;
    for (const Match& m : ml) {
        Sizes s;
        bool too_many;
        match_sizes(m, prev_anchor, r0, r1, P.min_match, s, too_many);
        if (too_many) return -1;
        match_emit(m, prev_anchor, r0, r1, P.min_match, &tk[tk_idx], &mb[m_idx], &mlenb[mlen_idx], dst + dst_idx);
        const int lit_len = int(m.start) - prev_anchor;
        memcpy(dst + dst_idx + (s.lit_bytes - lit_len), src + prev_anchor, size_t(lit_len));
        tk_idx++;
        m_idx += s.m_bytes;
        mlen_idx += s.mlen_bytes;
        dst_idx += s.lit_bytes;
        prev_anchor = int(m.start + m.len);
        r1 = r0;
        r0 = int(m.dist);
    }
    if (prev_anchor != final_anchor && !ml.empty()) return -3;
    const int lit_len = count - prev_anchor;
    if (dst_idx + size_t(lit_len) + tk_idx + m_idx >= size_t(count)) return -1;
    if (lit_len >= 7) {
        tk[tk_idx++] = uint8_t(7 << 5);
        dst_idx += size_t(emit_length(dst + dst_idx, lit_len - 7));
    } else {
        tk[tk_idx++] = uint8_t(lit_len << 5);
    }
    memcpy(dst + dst_idx, src + prev_anchor, size_t(lit_len));
    dst_idx += size_t(lit_len);
    const uint32_t a = uint32_t(dst_idx), b = uint32_t(tk_idx), c = uint32_t(m_idx);
    for (int i = 0; i < 4; i++) {
        dst[i] = uint8_t(a >> (8 * i));
        dst[4 + i] = uint8_t(b >> (8 * i));
        dst[8 + i] = uint8_t(c >> (8 * i));
    }
    dst[12] = flags;
    memcpy(dst + dst_idx, tk.data(), tk_idx);
    dst_idx += tk_idx;
    memcpy(dst + dst_idx, mb.data(), m_idx);
    dst_idx += m_idx;
    memcpy(dst + dst_idx, mlenb.data(), mlen_idx);
    dst_idx += mlen_idx;
    if (dst_idx > size_t(count - count / 100)) return -1;
    return int64_t(dst_idx);
}
}
