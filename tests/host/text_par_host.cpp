// Host harness for kanzi-go_b200/csrc/kz_text_par_core.cuh (tests/test_text_par_host.py): the phases of the parallel TEXT forward (what the
// kernels of kz_text_par.cu run, here one array pass after the other) against the one-pass forward_walk of kz_text_core.cuh.
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kanzi-go_b200/csrc/kz_text_par_core.cuh"
#include "../../kanzi-go_b200/csrc/_gen/kz_text_dict.inc"

using namespace kz::textc;

namespace {
struct Static {
    std::vector<uint8_t> words;
    std::vector<Entry> entries;
    int n = 0;
    Static() {
        const int len = int(sizeof(KZ_TC_DICT_EN_1024) - 1);
        words.assign(KZ_TC_DICT_EN_1024, KZ_TC_DICT_EN_1024 + len);
        entries.resize(STATIC_WORDS);
        n = create_static_dictionary(words.data(), len, entries.data());
    }
};
const Static& stat() {
    static const Static s;
    return s;
}
struct Work {
    std::vector<int32_t> map;
    std::vector<Entry> list;
    Dict D;
    Work(uint64_t bs, int count) {
        const uint32_t lg = log_hash_size(bs);
        map.assign(size_t(1) << lg, -1);
        list.resize(MAX_DICT_SIZE);
        const Static& S = stat();
        for (int i = 0; i < MAX_DICT_SIZE; i++) list[i] = i < S.n ? S.entries[i] : Entry{0, i, NIL};
        D.map = map.data();
        D.hash_mask = (1u << lg) - 1;
        D.list = list.data();
        D.dict_size = initial_dict_size(count);
        D.static_size = S.n;
        D.swords = S.words.data();
        for (int i = 0; i < S.n; i++) D.map[uint32_t(list[i].hash) & D.hash_mask] = i;
    }
};
}  // namespace

extern "C" {
// both return the encoded length (-1 = skip); *ncand / *nfound report the sizes of the intermediate lists of the three-phase version
int64_t tp_walk(const uint8_t* src, int count, uint8_t* dst, uint64_t bs, int mode) {
    Work w(bs, count);
    return forward_walk(src, count, dst, uint8_t(mode), w.D);
}
int64_t tp_phases(const uint8_t* src, int count, uint8_t* dst, uint64_t bs, int mode, int* ncand, int* nfound) {
    std::vector<Candidate> cand;
    int prev = -1;
    for (int p = 0; p < count; p++) {  // phase A (a max-scan + one thread per position on the GPU)
        Candidate c;
        if (candidate_at(src, count, p, prev, &c)) cand.push_back(c);
        if (!is_text(src[p])) prev = p;
    }
    Work w(bs, count);
    std::vector<Found> found(cand.size() + 1);
    const int nf = dictionary_pass(src, cand.data(), int(cand.size()), w.D, found.data());  // phase B
    *ncand = int(cand.size());
    *nfound = nf;
    return emit_pass(src, count, uint8_t(mode), found.data(), nf, dst);  // phase C
}

// The fully parallel statement: candidates, dictionary rounds to the fixed point, emission through prefix sums (array passes only).
// *rounds = dictionary rounds needed; returns the encoded length, -1 = skip, -2 = the parallel statement does not apply (list wrap / slot 0 overflow)
int64_t tp_parallel(const uint8_t* src, int count, uint8_t* dst, uint64_t bs, int mode, int* rounds, int* ncand_out, int* nfound_out) {
    std::vector<Candidate> cand;
    int prev = -1;
    for (int p = 0; p < count; p++) {
        Candidate c;
        if (candidate_at(src, count, p, prev, &c)) cand.push_back(c);
        if (!is_text(src[p])) prev = p;
    }
    const int nc = int(cand.size());
    *ncand_out = nc;
    const Static& S = stat();
    const uint32_t lg = log_hash_size(bs);
    const uint32_t mask = (1u << lg) - 1;
    std::vector<int32_t> owner0(size_t(1) << lg, OCC_EMPTY), owner;
    for (int i = 0; i < S.n; i++) owner0[uint32_t(S.entries[i].hash) & mask] = -(i + 1);  // later static words win a shared slot
    std::vector<uint8_t> ins(nc), ins_new(nc);
    for (int k = 0; k < nc; k++) ins[k] = cand[k].end - cand[k].start >= 3;
    std::vector<uint32_t> rank(nc + 1);
    std::vector<int32_t> focc(nc), fflip(nc);
    SlotZero z0;
    DictView V;
    V.src = src;
    V.cand = cand.data();
    V.mask = mask;
    V.sdict = S.entries.data();
    V.swords = S.words.data();
    V.z0 = &z0;
    int r = 0;
    for (;; r++) {
        if (r > 64) return -2;
        owner = owner0;
        z0.n = 0;
        z0.first_ins = INT32_MAX;
        uint32_t acc = 0;
        for (int k = 0; k < nc; k++) {  // kernel: owners (atomicMin), slot 0 list, first insert; scan: rank
            rank[k] = acc;
            if (!ins[k]) continue;
            acc++;
            if (z0.first_ins == INT32_MAX) z0.first_ins = k;
            const uint32_t s1 = cand[k].h1 & mask;
            if (s1 == 0) {
                if (z0.n >= Z0_MAX) return -2;
                z0.idx[z0.n] = k;
                z0.nx[z0.n] = INT32_MAX;
                z0.n++;
            } else if (owner[s1] >= 0 && k < owner[s1]) {
                owner[s1] = k;
            }
        }
        rank[nc] = acc;
        for (int k = 0; k < nc; k++)  // kernel: next insert behind every slot 0 word
            if (ins[k])
                for (int z = 0; z < z0.n; z++)
                    if (k > z0.idx[z] && k < z0.nx[z]) z0.nx[z] = k;
        V.owner = owner.data();
        bool changed = false;
        for (int k = 0; k < nc; k++) {  // kernel: one thread per candidate
            ins_new[k] = eval_candidate(V, k, rank[k], &focc[k], &fflip[k]);
            changed = changed || ins_new[k] != ins[k];
        }
        if (!changed) {
            if (STATIC_WORDS + (int64_t)acc >= MAX_DICT_SIZE) return -2;  // the entry list would wrap around: the serial walk handles that
            break;
        }
        ins.swap(ins_new);
    }
    *rounds = r + 1;
    // found words in order
    std::vector<int32_t> fstart, flen, findex, fflp;
    for (int k = 0; k < nc; k++)
        if (focc[k] != OCC_EMPTY) {
            fstart.push_back(cand[k].start);
            flen.push_back(cand[k].end - cand[k].start);
            findex.push_back(occupant_index(focc[k], rank.data()));
            fflp.push_back(fflip[k]);
        }
    const int nf = int(fstart.size());
    *nfound_out = nf;
    // emission through prefix sums
    const bool is_crlf = (mode & MASK_CRLF) != 0;
    std::vector<uint32_t> cp(size_t(count) + 1);
    cp[0] = 0;
    for (int i = 0; i < count; i++) cp[i + 1] = cp[i] + uint32_t(symbol_cost(src[i], is_crlf));
    int lead = 0;
    while (lead < count && src[lead] == ' ') lead++;
    if (lead >= count) return -1;
    const int64_t dst_end = count, dst_end3 = dst_end - 3;
    std::vector<int32_t> from(nf + 1), to(nf + 1);
    std::vector<int64_t> outpos(nf + 1);
    int64_t pos = 1 + lead;
    bool fail = false;
    for (int f = 0; f < nf; f++) {  // kernel: per found word cost; scan: outpos
        const int emit_anchor = f == 0 ? lead : fstart[f - 1] + flen[f - 1];
        from[f] = emit_anchor;
        to[f] = fstart[f];
        if (emit_anchor == fstart[f] - 1 && src[emit_anchor] == ' ') to[f] = emit_anchor;
        outpos[f] = pos;
        const int64_t span = int64_t(cp[to[f]]) - int64_t(cp[from[f]]);
        if (pos + span >= dst_end3) fail = true;
        pos += span + fflp[f] + index_cost(findex[f]);
    }
    const int tail_from = nf == 0 ? lead : fstart[nf - 1] + flen[nf - 1];
    from[nf] = tail_from;
    to[nf] = count;
    outpos[nf] = pos;
    const int64_t total = pos + (int64_t(cp[count]) - int64_t(cp[tail_from]));
    if (fail || total > dst_end) return -1;
    dst[0] = uint8_t(mode);
    for (int i = 0; i < lead; i++) dst[1 + i] = ' ';
    for (int f = 0; f < nf; f++) {  // kernel: per found word: flip + index bytes
        int64_t at = outpos[f] + (int64_t(cp[to[f]]) - int64_t(cp[from[f]]));
        if (fflp[f]) dst[at++] = MASK_FLIP_CASE;
        emit_word_index2(dst + at, findex[f]);
    }
    for (int i = lead; i < count; i++) {  // kernel: per byte (the found word in front of i by binary search)
        int lo = 0, hi = nf;  // F = number of found words with start <= i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (fstart[mid] <= i) lo = mid + 1;
            else hi = mid;
        }
        const int F = lo;
        if (F >= 1 && i < fstart[F - 1] + flen[F - 1]) continue;  // inside a dictionary word
        if (i >= to[F]) continue;                                    // the implied single space
        uint8_t* d = dst + outpos[F] + (int64_t(cp[i]) - int64_t(cp[from[F]]));
        const uint8_t cur = src[i];
        if (cur == ESCAPE_TOKEN1) {
            d[0] = ESCAPE_TOKEN1;
            d[1] = ESCAPE_TOKEN1;
        } else if (cur == CR) {
            if (!is_crlf) d[0] = cur;
        } else if (cur >= 0x80) {
            d[0] = ESCAPE_TOKEN1;
            d[1] = cur;
        } else {
            d[0] = cur;
        }
    }
    return total;
}

int64_t tp_inverse_walk(const uint8_t* src, int len, uint8_t* dst, int64_t cap, uint64_t bs) {
    Work w(bs, int(std::min<int64_t>(cap, 0x7FFFFFFF)));
    return inverse_walk(src, len, dst, cap, w.D);
}
// The data-parallel statement of the inverse, array pass by array pass. -2: an anomaly the parallel statement leaves to the serial walk.
int64_t tp_inverse_parallel(const uint8_t* src_in, int len, uint8_t* dst, int64_t cap, uint64_t bs, int* rounds) {
    if (len < 2) return -2;
    std::vector<uint8_t> padded(size_t(len) + 8, 0);
    memcpy(padded.data(), src_in, size_t(len));
    const uint8_t* src = padded.data();
    const Static& S = stat();
    const bool is_crlf = (src[0] & MASK_CRLF) != 0;
    // token starts: scan of the automaton from state 0 at position 1
    std::vector<uint8_t> start(size_t(len) + 1, 0);
    {
        uint32_t st = 0;
        for (int i = 1; i < len; i++) {
            start[i] = st == 0;
            st = token_step(src, i, st);
        }
        if (st != 0) return -2;  // the last token runs past the end of the stream
    }
    // previous non-letter token of every non-letter token (a max-scan), delim_anchor, candidates
    std::vector<Candidate> cand;
    std::vector<int32_t> prevq(size_t(len) + 1, -1);  // per non-letter token start: start of the previous non-letter token (-1: none)
    {
        int q = -1;
        for (int p = 1; p < len; p++) {
            if (!start[p] || token_kind(src[p]) == TK_LETTER) continue;
            prevq[p] = q;
            q = p;
        }
    }
    auto tok_wordlen = [&](int q) -> int {  // length of the dictionary word of an index token, as far as delim_anchor / word_run need it: > 1 or not
        int flip;
        const int idx = token_index(src, q, &flip);
        if (idx >= 0 && idx < S.n) return S.entries[idx].data >> 24 & 0xFF;
        return 3;  // dynamic words are at least 3 letters long
    };
    auto delim_anchor_at = [&](int p) -> int {
        const int q = prevq[p];
        if (q < 0) return is_text(src[1]) ? 0 : 1;
        const int qend = q + token_len(src, q);
        if (token_kind(src[q]) == TK_IDX && tok_wordlen(q) > 1) return qend;
        return qend - 1;
    };
    for (int p = 1; p < len; p++) {
        if (!start[p] || token_kind(src[p]) == TK_LETTER) continue;
        const int a = delim_anchor_at(p);
        if (p > a + 3 && is_delimiter(src[p])) {
            const int length = p - a - 1;
            if (length <= MAX_WORD_LENGTH) {
                Candidate c;
                c.end = p;
                c.start = a + 1;
                uint32_t h = HASH1;
                for (int i = a + 1; i < p; i++) h = hash_step(h, src[i]);
                c.h1 = h;
                c.h2 = 0;
                cand.push_back(c);
            }
        }
    }
    const int nc = int(cand.size());
    const uint32_t lg = log_hash_size(bs);
    const uint32_t mask = (1u << lg) - 1;
    std::vector<int32_t> owner0(size_t(1) << lg, OCC_EMPTY), owner;
    for (int i = 0; i < S.n; i++) owner0[uint32_t(S.entries[i].hash) & mask] = -(i + 1);
    std::vector<uint8_t> ins(nc, 1), ins_new(nc);
    std::vector<uint32_t> rank(nc + 1);
    std::vector<int32_t> focc(nc), fflip(nc);
    SlotZero z0;
    DictView V;
    V.src = src;
    V.cand = cand.data();
    V.mask = mask;
    V.sdict = S.entries.data();
    V.swords = S.words.data();
    V.z0 = &z0;
    int r = 0;
    uint32_t total_ins = 0;
    for (;; r++) {
        if (r > 64) return -2;
        owner = owner0;
        z0.n = 0;
        z0.first_ins = INT32_MAX;
        uint32_t acc = 0;
        bool z0_over = false;
        for (int k = 0; k < nc; k++) {
            rank[k] = acc;
            if (!ins[k]) continue;
            acc++;
            if (z0.first_ins == INT32_MAX) z0.first_ins = k;
            const uint32_t s1 = cand[k].h1 & mask;
            if (s1 == 0) {
                if (z0.n >= Z0_MAX) z0_over = true;
                else {
                    z0.idx[z0.n] = k;
                    z0.nx[z0.n] = INT32_MAX;
                    z0.n++;
                }
            } else if (owner[s1] >= 0 && k < owner[s1]) {
                owner[s1] = k;
            }
        }
        rank[nc] = acc;
        for (int k = 0; k < nc; k++)
            if (ins[k])
                for (int z = 0; z < z0.n; z++)
                    if (k > z0.idx[z] && k < z0.nx[z]) z0.nx[z] = k;
        V.owner = owner.data();
        bool changed = false;
        for (int k = 0; k < nc; k++) {
            ins_new[k] = eval_candidate(V, k, rank[k], &focc[k], &fflip[k], true);
            changed = changed || ins_new[k] != ins[k];
        }
        if (!changed) {
            if (z0_over || STATIC_WORDS + (int64_t)acc >= MAX_DICT_SIZE) return -2;
            total_ins = acc;
            break;
        }
        ins.swap(ins_new);
    }
    *rounds = r + 1;
    std::vector<int32_t> ins_list;  // k-th inserted word -> candidate
    for (int k = 0; k < nc; k++)
        if (ins[k]) ins_list.push_back(k);
    (void)total_ins;
    // per token: output length; anomalies
    std::vector<uint32_t> outpos(size_t(len) + 1, 0);
    uint64_t pos = 0;
    for (int p = 1; p < len; p++) {
        if (!start[p]) continue;
        outpos[p] = uint32_t(pos);
        const uint8_t cur = src[p];
        const int kind = token_kind(cur);
        if (kind == TK_LETTER || kind == TK_ESC) pos += 1;
        else if (kind == TK_LIT) pos += (is_crlf && cur == LF) ? 2 : 1;
        else {
            int flip;
            const int idx = token_index(src, p, &flip);
            if (idx < 0) return -2;
            int wl;
            if (idx < STATIC_WORDS) {
                if (idx >= S.n) return -2;
                wl = (S.entries[idx].data >> 24) & 0xFF;
            } else {
                // words inserted before this token: candidates that end before p
                int lo = 0, hi = nc;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cand[mid].end < p) lo = mid + 1;
                    else hi = mid;
                }
                const uint32_t before = rank[lo];
                if (uint32_t(idx - STATIC_WORDS) >= before) return -2;
                const Candidate& c = cand[ins_list[idx - STATIC_WORDS]];
                wl = c.end - c.start;
            }
            const int q = prevq[p];
            const bool word_run = q >= 0 && token_kind(src[q]) == TK_IDX && tok_wordlen(q) > 1;
            pos += uint64_t(wl) + ((word_run && wl > 1) ? 1 : 0);
        }
    }
    if (pos + 64 >= uint64_t(cap)) return -2;
    // emit
    for (int p = 1; p < len; p++) {
        if (!start[p]) continue;
        uint8_t* d = dst + outpos[p];
        const uint8_t cur = src[p];
        const int kind = token_kind(cur);
        if (kind == TK_LETTER) d[0] = cur;
        else if (kind == TK_ESC) d[0] = src[p + 1];
        else if (kind == TK_LIT) {
            if (is_crlf && cur == LF) {
                d[0] = CR;
                d[1] = LF;
            } else d[0] = cur;
        } else {
            int flip;
            const int idx = token_index(src, p, &flip);
            const uint8_t* w;
            int wl;
            if (idx < STATIC_WORDS) {
                w = S.words.data() + (~S.entries[idx].ptr);
                wl = (S.entries[idx].data >> 24) & 0xFF;
            } else {
                const Candidate& c = cand[ins_list[idx - STATIC_WORDS]];
                w = src + c.start;
                wl = c.end - c.start;
            }
            const int q = prevq[p];
            const bool word_run = q >= 0 && token_kind(src[q]) == TK_IDX && tok_wordlen(q) > 1;
            if (word_run && wl > 1) *d++ = ' ';
            for (int i = 0; i < wl; i++) d[i] = w[i];
            if (flip) d[0] ^= 0x20;
        }
    }
    return int64_t(pos);
}
}
