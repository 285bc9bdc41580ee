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
}
