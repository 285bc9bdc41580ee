// Host harness for kanzi-go_b200/csrc/next/kz_text_par_core.cuh (tests/test_text_par_host.py): the three-phase TEXT forward planned for
// the next round against the one-pass forward_walk of kz_text_core.cuh.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kanzi-go_b200/csrc/next/kz_text_par_core.cuh"
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
}
