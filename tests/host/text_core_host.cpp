// Host harness for kanzi-go_b200/csrc/kz_text_core.cuh (tests/test_text_core_host.py): runs the product's TEXT state machine on the CPU
// exactly as the kernels of kz_text.cu drive it (statistics -> mode -> dictionary set-up -> walk), so that the oracle can check it where
// no GPU is available. This is a test of the product's host/device-shared logic, not a CPU path of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kanzi-go_b200/csrc/kz_text_core.cuh"
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
    Work(uint64_t stream_block_size, int count_for_size) {
        const uint32_t lg = log_hash_size(stream_block_size);
        map.assign(size_t(1) << lg, -1);
        list.resize(MAX_DICT_SIZE);
        const Static& S = stat();
        for (int i = 0; i < MAX_DICT_SIZE; i++) {
            if (i < S.n) list[i] = S.entries[i];
            else list[i] = Entry{0, i, NIL};
        }
        D.map = map.data();
        D.hash_mask = (1u << lg) - 1;
        D.list = list.data();
        D.dict_size = initial_dict_size(count_for_size);
        D.static_size = S.n;
        D.swords = S.words.data();
        for (int i = 0; i < S.n; i++) D.map[uint32_t(list[i].hash) & D.hash_mask] = i;
    }
};
}  // namespace

extern "C" {
// returns the encoded length, -1 = skip; *dt_out = ctx["dataType"] after the call. has_magic is decided by the caller (the kernels get it
// from the block's first four bytes).
int64_t tc_forward(const uint8_t* src, int count, uint8_t* dst, uint64_t stream_block_size, int dt_in, int has_magic, int* dt_out) {
    *dt_out = dt_in;
    if (count < 1024) return -1;
    if (dt_in != 0 && dt_in != 1 && dt_in != 7) return -1;  // DT_UNDEFINED / DT_TEXT / DT_BIN (:1232-1241)
    std::vector<uint32_t> f0(256, 0), f1(65536, 0);
    uint8_t prv = 0;
    for (int i = 0; i < count; i++) {
        f0[src[i]]++;
        f1[(uint32_t(prv) << 8) | src[i]]++;
        prv = src[i];
    }
    const uint8_t mode = text_stats_mode(f0.data(), f1.data(), count, has_magic != 0);
    if (mode & MASK_NOT_TEXT) {
        *dt_out = mode & MASK_DT;
        return -1;
    }
    *dt_out = 1;  // DT_TEXT
    Work w(stream_block_size, count);
    return forward_walk(src, count, dst, mode, w.D);
}
int64_t tc_inverse(const uint8_t* src, int len, uint8_t* dst, int64_t cap, uint64_t stream_block_size) {
    if (len < 2) return -1;
    Work w(stream_block_size, int(cap));
    return inverse_walk(src, len, dst, cap, w.D);
}
}
