// Host harness for kanzi-go_b200/csrc/kz_utf_core.cuh (tests/test_utf_core_host.py): the product's UTF logic driven on the CPU the way the
// kernels of kz_utf.cu drive it (histograms -> validation -> walk). A test of host/device-shared logic, not a CPU path of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kanzi-go_b200/csrc/kz_utf_core.cuh"

using namespace kz::utfc;

extern "C" {
int64_t uc_forward(const uint8_t* src, int count, uint8_t* dst, int dt_in, int* dt_out) {
    *dt_out = dt_in;
    if (count < MIN_BLOCKSIZE) return -1;
    if (dt_in != 0 && dt_in != 8) return -1;  // DT_UNDEFINED / DT_UTF8 (:105-114)
    const bool must_validate = dt_in != 8;
    const int start = utf_start(src);
    if (must_validate) {
        const uint8_t* blk = src + start;
        const uint32_t len = uint32_t(count - 4 - start);
        std::vector<uint32_t> f0(256, 0), f1(65536, 0);
        uint8_t prv = 0;
        uint32_t forb = 0;
        const uint32_t checked = validate_checked_len(len);
        for (uint32_t i = 0; i < len; i++) {
            f0[blk[i]]++;
            f1[(uint32_t(prv) << 8) | blk[i]]++;
            prv = blk[i];
            if (i < checked && forbidden_byte(blk[i])) forb++;
        }
        if (!validate_utf(f0.data(), f1.data(), len, forb)) return -1;
    }
    *dt_out = 8;
    std::vector<int32_t> alias(ALIAS_MAP_SIZE, 0);
    std::vector<uint64_t> keys(MAX_SYMBOLS);
    return forward_walk(src, count, start, dst, alias.data(), keys.data());
}
int64_t uc_inverse(const uint8_t* src, int count, uint8_t* dst, int64_t cap) {
    std::vector<Sym> m(MAX_SYMBOLS);
    return inverse_walk(src, count, dst, cap, m.data());
}
}
