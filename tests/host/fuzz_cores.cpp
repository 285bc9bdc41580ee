// ASan fuzz of the shared host/device cores: corrupted inputs to the inverse walks must fail cleanly, never touch memory out of bounds
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../kanzi-go_b200/csrc/kz_text_core.cuh"
#include "../../kanzi-go_b200/csrc/_gen/kz_text_dict.inc"
#include "../../kanzi-go_b200/csrc/kz_utf_core.cuh"
#include "../../kanzi-go_b200/csrc/kz_exe_core.cuh"
using namespace kz;
struct TextWork {
    std::vector<uint8_t> words; std::vector<textc::Entry> stat; int n;
    std::vector<int32_t> map; std::vector<textc::Entry> list; textc::Dict D;
    TextWork() {
        int len = sizeof(KZ_TC_DICT_EN_1024) - 1;
        words.assign(KZ_TC_DICT_EN_1024, KZ_TC_DICT_EN_1024 + len);
        stat.resize(1024); n = textc::create_static_dictionary(words.data(), len, stat.data());
        list.resize(textc::MAX_DICT_SIZE);
    }
    void reset(uint64_t bs, int count) {
        uint32_t lg = textc::log_hash_size(bs);
        map.assign(size_t(1) << lg, -1);
        for (int i = 0; i < textc::MAX_DICT_SIZE; i++) list[i] = i < n ? stat[i] : textc::Entry{0, i, textc::NIL};
        D.map = map.data(); D.hash_mask = (1u << lg) - 1; D.list = list.data(); D.dict_size = textc::initial_dict_size(count);
        D.static_size = n; D.swords = words.data();
        for (int i = 0; i < n; i++) D.map[uint32_t(list[i].hash) & D.hash_mask] = i;
    }
};
int main(int argc, char** argv) {
    const int ITERS = argc > 1 ? atoi(argv[1]) : 500;
    std::mt19937 rng(12345);
    TextWork tw;
    const char* sample = "the quick brown fox jumps over the lazy dog, and that is what people have with because The And said. ";
    long fails = 0, oks = 0;
    for (int iter = 0; iter < ITERS; iter++) {
        int n = 1024 + rng() % 6000;
        std::vector<uint8_t> x(n);
        for (int i = 0; i < n; i++) x[i] = (rng() % 10 < 9) ? sample[(i * 7 + iter) % strlen(sample)] : uint8_t(rng());
        // TEXT forward then corrupt then inverse
        std::vector<uint8_t> enc(n + 64), dec(n + 4096 + 16);
        tw.reset(65536, n);
        int m = textc::forward_walk(x.data(), n, enc.data(), 0, tw.D);
        if (m > 0) {
            std::vector<uint8_t> e(enc.begin(), enc.begin() + m);  // exact-size buffer: ASan sees any overread
            for (int k = 0; k < 1 + int(rng() % 4); k++) e[rng() % m] = uint8_t(rng());
            if (rng() % 3 == 0) e.resize(std::max<size_t>(2, rng() % m));
            std::vector<uint8_t> d(n + (rng() % 3 ? 4096 : 0) - (rng() % 5 == 0 ? n / 2 : 0));
            tw.reset(65536, int(d.size()));
            int64_t r = textc::inverse_walk(e.data(), int(e.size()), d.data(), int64_t(d.size()), tw.D);
            (r < 0 ? fails : oks)++;
        }
        // UTF inverse on garbage and on corrupted valid streams
        {
            std::vector<uint8_t> g(16 + rng() % 3000);
            for (auto& b : g) b = uint8_t(rng());
            if (rng() % 2) { g[2] = 0; g[3] = uint8_t(1 + rng() % 40); }
            std::vector<uint8_t> d(64 + rng() % 6000);
            std::vector<utfc::Sym> sym(utfc::MAX_SYMBOLS);
            int64_t r = utfc::inverse_walk(g.data(), int(g.size()), d.data(), int64_t(d.size()), sym.data());
            (r < 0 ? fails : oks)++;
        }
        // EXE inverse on garbage with plausible headers
        {
            std::vector<uint8_t> g(16 + rng() % 3000);
            for (auto& b : g) b = uint8_t(rng());
            g[0] = (rng() % 2) ? 0x40 : 0x20;
            uint32_t cs = rng() % 64, ce = 9 + rng() % g.size();
            memcpy(&g[1], &cs, 4); memcpy(&g[5], &ce, 4);
            std::vector<uint8_t> d(rng() % 4000 + 1);
            int64_t r = exec::inverse(g.data(), int64_t(g.size()), d.data(), int64_t(d.size()));
            (r < 0 ? fails : oks)++;
        }
        // EXE forward on random data with executable magics (header parsers)
        {
            int n2 = 4096 + rng() % 5000;
            std::vector<uint8_t> g(n2);
            for (auto& b : g) b = uint8_t(rng());
            static const uint8_t mg[5][4] = {{0x7F, 'E', 'L', 'F'}, {'M', 'Z', 0x90, 0}, {0xCF, 0xFA, 0xED, 0xFE}, {0xFE, 0xED, 0xFA, 0xCE}, {0xCE, 0xFA, 0xED, 0xFE}};
            memcpy(g.data(), mg[rng() % 5], 4);
            if (rng() % 2) { g[4] = 1 + rng() % 2; g[5] = 1 + rng() % 2; }
            if (rng() % 2) { uint32_t two = 2; memcpy(&g[12], &two, 4); uint32_t nc = rng() % 40; memcpy(&g[16], &nc, 4); }
            std::vector<uint8_t> d(n2 + n2 / 8 + 64);
            int dt = 0;
            exec::forward(g.data(), n2, d.data(), int64_t(d.size()), &dt);
        }
    }
    printf("fuzz done: %ld clean failures, %ld accepted\n", fails, oks);
    return 0;
}
