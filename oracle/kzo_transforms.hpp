// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Internal declarations of the restated transforms.
#pragma once
#include "kzo.hpp"

namespace kzo {
bool bwt_block_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool bwt_block_inverse(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t lz_max_encoded_len(size_t n);
bool lz_forward(Ctx& ctx, bool extra, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool lz_inverse(Ctx& ctx, bool extra, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool sbrt_forward(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool sbrt_inverse(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool zrlt_forward(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool zrlt_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool bwts_forward(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool bwts_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t rolz_max_encoded_len(size_t n);
bool rolz_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool rolz_inverse(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
int detect_simple_type(size_t count, const int* freqs0);
uint32_t get_magic_type(const uint8_t* src, size_t n);  // internal/Magic.go:73-112
// pre-transforms (kzo_pre.cpp, kzo_text.cpp)
size_t alias_max_encoded_len(size_t n);
bool alias_forward(Ctx& ctx, bool only_dna, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool alias_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t fsd_max_encoded_len(size_t n);
bool fsd_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool fsd_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t utf_max_encoded_len(size_t n);
bool utf_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool utf_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t exe_max_encoded_len(size_t n);
bool exe_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool exe_inverse(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool text_available();  // false when the static dictionary could not be extracted (no reference tree at build time)
size_t text_max_encoded_len(size_t n);
bool text_forward(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool text_inverse(Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
}  // namespace kzo
