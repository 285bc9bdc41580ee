// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Transform dispatch (v2/transform/Factory.go:97-185).
#include <algorithm>
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {

size_t transform_max_encoded_len(uint64_t type, size_t n) {
    switch (type) {
        case T_NONE: return n;                                   // NullTransform.go MaxEncodedLen
        case T_BWT: return n + 33;                               // BWTBlockCodec.go:228-230 (_BWT_MAX_HEADER_SIZE = 8*4+1)
        case T_LZ: case T_LZX: return lz_max_encoded_len(n);     // LZCodec.go:935-941
        case T_BWTS: return n;                                   // BWTS.go:291-293
        case T_ROLZ: return rolz_max_encoded_len(n);             // ROLZCodec.go:916-918
        case T_RANK: case T_MTFT: return n + 33;                 // SBRT.go:229-231
        case T_ZRLT: return n;                                   // ZRLT.go:228-230
        case T_PACK: case T_DNA: return alias_max_encoded_len(n);
        case T_MM: return fsd_max_encoded_len(n);
        case T_UTF: return utf_max_encoded_len(n);
        case T_DICT: return text_max_encoded_len(n);
        case T_EXE: return exe_max_encoded_len(n);
        default: throw Error(ERR_CREATE_CODEC, "transform not restated in the oracle");
    }
}

bool transform_forward(uint64_t type, Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    switch (type) {
        case T_NONE:  // NullTransform.Forward: plain copy, never fails for valid buffers
            if (cap < n) return false;
            std::memcpy(dst, src, n);
            *out_n = n;
            return true;
        case T_BWT: return bwt_block_forward(ctx, src, n, dst, cap, out_n);
        case T_LZ: return lz_forward(ctx, false, src, n, dst, cap, out_n);
        case T_LZX: return lz_forward(ctx, true, src, n, dst, cap, out_n);
        case T_BWTS: return bwts_forward(src, n, dst, cap, out_n);
        case T_ROLZ: return rolz_forward(ctx, src, n, dst, cap, out_n);
        case T_RANK: return sbrt_forward(2, src, n, dst, cap, out_n);  // Factory.go: RANK -> SBRT_MODE_RANK
        case T_MTFT: return sbrt_forward(1, src, n, dst, cap, out_n);  // MTFT -> SBRT_MODE_MTF
        case T_ZRLT: return zrlt_forward(src, n, dst, cap, out_n);
        case T_PACK: return alias_forward(ctx, false, src, n, dst, cap, out_n);
        case T_DNA: return alias_forward(ctx, true, src, n, dst, cap, out_n);  // Factory.go:151-153 packOnlyDNA
        case T_MM: return fsd_forward(ctx, src, n, dst, cap, out_n);
        case T_UTF: return utf_forward(ctx, src, n, dst, cap, out_n);
        case T_DICT: return text_forward(ctx, src, n, dst, cap, out_n);
        case T_EXE: return exe_forward(ctx, src, n, dst, cap, out_n);
        default: throw Error(ERR_CREATE_CODEC, "transform not restated in the oracle");
    }
}

bool transform_inverse(uint64_t type, Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    switch (type) {
        case T_NONE:
            if (cap < n) return false;
            std::memcpy(dst, src, n);
            *out_n = n;
            return true;
        case T_BWT: return bwt_block_inverse(ctx, src, n, dst, cap, out_n);
        case T_LZ: return lz_inverse(ctx, false, src, n, dst, cap, out_n);
        case T_LZX: return lz_inverse(ctx, true, src, n, dst, cap, out_n);
        case T_BWTS: return bwts_inverse(src, n, dst, cap, out_n);
        case T_ROLZ: return rolz_inverse(ctx, src, n, dst, cap, out_n);
        case T_RANK: return sbrt_inverse(2, src, n, dst, cap, out_n);
        case T_MTFT: return sbrt_inverse(1, src, n, dst, cap, out_n);
        case T_ZRLT: return zrlt_inverse(src, n, dst, cap, out_n);
        case T_PACK: case T_DNA: return alias_inverse(src, n, dst, cap, out_n);
        case T_MM: return fsd_inverse(src, n, dst, cap, out_n);
        case T_UTF: return utf_inverse(src, n, dst, cap, out_n);
        case T_DICT: return text_inverse(ctx, src, n, dst, cap, out_n);
        case T_EXE: return exe_inverse(src, n, dst, cap, out_n);
        default: throw Error(ERR_INVALID_CODEC, "transform not restated in the oracle");
    }
}

}  // namespace kzo
