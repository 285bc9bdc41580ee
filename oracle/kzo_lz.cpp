// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). LZ / LZX codec (v2/transform/LZCodec.go) — placeholder until restated.
#include "kzo.hpp"
#include "kzo_transforms.hpp"

namespace kzo {
size_t lz_max_encoded_len(size_t n) { return n <= 1024 ? n + 16 : n + n / 64; }
bool lz_forward(Ctx&, bool, const uint8_t*, size_t, uint8_t*, size_t, size_t*) { throw Error(ERR_CREATE_CODEC, "LZ not restated yet"); }
bool lz_inverse(Ctx&, bool, const uint8_t*, size_t, uint8_t*, size_t, size_t*) { throw Error(ERR_CREATE_CODEC, "LZ not restated yet"); }
}  // namespace kzo
