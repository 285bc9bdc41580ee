// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline).
#include <chrono>
#include <cstring>

#include "kzo.hpp"
#include "kzo_transforms.hpp"

using namespace kzo;

#define KZO_GUARD(body)                       \
    try {                                     \
        body;                                 \
        return 0;                             \
    } catch (const Error& e) {                \
        set_err(e.what());                    \
        return -e.code;                       \
    } catch (const std::exception& e) {       \
        set_err(e.what());                    \
        return -ERR_UNKNOWN;                  \
    }

static thread_local char g_err[256];
static void set_err(const char* m) {
    std::strncpy(g_err, m, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}

extern "C" {

const char* kzo_last_error() { return g_err; }

// Encodes `n` bytes with EntropyEncoder.Write semantics into a fresh bit stream. dst gets the zero padded bytes.
int kzo_entropy_encode(uint32_t type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, uint64_t* out_bits) {
    KZO_GUARD({
        BitWriter bw;
        bw.reserve(n + n / 8 + 1024);
        entropy_encode(type, bw, src, n);
        *out_bits = bw.written();
        auto& b = bw.close();
        if (b.size() > cap) throw Error(ERR_WRITE_FILE, "output buffer too small");
        std::memcpy(dst, b.data(), b.size());
    })
}

int kzo_entropy_decode(uint32_t type, const uint8_t* src, size_t src_bytes, uint8_t* dst, size_t n, uint64_t* consumed_bits) {
    KZO_GUARD({
        BitReader br = BitReader::over_bytes(src, src_bytes);
        entropy_decode(type, br, dst, n);
        if (consumed_bits) *consumed_bits = br.read();
    })
}

// ctx["blockSize"] for the single-transform calls below (0 = the call's own input length); thread local
static thread_local size_t g_stream_block_size = 0;
void kzo_set_stream_block_size(size_t n) { g_stream_block_size = n; }
int kzo_text_available() { return text_available() ? 1 : 0; }

int kzo_transform_forward(uint64_t type, int data_type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n, int* out_data_type) {
    try {
        Ctx ctx;
        ctx.data_type = data_type;
        ctx.block_size = n;
        ctx.stream_block_size = g_stream_block_size;
        bool ok = transform_forward(type, ctx, src, n, dst, cap, out_n);
        if (out_data_type) *out_data_type = ctx.data_type;
        return ok ? 0 : 1;  // 1 = "skip me" (the reference returns a non-nil error)
    } catch (const Error& e) {
        set_err(e.what());
        return -e.code;
    } catch (const std::exception& e) {
        set_err(e.what());
        return -ERR_UNKNOWN;
    }
}

int kzo_transform_inverse(uint64_t type, int data_type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    try {
        Ctx ctx;
        ctx.data_type = data_type;
        ctx.block_size = n;
        ctx.stream_block_size = g_stream_block_size;
        bool ok = transform_inverse(type, ctx, src, n, dst, cap, out_n);
        return ok ? 0 : -ERR_PROCESS_BLOCK;
    } catch (const Error& e) {
        set_err(e.what());
        return -e.code;
    } catch (const std::exception& e) {
        set_err(e.what());
        return -ERR_UNKNOWN;
    }
}

size_t kzo_transform_max_encoded_len(uint64_t type, size_t n) {
    try {
        return transform_max_encoded_len(type, n);
    } catch (...) {
        return 0;
    }
}

void kzo_bwt_forward_raw(const uint8_t* src, uint8_t* dst, size_t n, uint32_t* primary8) { bwt_forward_raw(src, dst, n, primary8); }
int kzo_bwt_inverse_raw(const uint8_t* src, uint8_t* dst, size_t n, const uint32_t* primary8) { return bwt_inverse_raw(src, dst, n, primary8) ? 0 : -1; }

int kzo_encode_block(const uint8_t* data, size_t n, uint64_t transform48, uint32_t entropy, int checksum_bits, int skip_blocks, uint8_t* dst,
                     size_t cap, uint64_t* out_bits) {
    KZO_GUARD({
        BlockResult r;
        encode_block(data, n, transform48, entropy, checksum_bits, skip_blocks != 0, r);
        if (r.bytes.size() > cap) throw Error(ERR_WRITE_FILE, "output buffer too small");
        std::memcpy(dst, r.bytes.data(), r.bytes.size());
        *out_bits = r.bits;
    })
}

int kzo_decode_block(const uint8_t* payload, uint64_t bits, uint64_t transform48, uint32_t entropy, int checksum_bits, size_t block_size,
                     uint8_t* dst, size_t cap, size_t* out_n) {
    KZO_GUARD({
        std::vector<uint8_t> out;
        size_t n = decode_block(payload, bits, transform48, entropy, checksum_bits, block_size, out);
        if (n > cap) throw Error(ERR_WRITE_FILE, "output buffer too small");
        std::memcpy(dst, out.data(), n);
        *out_n = n;
    })
}

// Whole stream. Returns the elapsed seconds of the compression itself in *seconds (for the CPU baseline).
int kzo_compress(const uint8_t* src, size_t n, uint64_t transform48, uint32_t entropy, uint32_t block_size, int checksum_bits, int jobs,
                 int64_t input_size, uint8_t* dst, size_t cap, size_t* out_n, double* seconds) {
    KZO_GUARD({
        StreamParams p;
        p.transform48 = transform48;
        p.entropy = entropy;
        p.block_size = block_size;
        p.checksum_bits = checksum_bits;
        p.jobs = jobs;
        p.input_size = input_size;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<uint8_t> out = compress_stream(src, n, p);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (out.size() > cap) throw Error(ERR_WRITE_FILE, "output buffer too small");
        std::memcpy(dst, out.data(), out.size());
        *out_n = out.size();
    })
}

int kzo_decompress(const uint8_t* src, size_t n, int jobs, uint8_t* dst, size_t cap, size_t* out_n, double* seconds) {
    KZO_GUARD({
        auto t0 = std::chrono::steady_clock::now();
        std::vector<uint8_t> out = decompress_stream(src, n, jobs);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        if (out.size() > cap) throw Error(ERR_WRITE_FILE, "output buffer too small");
        std::memcpy(dst, out.data(), out.size());
        *out_n = out.size();
    })
}

uint64_t kzo_parse_transform(const char* names) {
    try {
        return parse_transform_names(names);
    } catch (...) {
        return ~0ull;
    }
}
int kzo_parse_entropy(const char* name) {
    try {
        return int(parse_entropy_name(name));
    } catch (...) {
        return -1;
    }
}

uint32_t kzo_xxhash32(const uint8_t* p, size_t n, uint32_t seed) { return xxhash32(p, n, seed); }
uint64_t kzo_xxhash64(const uint8_t* p, size_t n, uint64_t seed) { return xxhash64(p, n, seed); }

// small hooks for the pin tests
int kzo_varint_len(uint32_t v) {
    BitWriter bw;
    return write_varint(bw, v);
}
int kzo_normalize(int* freqs256, int* alphabet256, int total, int scale) {
    try {
        return normalize_frequencies(freqs256, alphabet256, 256, total, scale);
    } catch (...) {
        return -1;
    }
}
int kzo_entropy1024(const uint8_t* p, size_t n) {
    int h[256];
    std::memset(h, 0, sizeof(h));
    histogram_order0(p, n, h);
    return first_order_entropy_1024(n, h);
}
// writes (len<<9)|bits of the signed Exp-Golomb code of every byte value into out[256] (out[0] = 1<<9 | 1)
}
