// Shared device/host helpers for the kanzi B200 engine.
// Bit order everywhere is kanzi's: MSB first inside bytes, fields big-endian
// (reference: v2/bitstream/DefaultOutputBitStream.go:78-96, DefaultInputBitStream.go:78-94).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KZ_HD __host__ __device__ __forceinline__
#define KZ_D __device__ __forceinline__

// kanzi error codes (v2/Definitions.go:25-46), stored negated in status words
#define KZ_E_BLOCK_SIZE 2
#define KZ_E_INVALID_CODEC 3
#define KZ_E_PROCESS_BLOCK 13
#define KZ_E_CREATE_CODEC 14
#define KZ_E_CRC 19

namespace kz {

static const int ANS_TOP = 1 << 15;           // entropy/ANSRangeCodec.go:32
static const int ANS0_CHUNK = 16384;          // :33
static const int ANS_MAX_CHUNK = 1 << 27;     // :35
static const int HUF_CHUNK = 16384;           // entropy/HuffmanCodec.go:30
static const int HUF_MAX_BITS = 12;           // :31
static const int RANGE_CHUNK = 32768;         // entropy/RangeCodec.go:34

KZ_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

KZ_HD uint32_t log2_floor(uint32_t x) {  // internal/Global.go:156 Log2NoCheck
#ifdef __CUDA_ARCH__
    return 31u - (uint32_t)__clz((int)x);
#else
    return 31u - (uint32_t)__builtin_clz(x);
#endif
}

// Reads `n` (1..32) bits at absolute bit position `pos` of a big-endian bit string stored in 32-bit words
// (the buffer must be readable one word past the last touched bit).
KZ_HD uint32_t bits_at(const uint32_t* __restrict__ words, uint64_t pos, uint32_t n) {
    uint64_t i = pos >> 5;
    uint32_t o = (uint32_t)(pos & 31);
    uint32_t hi = bswap32(words[i]);
    uint32_t lo = bswap32(words[i + 1]);
    uint32_t v = o ? ((hi << o) | (lo >> (32 - o))) : hi;
    return v >> (32 - n);
}

// Sequential MSB-first reader over a word buffer. `end` is the first bit that must not be consumed;
// reads past it return zeros and set `overrun` (mirrors the panic of DefaultInputBitStream.pull, :268-294).
struct BitReader {
    const uint32_t* words;
    uint64_t pos;
    uint64_t end;
    bool overrun;
    KZ_HD BitReader(const uint32_t* w, uint64_t p, uint64_t e) : words(w), pos(p), end(e), overrun(false) {}
    KZ_HD uint32_t read(uint32_t n) {  // n in 1..32
        if (pos + n > end) {
            overrun = true;
            pos += n;
            return 0;
        }
        uint32_t v = bits_at(words, pos, n);
        pos += n;
        return v;
    }
    KZ_HD void skip(uint64_t n) {
        pos += n;
        if (pos > end) overrun = true;
    }
    // entropy/EntropyUtils.go:278-296
    KZ_HD uint32_t read_varint() {
        uint32_t res = 0, shift = 0;
        for (int i = 0; i < 4; i++) {
            uint32_t v = read(8);
            res |= (v & 0x7F) << shift;
            if (v < 128) return res;
            shift += 7;
        }
        uint32_t v = read(8);
        return res | ((v & 0x0F) << 28);
    }
};

// MSB-first writer into a zero-initialised, thread-private byte buffer (no atomics).
struct BitWriter {
    uint8_t* buf;
    uint32_t pos;  // bits
    KZ_HD BitWriter(uint8_t* b, uint32_t p = 0) : buf(b), pos(p) {}
    KZ_HD void write(uint32_t value, uint32_t n) {  // n in 1..32, buffer bytes beyond pos must be zero
        if (n < 32) value &= (1u << n) - 1u;
        uint32_t done = 0;
        while (done < n) {
            uint32_t byte = pos >> 3, off = pos & 7, room = 8 - off;
            uint32_t take = (n - done) < room ? (n - done) : room;
            uint32_t bits = (value >> (n - done - take)) & ((1u << take) - 1u);
            buf[byte] |= (uint8_t)(bits << (room - take));
            pos += take;
            done += take;
        }
    }
    // entropy/EntropyUtils.go:264-275
    KZ_HD void write_varint(uint32_t v) {
        while (v >= 128) {
            write(0x80 | (v & 0x7F), 8);
            v >>= 7;
        }
        write(v, 8);
    }
};

// ---- block framing shared by host and device (io/CompressedStream.go) --------------------------------------
struct BlockHeader {
    uint8_t mode;
    uint8_t skip_flags;
    uint32_t pre_len;       // "preTransformLength" on decode (= postTransformLength written by the encoder)
    uint64_t checksum;
    uint64_t data_bit;      // absolute bit offset of the entropy coded data
    int32_t status;         // 0 or -(error code)
    bool copy;
};

// decodingTask.decode :1878-1914
KZ_HD BlockHeader parse_block_header(const uint32_t* words, uint64_t bit_off, uint64_t bits, uint32_t checksum_bits, uint32_t block_size) {
    BlockHeader h;
    BitReader br(words, bit_off, bit_off + bits);
    h.status = 0;
    h.mode = (uint8_t)br.read(8);
    h.skip_flags = 0;
    h.copy = (h.mode & 0x80) != 0;
    if (!h.copy) {
        if (h.mode & 0x10) h.skip_flags = (uint8_t)br.read(8);
        else h.skip_flags = (uint8_t)((h.mode << 4) | 0x0F);
    }
    uint32_t data_size = 1 + ((h.mode >> 5) & 3);
    h.pre_len = br.read(8 * data_size);
    uint64_t max_len = (uint64_t)block_size + block_size / 2;
    if (max_len < 2048) max_len = 2048;
    if (max_len > (1ull << 30)) max_len = 1ull << 30;
    if (h.pre_len == 0 || h.pre_len > max_len) h.status = -KZ_E_BLOCK_SIZE;
    h.checksum = 0;
    if (checksum_bits == 32) h.checksum = br.read(32);
    else if (checksum_bits == 64) {
        uint64_t hi = br.read(32);
        h.checksum = (hi << 32) | br.read(32);
    }
    if (br.overrun) h.status = -KZ_E_PROCESS_BLOCK;
    h.data_bit = br.pos;
    return h;
}

// encodingTask.encode :836-887: mode byte, optional skip-flag byte, length, optional checksum. Returns bit count.
KZ_HD uint32_t make_block_header(uint8_t* dst /*>=14 zeroed bytes*/, bool copy, uint32_t ntransforms, uint8_t skip_flags, uint32_t post_len,
                                 uint32_t checksum_bits, uint64_t checksum) {
    uint32_t data_size = 1;
    if (post_len >= 256) data_size = (log2_floor(post_len) >> 3) + 1;
    uint8_t mode = copy ? 0x80 : 0;
    mode |= (uint8_t)(((data_size - 1) & 3) << 5);
    BitWriter bw(dst);
    if (copy || ntransforms <= 4) {
        mode |= (uint8_t)(skip_flags >> 4);
        bw.write(mode, 8);
    } else {
        mode |= 0x10;
        bw.write(mode, 8);
        bw.write(skip_flags, 8);
    }
    bw.write(post_len, 8 * data_size);
    if (checksum_bits == 32) bw.write((uint32_t)checksum, 32);
    else if (checksum_bits == 64) {
        bw.write((uint32_t)(checksum >> 32), 32);
        bw.write((uint32_t)checksum, 32);
    }
    return bw.pos;
}

}  // namespace kz
