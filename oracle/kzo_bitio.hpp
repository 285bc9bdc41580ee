// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// CPU restatement of kanzi-go's MSB-first bitstream semantics, used by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.
//
// Follows (semantics, not code):
//   v2/bitstream/DefaultOutputBitStream.go:64-96   WriteBit / WriteBits  (MSB first, 64-bit accumulator)
//   v2/bitstream/DefaultOutputBitStream.go:101-199 WriteArray (first `count` bits of a byte array;
//                                                   last partial byte contributes its HIGH bits, :194-196)
//   v2/bitstream/DefaultOutputBitStream.go:232-273 Close (zero padding of last byte) / Written (exact bit count)
//   v2/bitstream/DefaultInputBitStream.go:66-94    ReadBit / ReadBits
//   v2/bitstream/DefaultInputBitStream.go:99-211   ReadArray
//   v2/bitstream/DefaultInputBitStream.go:268-294  pull: reading past the end panics ("No more data to read")
//
// The reference streams through an io.Writer with a flush buffer; the bit sequence it produces is
// simply "all written bits, in order, MSB first, zero padded to a byte at Close". That sequence is
// what is restated here, over an in-memory byte vector.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kzo {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// kanzi error codes, v2/Definitions.go:25-46
enum {
    ERR_MISSING_PARAM = 1, ERR_BLOCK_SIZE = 2, ERR_INVALID_CODEC = 3, ERR_CREATE_COMPRESSOR = 4,
    ERR_CREATE_DECOMPRESSOR = 5, ERR_OUTPUT_IS_DIR = 6, ERR_OVERWRITE_FILE = 7, ERR_CREATE_FILE = 8,
    ERR_CREATE_BITSTREAM = 9, ERR_OPEN_FILE = 10, ERR_READ_FILE = 11, ERR_WRITE_FILE = 12,
    ERR_PROCESS_BLOCK = 13, ERR_CREATE_CODEC = 14, ERR_INVALID_FILE = 15, ERR_STREAM_VERSION = 16,
    ERR_CREATE_STREAM = 17, ERR_INVALID_PARAM = 18, ERR_CRC_CHECK = 19, ERR_UNKNOWN = 127
};

static inline uint64_t load_be64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return __builtin_bswap64(v);
}
static inline void store_be64(uint8_t* p, uint64_t v) {
    v = __builtin_bswap64(v);
    std::memcpy(p, &v, 8);
}

class BitWriter {
public:
    std::vector<uint8_t> buf;  // completed bytes
    uint64_t acc = 0;          // pending bits, left aligned (bit 63 = oldest)
    unsigned used = 0;         // number of pending bits in acc (0..63)

    void reserve(size_t n) { buf.reserve(n); }
    uint64_t written() const { return (uint64_t(buf.size()) << 3) + used; }

    inline void write_bit(int bit) { write_bits(uint64_t(bit & 1), 1); }

    // appends the low `count` bits of value, most significant first. count in [1..64]
    inline void write_bits(uint64_t value, unsigned count) {
        if (count == 0) return;
        if (count < 64) value &= ((uint64_t(1) << count) - 1);
        unsigned room = 64 - used;
        if (count < room) {
            acc |= value << (room - count);
            used += count;
            return;
        }
        // fill the accumulator, flush it, keep the rest
        unsigned rest = count - room;
        uint64_t full = acc | (rest == 64 ? 0 : (value >> rest));
        if (room == 64 && rest == 0) full = value;
        size_t n = buf.size();
        buf.resize(n + 8);
        store_be64(&buf[n], full);
        acc = rest ? (value << (64 - rest)) : 0;
        used = rest;
    }

    // appends the first `count` bits of the byte array
    void write_array(const uint8_t* bits, uint64_t count) {
        uint64_t nbytes = count >> 3;
        unsigned tail = unsigned(count & 7);
        if ((used & 7) == 0) {
            flush_whole_bytes();
            buf.insert(buf.end(), bits, bits + nbytes);
        } else if (nbytes >= 8) {
            // unaligned bulk path (DefaultOutputBitStream.go:144-185): shift whole 64-bit words through the accumulator
            const unsigned u = used;  // 1..63 pending bits, unchanged by whole words
            const uint64_t nwords = nbytes >> 3;
            size_t n = buf.size();
            buf.resize(n + size_t(nwords) * 8);
            uint8_t* out = buf.data() + n;
            uint64_t a = acc;
            for (uint64_t i = 0; i < nwords; i++) {
                const uint64_t v = load_be64(bits + 8 * i);
                store_be64(out + 8 * i, a | (v >> u));
                a = v << (64 - u);
            }
            acc = a;
            for (uint64_t i = nwords * 8; i < nbytes; i++) write_bits(bits[i], 8);
        } else {
            for (uint64_t i = 0; i < nbytes; i++) write_bits(bits[i], 8);
        }
        if (tail) write_bits(uint64_t(bits[nbytes]) >> (8 - tail), tail);
    }

    // zero-pads to a byte boundary and returns the bytes (Close semantics). Written() is unchanged by padding
    // in the reference (exact bit count); callers must sample written() before calling this.
    std::vector<uint8_t>& close() {
        while (used > 0) {
            buf.push_back(uint8_t(acc >> 56));
            acc <<= 8;
            used = used > 8 ? used - 8 : 0;
        }
        acc = 0;
        return buf;
    }

private:
    void flush_whole_bytes() {
        while (used >= 8) {
            buf.push_back(uint8_t(acc >> 56));
            acc <<= 8;
            used -= 8;
        }
    }
};

class BitReader {
public:
    const uint8_t* data;
    uint64_t nbits;  // total readable bits
    uint64_t pos = 0;

    BitReader(const uint8_t* d, uint64_t bits) : data(d), nbits(bits) {}
    // The reference's local block streams are built over whole bytes (data[0:r], r=(bits+7)>>3), so the
    // padding bits of the last byte are readable there too.
    static BitReader over_bytes(const uint8_t* d, size_t nbytes) { return BitReader(d, uint64_t(nbytes) << 3); }

    uint64_t read() const { return pos; }
    bool has_more() const { return pos < nbits; }

    inline int read_bit() { return int(read_bits(1)); }

    inline uint64_t read_bits(unsigned count) {
        if (count == 0 || count > 64) throw Error(ERR_PROCESS_BLOCK, "Invalid bit count");
        if (pos + count > nbits) throw Error(ERR_PROCESS_BLOCK, "No more data to read in the bitstream");
        uint64_t byte = pos >> 3;
        unsigned off = unsigned(pos & 7);
        uint64_t avail_bytes = ((nbits + 7) >> 3) - byte;
        uint64_t hi, lo = 0;
        if (avail_bytes >= 9) {
            hi = load_be64(data + byte);
            lo = data[byte + 8];
        } else {
            uint8_t tmp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            std::memcpy(tmp, data + byte, size_t(avail_bytes));
            hi = load_be64(tmp);
            lo = tmp[8];
        }
        uint64_t v = off ? ((hi << off) | (lo >> (8 - off))) : hi;
        pos += count;
        return count == 64 ? v : (v >> (64 - count));
    }

    // reads `count` bits into the byte array (last partial byte: bits land in its HIGH positions)
    void read_array(uint8_t* bits, uint64_t count) {
        if (count == 0) return;
        if (pos + count > nbits) throw Error(ERR_PROCESS_BLOCK, "No more data to read in the bitstream");
        uint64_t nbytes = count >> 3;
        unsigned tail = unsigned(count & 7);
        if ((pos & 7) == 0) {
            std::memcpy(bits, data + (pos >> 3), size_t(nbytes));
            pos += nbytes << 3;
        } else {
            const unsigned off = unsigned(pos & 7);
            const uint8_t* p = data + (pos >> 3);
            const uint64_t total_bytes = (nbits + 7) >> 3;
            uint64_t i = 0;
            // bulk: needs 9 readable bytes per word
            const uint64_t avail = total_bytes - (pos >> 3);
            for (; i + 8 <= nbytes && i + 9 <= avail; i += 8) {
                const uint64_t hi = load_be64(p + i);
                store_be64(bits + i, (hi << off) | (uint64_t(p[i + 8]) >> (8 - off)));
            }
            pos += i << 3;
            for (; i < nbytes; i++) bits[i] = uint8_t(read_bits(8));
        }
        if (tail) bits[nbytes] = uint8_t(read_bits(tail) << (8 - tail));
    }
};

}  // namespace kzo
