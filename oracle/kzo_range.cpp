// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Order-0 range coder. Follows v2/entropy/RangeCodec.go.
#include <algorithm>
#include <cstring>

#include "kzo.hpp"

namespace kzo {
namespace {
const uint64_t TOP_RANGE = 0x0FFFFFFFFFFFFFFFull;     // RangeCodec.go:31
const uint64_t BOTTOM_RANGE = 0x000000000000FFFFull;  // :32
const uint64_t RANGE_MASK = 0x0FFFFFFF00000000ull;    // :33
const size_t RANGE_CHUNK = 1 << 15;                   // :34
const unsigned RANGE_LOG = 12;                        // :35

// :158-210 (same group layout as the ANS header, but written AFTER the alphabet and only if it is non-empty)
void range_encode_header(BitWriter& bw, const int* alphabet, int asz, const int* f, unsigned lr) {
    encode_alphabet(bw, alphabet, asz);
    if (asz == 0) return;
    bw.write_bits(lr - 8, 3);
    int chk = asz < 64 ? 6 : 8;
    unsigned llr = 3;
    while ((1u << llr) <= lr) llr++;
    for (int i = 1; i < asz; i += chk) {
        int mx = f[alphabet[i]] - 1;
        unsigned log_max = 0;
        int endj = std::min(i + chk, asz);
        for (int j = i + 1; j < endj; j++) mx = std::max(mx, f[alphabet[j]] - 1);
        while ((1 << log_max) <= mx) log_max++;
        bw.write_bits(log_max, llr);
        if (log_max == 0) continue;
        for (int j = i; j < endj; j++) bw.write_bits(uint64_t(f[alphabet[j]] - 1), log_max);
    }
}
}  // namespace

// :216-305
void range_encode(BitWriter& bw, const uint8_t* block, size_t len) {
    if (len == 0) return;
    int alphabet[256];
    int freqs[256];
    uint64_t cum[257];
    size_t start = 0;
    while (start < len) {
        uint64_t rng = TOP_RANGE, low = 0;
        unsigned lr = RANGE_LOG;
        size_t end = std::min(start + RANGE_CHUNK, len);
        while (lr > 8 && (size_t(1) << lr) > end - start) lr--;
        const uint8_t* buf = block + start;
        size_t n = end - start;
        std::memset(freqs, 0, sizeof(freqs));
        std::memset(alphabet, 0, sizeof(alphabet));
        histogram_order0(buf, n, freqs);
        int asz = normalize_frequencies(freqs, alphabet, 256, int(n), 1 << lr);
        if (asz > 0) {
            cum[0] = 0;
            for (int i = 0; i < 256; i++) cum[i + 1] = cum[i] + uint64_t(freqs[i]);
        }
        range_encode_header(bw, alphabet, asz, freqs, lr);
        if (asz <= 1) {
            start = end;
            continue;
        }
        for (size_t i = 0; i < n; i++) {  // :283-304 encodeByte
            int s = buf[i];
            uint64_t c = cum[s];
            rng >>= lr;
            low += c * rng;
            rng *= (cum[s + 1] - c);
            for (;;) {
                if (((low ^ (low + rng)) & RANGE_MASK) != 0) {
                    if (rng > BOTTOM_RANGE) break;
                    rng = (0 - low) & BOTTOM_RANGE;
                }
                bw.write_bits(low >> 32, 28);
                rng <<= 28;
                low <<= 28;
            }
        }
        bw.write_bits(low, 60);
        start = end;
    }
}

// :403-557
void range_decode(BitReader& br, uint8_t* block, size_t len) {
    int alphabet[256];
    int freqs[256];
    uint64_t cum[257];
    std::vector<uint16_t> f2s;
    std::memset(freqs, 0, sizeof(freqs));
    size_t start = 0;
    while (start < len) {
        size_t end = std::min(start + RANGE_CHUNK, len);
        int asz = decode_alphabet(br, alphabet);
        if (asz == 0) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: empty alphabet");  // reference: silent short read
        if (asz != 256) std::memset(freqs, 0, sizeof(freqs));
        unsigned lr = unsigned(8 + br.read_bits(3));
        int scale = 1 << lr;
        int sum = 0;
        int chk = asz < 64 ? 6 : 8;
        unsigned llr = 3;
        while ((1u << llr) <= lr) llr++;
        for (int i = 1; i < asz; i += chk) {
            unsigned log_max = unsigned(br.read_bits(llr));
            if ((1 << log_max) > scale) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect frequency size");
            int endj = std::min(i + chk, asz);
            for (int j = i; j < endj; j++) {
                int freq = 1;
                if (log_max > 0) {
                    freq = int(1 + br.read_bits(log_max));
                    if (freq <= 0 || freq >= scale) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect frequency");
                }
                freqs[alphabet[j]] = freq;
                sum += freq;
            }
        }
        if (scale <= sum) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect frequency");
        freqs[alphabet[0]] = scale - sum;
        cum[0] = 0;
        if (f2s.size() < size_t(scale)) f2s.assign(size_t(scale), 0);
        for (int i = 0; i < 256; i++) {
            cum[i + 1] = cum[i] + uint64_t(freqs[i]);
            if (cum[i + 1] > uint64_t(scale)) throw Error(ERR_PROCESS_BLOCK, "index out of range");
            for (uint64_t j = cum[i]; j < cum[i + 1]; j++) f2s[size_t(j)] = uint16_t(i);
        }
        if (asz == 1) {
            std::memset(block + start, alphabet[0], end - start);
            start = end;
            continue;
        }
        uint64_t rng = TOP_RANGE, low = 0;
        uint64_t code = br.read_bits(60);
        for (size_t i = start; i < end; i++) {  // :530-557 decodeByte
            rng >>= lr;
            uint64_t count = (code - low) / rng;
            if (count >= uint64_t(scale)) throw Error(ERR_PROCESS_BLOCK, "index out of range");
            int s = f2s[size_t(count)];
            uint64_t c = cum[s];
            low += c * rng;
            rng *= (cum[s + 1] - c);
            for (;;) {
                if (((low ^ (low + rng)) & RANGE_MASK) != 0) {
                    if (rng > BOTTOM_RANGE) break;
                    rng = (0 - low) & BOTTOM_RANGE;
                }
                code = (code << 28) | br.read_bits(28);
                rng <<= 28;
                low <<= 28;
            }
            block[i] = uint8_t(s);
        }
        start = end;
    }
}

}  // namespace kzo
