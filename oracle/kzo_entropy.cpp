// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Entropy utilities + static rANS (order 0 / order 1).
#include <algorithm>
#include <cstring>

#include "kzo.hpp"

namespace kzo {

// ---------------------------------------------------------------------------------------------
// v2/internal/Global.go:156-172 Log2NoCheck (floor(log2 x), x > 0)
uint32_t log2_floor(uint32_t x) { return 31u - uint32_t(__builtin_clz(x)); }

// v2/internal/Global.go:175-191 Log2ScaledBy1024 — needs the LOG2_4096 table semantics: 4096*log2(x/128), x in [0..255].
// The table is data, not arithmetic; it is regenerated here from its definition and rounded like the reference's.
static uint32_t log2_4096_entry(uint32_t i) {
    // reference table LOG2_4096[i] = round(4096*log2(i)) for i >= 1 (entry 0 = 0)
    static uint32_t tab[257];
    static bool init = false;
    if (!init) {
        tab[0] = 0;
        for (int k = 1; k <= 256; k++) {
            long double v = 4096.0L * __builtin_log2l((long double)k);
            tab[k] = uint32_t(v + 0.5L);
        }
        init = true;
    }
    return tab[i];
}
static uint32_t log2_scaled_1024(uint32_t x) {
    if (x < 256) return (log2_4096_entry(x) + 2) >> 2;
    uint32_t lg = log2_floor(x);
    if ((x & (x - 1)) == 0) return lg << 10;
    return ((lg - 7) * 1024) + ((log2_4096_entry(x >> (lg - 7)) + 2) >> 2);
}
// v2/internal/Global.go:196-214
int first_order_entropy_1024(size_t len, const int* histo) {
    if (len == 0) return 0;
    uint64_t sum = 0;
    uint32_t ll = log2_scaled_1024(uint32_t(len));
    for (int i = 0; i < 256; i++) {
        if (histo[i] == 0) continue;
        uint32_t l = log2_scaled_1024(uint32_t(histo[i]));
        sum += (uint64_t(histo[i]) * uint64_t(ll - l)) >> 3;
    }
    return int(sum / uint64_t(len));
}

// v2/internal/Global.go:220-251 (order 0). The unrolling of the reference does not change the counts.
void histogram_order0(const uint8_t* p, size_t n, int* freqs) {
    for (size_t i = 0; i < n; i++) freqs[p[i]]++;
}
// v2/internal/Global.go:252-303 (order 1, withTotal): every path of the reference counts the pairs
// (prev, cur) with prev = 0 for the first byte of the slice and the true preceding byte elsewhere
// (:262-268 short path; :270-273 the 4 cursors start with 0 / block[n_k - 1]; :298-302 tail).
void histogram_order1(const uint8_t* p, size_t n, int* freqs) {
    unsigned prv = 0;
    for (size_t i = 0; i < n; i++) {
        freqs[257 * prv + p[i]]++;
        freqs[257 * prv + 256]++;
        prv = p[i];
    }
}

// ---------------------------------------------------------------------------------------------
// v2/entropy/EntropyUtils.go:38-67
int encode_alphabet(BitWriter& bw, const int* alphabet, int count) {
    if (count > 256) throw Error(ERR_PROCESS_BLOCK, "The max alphabet length is 256");
    if (count == 0) {
        bw.write_bit(0);  // _FULL_ALPHABET
        bw.write_bit(1);  // _ALPHABET_0
    } else if (count == 256) {
        bw.write_bit(0);
        bw.write_bit(0);  // _ALPHABET_256
    } else {
        bw.write_bit(1);  // _PARTIAL_ALPHABET
        uint8_t masks[32];
        std::memset(masks, 0, sizeof(masks));
        for (int i = 0; i < count; i++) masks[alphabet[i] >> 3] |= uint8_t(1u << (alphabet[i] & 7));
        int last = alphabet[count - 1] >> 3;
        bw.write_bits(uint64_t(last), 5);
        bw.write_array(masks, 8 * uint64_t(last + 1));
    }
    return count;
}

// v2/entropy/EntropyUtils.go:71-118
int decode_alphabet(BitReader& br, int* alphabet) {
    if (br.read_bit() == 0) {
        if (br.read_bit() == 1) return 0;
        for (int i = 0; i < 256; i++) alphabet[i] = i;
        return 256;
    }
    int last = int(br.read_bits(5));
    uint8_t masks[32];
    std::memset(masks, 0, sizeof(masks));
    br.read_array(masks, 8 * uint64_t(last + 1));
    int count = 0;
    for (int i = 0; i <= last; i++)
        for (int j = 0; j < 8; j++)
            if ((masks[i] >> j) & 1) alphabet[count++] = 8 * i + j;
    return count;
}

// v2/entropy/EntropyUtils.go:123-260
int normalize_frequencies(int* freqs, int* alphabet, int alphabet_len, int total, int scale) {
    if (alphabet_len > 256) throw Error(ERR_PROCESS_BLOCK, "Invalid alphabet size parameter");
    if (scale < 256 || scale > 65536) throw Error(ERR_PROCESS_BLOCK, "Invalid range parameter");
    if (alphabet_len == 0 || total == 0) return 0;
    int asz = 0;
    if (total == scale) {  // :139-148 (indexes all 256 entries regardless of len(alphabet))
        if (alphabet_len < 256) throw Error(ERR_PROCESS_BLOCK, "index out of range");
        for (int i = 0; i < 256; i++)
            if (freqs[i] != 0) alphabet[asz++] = i;
        return asz;
    }
    int sum_scaled = 0, sum = 0, idx_max = 0;
    for (int i = 0; i < alphabet_len; i++) {  // :155-186
        alphabet[i] = 0;
        int f = freqs[i];
        if (f == 0) continue;
        int64_t sf = int64_t(f) * int64_t(scale);
        int scaled = (sf <= int64_t(total)) ? 1 : int((sf + (int64_t(total) >> 1)) / int64_t(total));
        alphabet[asz++] = i;
        sum_scaled += scaled;
        freqs[i] = scaled;
        sum += f;
        if (scaled > freqs[idx_max]) idx_max = i;
        if (sum >= total) break;
    }
    if (asz == 0) return 0;
    if (asz == 1) {
        freqs[alphabet[0]] = scale;
        return 1;
    }
    if (sum_scaled == scale) return asz;
    int delta = sum_scaled - scale;
    int err_thr = freqs[idx_max] >> 4;
    int abs_delta = delta < 0 ? -delta : delta;
    if (abs_delta <= err_thr) {  // :211-215
        freqs[idx_max] -= delta;
        return asz;
    }
    int inc;
    if (delta < 0) {  // :217-226
        delta += err_thr;
        freqs[idx_max] += err_thr;
        inc = 1;
        delta = -delta;
    } else {
        delta -= err_thr;
        freqs[idx_max] -= err_thr;
        inc = -1;
    }
    int round = 1;
    while (round < 6 && delta > 0) {  // :232-256
        int adjustments = 0;
        round++;
        for (int k = 0; k < asz; k++) {
            int idx = alphabet[k];
            if (freqs[idx] <= 2) continue;
            freqs[idx] += inc;
            adjustments++;
            delta--;
            if (delta == 0) break;
        }
        if (adjustments == 0) break;
    }
    freqs[idx_max] = std::max(freqs[idx_max] - delta, 1);  // :258
    return asz;
}

// v2/entropy/EntropyUtils.go:264-275
int write_varint(BitWriter& bw, uint32_t v) {
    int res = 1;
    while (v >= 128) {
        bw.write_bits(0x80 | (v & 0x7F), 8);
        v >>= 7;
        res++;
    }
    bw.write_bits(v, 8);
    return res;
}
// v2/entropy/EntropyUtils.go:278-296
uint32_t read_varint(BitReader& br) {
    uint32_t res = 0;
    unsigned shift = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t v = uint32_t(br.read_bits(8));
        res |= (v & 0x7F) << shift;
        if (v < 128) return res;
        shift += 7;
    }
    uint32_t v = uint32_t(br.read_bits(8));
    return res | ((v & 0x0F) << 28);
}

// ---------------------------------------------------------------------------------------------
// rANS. v2/entropy/ANSRangeCodec.go
namespace {
const int ANS_TOP = 1 << 15;                // :32
const int ANS0_CHUNK = 16384;               // :33
const int ANS_MAX_CHUNK = 1 << 27;          // :35

struct EncSym {  // :438-468
    int64_t x_max, bias, cmpl_freq;
    uint64_t inv_freq;
    unsigned inv_shift;
    void reset(int cum, int freq, unsigned lr) {
        freq = std::min(freq, (1 << lr) - 1);
        x_max = int64_t((ANS_TOP >> lr) << 16) * freq;
        cmpl_freq = (1 << lr) - freq;
        if (freq < 2) {
            inv_freq = 0xFFFFFFFFull;
            inv_shift = 32;
            bias = cum + (1 << lr) - 1;
        } else {
            unsigned shift = 0;
            while (freq > (1 << shift)) shift++;
            inv_freq = (((uint64_t(1) << (shift + 31)) + uint64_t(freq - 1)) / uint64_t(freq)) & 0xFFFFFFFFull;
            inv_shift = 32 + shift - 1;
            bias = cum;
        }
    }
};

struct AnsEncoder {
    BitWriter& bw;
    int order;
    unsigned lr;
    int chunk;
    std::vector<int> freqs;       // dim*257
    std::vector<EncSym> symbols;  // dim*256
    std::vector<uint8_t> buffer;

    AnsEncoder(BitWriter& w, int ord, int chunk_size, int log_range) : bw(w), order(ord) {  // :117-168
        int chk = chunk_size > 0 ? chunk_size : ANS0_CHUNK;
        if (order == 1) chk = std::min(chk << 8, ANS_MAX_CHUNK);
        int dim = 255 * order + 1;
        freqs.assign(size_t(dim) * 257, 0);
        symbols.resize(size_t(dim) * 256);
        lr = unsigned(std::max(log_range - order, 8));
        chunk = chk;
    }

    // :216-270
    void encode_header(const int* alphabet, int asz, const int* f, unsigned lr_) {
        encode_alphabet(bw, alphabet, asz);
        if (asz <= 1) return;
        int chk = asz < 64 ? 6 : 8;
        unsigned llr = 3;
        while ((1u << llr) <= lr_) llr++;
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

    // :171-213
    int update_frequencies(unsigned lr_) {
        int res = 0;
        int endk = 255 * order + 1;
        bw.write_bits(lr_ - 8, 3);
        int alphabet[256];
        std::memset(alphabet, 0, sizeof(alphabet));
        for (int k = 0; k < endk; k++) {
            int* f = &freqs[size_t(257) * k];
            EncSym* symb = &symbols[size_t(k) << 8];
            int asz = normalize_frequencies(f, alphabet, 256, f[256], 1 << lr_);
            if (asz > 0) {
                int sum = 0;
                for (int i = 0, count = 0; i < 256; i++) {
                    if (f[i] == 0) continue;
                    symb[i].reset(sum, f[i], lr_);
                    sum += f[i];
                    if (++count >= asz) break;
                }
            }
            encode_header(alphabet, asz, f, lr_);
            res += asz;
        }
        return res;
    }

    // :408-427
    int rebuild_statistics(const uint8_t* block, size_t n) {
        std::fill(freqs.begin(), freqs.end(), 0);
        if (order == 0) {
            histogram_order0(block, n, freqs.data());
            freqs[256] = int(n);
        } else {
            size_t q = n >> 2;
            if (q == 0) {
                histogram_order1(block, n, freqs.data());
            } else {
                for (int k = 0; k < 4; k++) histogram_order1(block + k * q, q, freqs.data());
            }
        }
        return update_frequencies(lr);
    }

    // :313-329
    inline void encode_symbol(int64_t& n, int64_t& st, const EncSym& s) {
        int x = st >= s.x_max ? 1 : 0;
        buffer[size_t(n)] = uint8_t(st);
        n -= x;
        buffer[size_t(n)] = uint8_t(st >> 8);
        n -= x;
        if (x) st >>= 16;
        st = st + s.bias + int64_t((uint64_t(st) * s.inv_freq) >> s.inv_shift) * s.cmpl_freq;
    }

    // :331-405
    void encode_chunk(const uint8_t* block, int64_t len) {
        int64_t st0 = ANS_TOP, st1 = ANS_TOP, st2 = ANS_TOP, st3 = ANS_TOP;
        int64_t n = int64_t(buffer.size()) - 1;
        int64_t end4 = len & -4;
        for (int64_t i = len - 1; i >= end4; i--) buffer[size_t(n--)] = block[i];
        if (order == 0) {
            const EncSym* symb = symbols.data();
            for (int64_t i = end4 - 1; i > 0; i -= 4) {
                encode_symbol(n, st0, symb[block[i]]);
                encode_symbol(n, st1, symb[block[i - 1]]);
                encode_symbol(n, st2, symb[block[i - 2]]);
                encode_symbol(n, st3, symb[block[i - 3]]);
            }
        } else if (len > 1) {
            int64_t quarter = end4 >> 2;
            int64_t i0 = quarter - 2, i1 = 2 * quarter - 2, i2 = 3 * quarter - 2, i3 = end4 - 2;
            // :359 block[i0+1] with quarter == 0 indexes block[-1] in the reference -> Go panic
            if (quarter == 0) throw Error(ERR_PROCESS_BLOCK, "index out of range [-1]");
            int prv0 = block[i0 + 1], prv1 = block[i1 + 1], prv2 = block[i2 + 1], prv3 = block[i3 + 1];
            while (i0 >= 0) {
                int c0 = block[i0], c1 = block[i1], c2 = block[i2], c3 = block[i3];
                encode_symbol(n, st0, symbols[size_t(c0 << 8) | prv0]);
                encode_symbol(n, st1, symbols[size_t(c1 << 8) | prv1]);
                encode_symbol(n, st2, symbols[size_t(c2 << 8) | prv2]);
                encode_symbol(n, st3, symbols[size_t(c3 << 8) | prv3]);
                prv0 = c0; prv1 = c1; prv2 = c2; prv3 = c3;
                i0--; i1--; i2--; i3--;
            }
            encode_symbol(n, st0, symbols[size_t(prv0)]);
            encode_symbol(n, st1, symbols[size_t(prv1)]);
            encode_symbol(n, st2, symbols[size_t(prv2)]);
            encode_symbol(n, st3, symbols[size_t(prv3)]);
        }
        n++;
        uint32_t sz = uint32_t(int64_t(buffer.size()) - n);
        write_varint(bw, sz);
        bw.write_bits(uint64_t(st0), 32);
        bw.write_bits(uint64_t(st1), 32);
        bw.write_bits(uint64_t(st2), 32);
        bw.write_bits(uint64_t(st3), 32);
        if (sz != 0) bw.write_array(&buffer[size_t(n)], 8 * uint64_t(sz));
    }

    // :274-311
    void write(const uint8_t* block, size_t len) {
        if (len <= 32) {
            bw.write_array(block, 8 * uint64_t(len));
            return;
        }
        size_t size = std::min<size_t>(2 * len, size_t(chunk) + (size_t(chunk) >> 3));
        size = std::max<size_t>(size, 65536);
        if (buffer.size() < size) buffer.assign(size, 0);
        size_t start = 0;
        while (start < len) {
            size_t end = std::min(start + size_t(chunk), len);
            int asz = rebuild_statistics(block + start, end - start);
            if (order == 1 || asz > 1) encode_chunk(block + start, int64_t(end - start));
            start = end;
        }
    }
};

struct DecSym {  // :968-978
    int cum, freq;
};

struct AnsDecoder {
    BitReader& br;
    int order;
    unsigned lr = 12;
    int chunk;
    std::vector<int> freqs;  // dim*256
    std::vector<DecSym> symbols;
    std::vector<uint8_t> f2s;
    std::vector<uint8_t> buffer;

    AnsDecoder(BitReader& r, int ord, int chunk_size) : br(r), order(ord) {  // :542-602
        int chk = chunk_size > 0 ? chunk_size : ANS0_CHUNK;
        if (order == 1) chk = std::min(chk << 8, ANS_MAX_CHUNK);
        int dim = 255 * order + 1;
        freqs.assign(size_t(dim) * 256, 0);
        symbols.assign(size_t(dim) * 256, DecSym{0, 0});
        chunk = chk;
    }

    // :605-710
    int decode_header(int* alphabet) {
        lr = unsigned(8 + br.read_bits(3));
        if (lr < 8 || lr > 16) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: range");
        int res = 0;
        int dim = 255 * order + 1;
        int scale = 1 << lr;
        if (f2s.size() < size_t(dim) * scale) f2s.assign(size_t(dim) * scale, 0);
        unsigned llr = 3;
        while ((1u << llr) <= lr) llr++;
        for (int k = 0; k < dim; k++) {
            int asz = decode_alphabet(br, alphabet);
            if (asz == 0) continue;
            int* f = &freqs[size_t(k) << 8];
            if (asz != 256) std::memset(f, 0, 256 * sizeof(int));
            int chk = asz < 64 ? 6 : 8;
            int sum = 0;
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
                    f[alphabet[j]] = freq;
                    sum += freq;
                }
            }
            if (scale <= sum) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect frequency");
            f[alphabet[0]] = scale - sum;
            sum = 0;
            DecSym* symb = &symbols[size_t(k) << 8];
            uint8_t* freq2sym = &f2s[size_t(k) << lr];
            for (int i = 0; i < 256; i++) {
                if (f[i] == 0) continue;
                // a corrupt header can make sum+f[i] exceed scale; the reference would index out of range (panic)
                if (sum + f[i] > scale) throw Error(ERR_PROCESS_BLOCK, "index out of range");
                std::memset(freq2sym + sum, i, size_t(f[i]));
                symb[i].cum = sum;
                symb[i].freq = std::min(f[i], (1 << lr) - 1);
                sum += f[i];
            }
            res += asz;
        }
        return res;
    }

    // :846-858
    inline void decode_symbol(size_t& n, int64_t& st, const DecSym& s, int64_t mask) {
        st = int64_t(s.freq) * (st >> lr) + (st & mask) - s.cum;
        if (st < ANS_TOP) {
            st = (st << 16) | (int64_t(buffer[n]) << 8) | int64_t(buffer[n + 1]);
            n += 2;
        }
    }

    // :860-957
    bool decode_chunk(uint8_t* block, size_t len) {
        uint32_t sz = read_varint(br);
        if (sz >= uint32_t(ANS_MAX_CHUNK)) return false;
        int64_t st0 = int64_t(br.read_bits(32)), st1 = int64_t(br.read_bits(32));
        int64_t st2 = int64_t(br.read_bits(32)), st3 = int64_t(br.read_bits(32));
        if (len == 0) return true;
        size_t min_buf = std::max<size_t>(2 * len, 256);
        if (buffer.size() < min_buf) buffer.assign(min_buf, 0);
        // the reference would panic (slice bounds) if 8*sz exceeds the buffer; treated as a block error
        if (size_t(sz) > buffer.size()) throw Error(ERR_PROCESS_BLOCK, "slice bounds out of range");
        br.read_array(buffer.data(), 8 * uint64_t(sz));
        if (size_t(sz) < buffer.size()) {
            size_t guard_end = std::min(size_t(sz) + 64, buffer.size());
            std::memset(&buffer[sz], 0, guard_end - sz);
        }
        size_t n = 0;
        int64_t mask = (int64_t(1) << lr) - 1;
        size_t end4 = len & ~size_t(3);
        if (order == 0) {
            const uint8_t* freq2sym = f2s.data();
            const DecSym* symb = symbols.data();
            for (size_t i = 0; i < end4; i += 4) {
                uint8_t c3 = freq2sym[st3 & mask]; block[i] = c3;     decode_symbol(n, st3, symb[c3], mask);
                uint8_t c2 = freq2sym[st2 & mask]; block[i + 1] = c2; decode_symbol(n, st2, symb[c2], mask);
                uint8_t c1 = freq2sym[st1 & mask]; block[i + 2] = c1; decode_symbol(n, st1, symb[c1], mask);
                uint8_t c0 = freq2sym[st0 & mask]; block[i + 3] = c0; decode_symbol(n, st0, symb[c0], mask);
            }
        } else {
            size_t quarter = end4 >> 2;
            size_t i0 = 0, i1 = quarter, i2 = 2 * quarter, i3 = 3 * quarter;
            int p0 = 0, p1 = 0, p2 = 0, p3 = 0;
            while (i0 < quarter) {
                uint8_t c3 = f2s[(size_t(p3) << lr) + size_t(st3 & mask)]; block[i3] = c3; decode_symbol(n, st3, symbols[(size_t(p3) << 8) + c3], mask);
                uint8_t c2 = f2s[(size_t(p2) << lr) + size_t(st2 & mask)]; block[i2] = c2; decode_symbol(n, st2, symbols[(size_t(p2) << 8) + c2], mask);
                uint8_t c1 = f2s[(size_t(p1) << lr) + size_t(st1 & mask)]; block[i1] = c1; decode_symbol(n, st1, symbols[(size_t(p1) << 8) + c1], mask);
                uint8_t c0 = f2s[(size_t(p0) << lr) + size_t(st0 & mask)]; block[i0] = c0; decode_symbol(n, st0, symbols[(size_t(p0) << 8) + c0], mask);
                p3 = c3; p2 = c2; p1 = c1; p0 = c0;
                i0++; i1++; i2++; i3++;
            }
        }
        for (size_t i = end4; i < len; i++) block[i] = buffer[n++];
        return true;
    }

    // :714-757
    void read(uint8_t* block, size_t len) {
        if (len <= 32) {
            br.read_array(block, 8 * uint64_t(len));
            return;
        }
        size_t start = 0;
        int alphabet[256];
        while (start < len) {
            size_t end = std::min(start + size_t(chunk), len);
            int asz = decode_header(alphabet);
            if (asz == 0) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: empty alphabet");  // reference: silent short read
            if (order == 0 && asz == 1) {
                std::memset(block + start, alphabet[0], end - start);
            } else if (!decode_chunk(block + start, end - start)) {
                throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect chunk size");
            }
            start = end;
        }
    }
};
}  // namespace

void ans_encode(BitWriter& bw, const uint8_t* block, size_t n, int order, int chunk_size, int log_range) {
    AnsEncoder e(bw, order, chunk_size, log_range);
    e.write(block, n);
}
void ans_decode(BitReader& br, uint8_t* block, size_t n, int order, int chunk_size) {
    AnsDecoder d(br, order, chunk_size);
    d.read(block, n);
}

// v2/entropy/NullEntropyCodec.go:43-61, :91-109
void null_encode(BitWriter& bw, const uint8_t* block, size_t n) { bw.write_array(block, 8 * uint64_t(n)); }
void null_decode(BitReader& br, uint8_t* block, size_t n) { br.read_array(block, 8 * uint64_t(n)); }

// v2/entropy/EntropyCodecFactory.go:45-134
void entropy_encode(uint32_t type, BitWriter& bw, const uint8_t* block, size_t n) {
    switch (type) {
        case E_NONE: null_encode(bw, block, n); break;
        case E_HUFFMAN: huffman_encode(bw, block, n); break;
        case E_ANS0: ans_encode(bw, block, n, 0); break;
        case E_ANS1: ans_encode(bw, block, n, 1); break;
        case E_RANGE: range_encode(bw, block, n); break;
        default: throw Error(ERR_CREATE_CODEC, "Unsupported entropy codec type");
    }
}
void entropy_decode(uint32_t type, BitReader& br, uint8_t* block, size_t n) {
    switch (type) {
        case E_NONE: null_decode(br, block, n); break;
        case E_HUFFMAN: huffman_decode(br, block, n); break;
        case E_ANS0: ans_decode(br, block, n, 0); break;
        case E_ANS1: ans_decode(br, block, n, 1); break;
        case E_RANGE: range_decode(br, block, n); break;
        default: throw Error(ERR_INVALID_CODEC, "Unsupported entropy codec type");
    }
}

}  // namespace kzo
