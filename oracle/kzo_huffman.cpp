// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp). Canonical length-limited Huffman, bitstream v6 layout.
// Follows v2/entropy/HuffmanCodec.go and v2/entropy/ExpGolombCodec.go (signed variant).
#include <algorithm>
#include <cstring>

#include "kzo.hpp"

namespace kzo {
namespace {
const int HUF_CHUNK = 1 << 14;     // HuffmanCodec.go:30
const int HUF_MAX_BITS = 12;       // :31

// ExpGolombCodec.go:104-112 with the signed table (:45-62). The table entry for an int8 value v != 0 is
// (length << 9) | bits with: m = |v|, k = floor(log2(m+1)), bits = [0]*k ++ [1] ++ (m+1-2^k in k bits) ++ sign,
// length = 2k+2 — derived from the decoder (:159-187) and checked against the table by tests/test_oracle_pins.py.
void expgolomb_signed_encode(BitWriter& bw, uint8_t val) {
    if (val == 0) {
        bw.write_bit(1);
        return;
    }
    int v = int(int8_t(val));
    unsigned m = unsigned(v < 0 ? -v : v);
    unsigned k = log2_floor(m + 1);
    uint64_t bits = (uint64_t(1) << (k + 1)) | (uint64_t(m + 1 - (1u << k)) << 1) | uint64_t(v < 0 ? 1 : 0);
    bw.write_bits(bits & 0x1FF, 2 * k + 2);
}
// ExpGolombCodec.go:159-187
uint8_t expgolomb_signed_decode(BitReader& br) {
    if (br.read_bit() == 1) return 0;
    unsigned lg = 1;
    while (br.read_bit() != 1) lg++;
    lg &= 7;
    uint64_t val = br.read_bits(lg + 1);
    uint64_t res = (val >> 1) + (uint64_t(1) << lg) - 1;
    if (val & 1) res = ~res + 1;
    return uint8_t(res);
}

// HuffmanCodec.go:37-77. symbols[] is re-ordered by (size, symbol); codes are canonical.
int generate_canonical_codes(const uint8_t* sizes, uint16_t* codes, int* symbols, int count, int max_size) {
    if (count == 0) return 0;
    if (count > 1) {
        static thread_local uint8_t buf[(HUF_MAX_BITS << 8) + 256];
        std::memset(buf, 0, sizeof(buf));
        for (int i = 0; i < count; i++) {
            int s = symbols[i];
            if (s > 255) throw Error(ERR_PROCESS_BLOCK, "Could not generate Huffman codes: invalid code length");
            if (sizes[s] > max_size) throw Error(ERR_PROCESS_BLOCK, "Could not generate Huffman codes: max code length exceeded");
            if (sizes[s] == 0) throw Error(ERR_PROCESS_BLOCK, "index out of range");  // (sizes-1)<<8 underflows in the reference
            buf[(int(sizes[s] - 1) << 8) | s] = 1;
        }
        for (int i = 0, n = 0; n < count; i++) {
            symbols[n] = i & 0xFF;
            n += buf[i];
        }
    }
    uint16_t code = 0;
    uint8_t cur = sizes[symbols[0]];
    for (int i = 0; i < count; i++) {
        int s = symbols[i];
        code = uint16_t(code << (sizes[s] - cur));
        cur = sizes[s];
        codes[s] = code;
        code++;
    }
    return count;
}

// :326-351 (Moffat & Katajainen, phase 1)
void inplace_sizes_phase1(int* data, int n) {
    for (int s = 0, r = 0, t = 0; t < n - 1; t++) {
        int sum = 0;
        for (int i = 0; i < 2; i++) {
            if (s >= n || (r < t && data[r] < data[s])) {
                sum += data[r];
                data[r] = t;
                r++;
                continue;
            }
            sum += data[s];
            if (s > t) data[s] = 0;
            s++;
        }
        data[t] = sum;
    }
}
// :354-385
int inplace_sizes_phase2(int* data, int n) {
    if (n < 2) return 0;
    int level_top = n - 2, depth = 1, i = n, total_nodes = 2;
    while (i > 0) {
        int k = level_top;
        while (k > 0 && data[k - 1] >= level_top) k--;
        int internal_nodes = level_top - k;
        int leaves = total_nodes - internal_nodes;
        for (int j = 0; j < leaves; j++) data[--i] = depth;
        total_nodes = internal_nodes << 1;
        level_top = k;
        depth++;
    }
    return depth - 1;
}

// :300-324. ranks: (freq<<8)|symbol on entry, symbols sorted by increasing (freq, symbol) on exit.
int compute_code_lengths(uint8_t* sizes, int* ranks, int count) {
    int f[256];
    std::sort(ranks, ranks + count);
    for (int i = 0; i < count; i++) {
        f[i] = ranks[i] >> 8;
        ranks[i] &= 0xFF;
        if (f[i] == 0) throw Error(ERR_PROCESS_BLOCK, "Could not generate Huffman codes: invalid code length 0");
    }
    inplace_sizes_phase1(f, count);
    int max_len = inplace_sizes_phase2(f, count);
    for (int i = 0; i < count; i++) sizes[ranks[i]] = uint8_t(f[i]);
    return max_len;
}

// :216-297
int limit_code_lengths(const int* symbols, int* freqs, uint8_t* sizes, int* ranks, int count) {
    int n = 0, debt = 0;
    while (sizes[ranks[n]] >= HUF_MAX_BITS) {
        debt += int(sizes[ranks[n]]) - HUF_MAX_BITS;
        sizes[ranks[n]] = HUF_MAX_BITS;
        n++;
    }
    std::vector<int> q[6];
    size_t head[6] = {0, 0, 0, 0, 0, 0};
    while (n < count) {
        int idx = HUF_MAX_BITS - 1 - sizes[ranks[n]];
        if (idx > 5 || debt < (1 << idx)) break;
        q[idx].push_back(ranks[n]);
        n++;
    }
    int idx = 5;
    while (debt > 0 && idx >= 0) {
        if (head[idx] == q[idx].size() || debt < (1 << idx)) {
            idx--;
            continue;
        }
        sizes[q[idx][head[idx]++]]++;
        debt -= (1 << idx);
    }
    idx = 0;
    while (debt > 0 && idx < 6) {
        if (head[idx] == q[idx].size()) {
            idx++;
            continue;
        }
        sizes[q[idx][head[idx]++]]++;
        debt -= (1 << idx);
    }
    if (debt > 0) {  // :272-294 slow path
        int f[256], alpha[256];
        int total = 0;
        for (int i = 0; i < count; i++) {
            f[i] = freqs[symbols[i]];
            total += f[i];
        }
        normalize_frequencies(f, alpha, count, total, HUF_CHUNK >> 3);
        for (int i = 0; i < count; i++) {
            freqs[symbols[i]] = f[i];
            ranks[i] = (f[i] << 8) | symbols[i];
        }
        return compute_code_lengths(sizes, ranks, count);
    }
    return HUF_MAX_BITS;
}

struct HuffEncoder {
    BitWriter& bw;
    uint16_t codes[256];
    std::vector<uint8_t> buffer;
    explicit HuffEncoder(BitWriter& w) : bw(w) {
        for (int i = 0; i < 256; i++) codes[i] = uint16_t(i);
    }

    // :128-214
    int update_frequencies(int* freqs) {
        int count = 0;
        uint8_t sizes[256];
        int alphabet[256];
        std::memset(sizes, 0, sizeof(sizes));
        for (int i = 0; i < 256; i++) {
            codes[i] = 0;
            if (freqs[i] > 0) alphabet[count++] = i;
        }
        encode_alphabet(bw, alphabet, count);
        if (count == 0) return 0;
        if (count == 1) {
            codes[alphabet[0]] = 1 << 12;
            sizes[alphabet[0]] = 1;
        } else {
            int ranks[256];
            for (int i = 0; i < count; i++) ranks[i] = (freqs[alphabet[i]] << 8) | alphabet[i];
            int max_len = compute_code_lengths(sizes, ranks, count);
            if (max_len > HUF_MAX_BITS) max_len = limit_code_lengths(alphabet, freqs, sizes, ranks, count);
            if (max_len > HUF_MAX_BITS) {
                for (int i = 0; i < count; i++) {
                    codes[alphabet[i]] = uint16_t(i);
                    sizes[alphabet[i]] = 8;
                }
            } else {
                generate_canonical_codes(sizes, codes, ranks, count, HUF_MAX_BITS);
            }
        }
        uint8_t prev = 2;
        for (int i = 0; i < count; i++) {
            int s = alphabet[i];
            uint8_t cur = sizes[s];
            codes[s] |= uint16_t(cur) << 12;
            expgolomb_signed_encode(bw, uint8_t(cur - prev));
            prev = cur;
        }
        return count;
    }

    // :435-511. Each of the 4 fragments is the plain MSB-first concatenation of its symbols' codes.
    void encode_chunk(const uint8_t* block, int count) {
        uint32_t nb_bits[4];
        int sz_frag = count / 4;
        std::vector<BitWriter> frag(4);
        for (int j = 0; j < 4; j++) {
            const uint8_t* src = block + j * sz_frag;
            BitWriter& f = frag[j];
            f.reserve(size_t(sz_frag) * 2 + 16);
            for (int i = 0; i < sz_frag; i++) {
                uint16_t c = codes[src[i]];
                f.write_bits(c & 0x0FFF, c >> 12);
            }
            nb_bits[j] = uint32_t(f.written());
            f.close();
        }
        for (int j = 0; j < 4; j++) write_varint(bw, nb_bits[j]);
        for (int j = 0; j < 4; j++) bw.write_array(frag[j].buf.data(), nb_bits[j]);
        for (int i = 4 * sz_frag; i < count; i++) bw.write_bits(block[i], 8);
    }

    // :390-433
    void write(const uint8_t* block, size_t len) {
        size_t start = 0;
        while (start < len) {
            int sz = int(std::min<size_t>(HUF_CHUNK, len - start));
            if (sz < 32) {
                bw.write_array(block + start, 8 * uint64_t(sz));
            } else {
                int freqs[256];
                std::memset(freqs, 0, sizeof(freqs));
                histogram_order0(block + start, size_t(sz), freqs);
                int count = update_frequencies(freqs);
                if (count > 1) encode_chunk(block + start, sz);
            }
            start += size_t(sz);
        }
    }
};

struct HuffDecoder {
    BitReader& br;
    uint16_t codes[256];
    int alphabet[256];
    uint8_t sizes[256];
    uint16_t table[1 << HUF_MAX_BITS];
    std::vector<uint8_t> buffer;
    explicit HuffDecoder(BitReader& r) : br(r) {
        for (int i = 0; i < 256; i++) {
            sizes[i] = 8;
            codes[i] = uint16_t(i);
        }
    }

    // :620-658
    int read_lengths() {
        int count = decode_alphabet(br, alphabet);
        if (count == 0) return 0;
        int8_t cur = 2;
        for (int i = 0; i < count; i++) {
            int s = alphabet[i];
            codes[s] = 0;
            cur = int8_t(cur + int8_t(expgolomb_signed_decode(br)));
            if (cur <= 0 || cur > HUF_MAX_BITS) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect size for Huffman symbol");
            sizes[s] = uint8_t(cur);
        }
        generate_canonical_codes(sizes, codes, alphabet, count, HUF_MAX_BITS);
        return count;
    }

    // :661-697
    bool build_table(int count) {
        for (auto& t : table) t = 7;
        int length = 0;
        for (int i = 0; i < count; i++) {
            int s = alphabet[i];
            if (sizes[s] > length) length = sizes[s];
            uint32_t idx = uint32_t(uint16_t(codes[s] << (HUF_MAX_BITS - length)));
            uint32_t end = idx + (1u << (HUF_MAX_BITS - length));
            if (end > (1u << HUF_MAX_BITS)) return false;
            uint16_t val = uint16_t((s << 8) | sizes[s]);
            for (uint32_t j = idx; j < end; j++) table[j] = val;
        }
        return true;
    }

    // :807-969. For every fragment the reference keeps a 64-bit window refilled to >= 56 bits and looks up the
    // top 12 bits; with 8 zero bytes appended past each fragment (:832-846) that is exactly "decode sz_frag
    // symbols sequentially from a zero-extended bit string", which is what is done here.
    void decode_chunk(uint8_t* block, int count) {
        uint32_t sz_bits[4];
        for (int j = 0; j < 4; j++) {
            sz_bits[j] = read_varint(br);
            if (int32_t(sz_bits[j]) < 0) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect stream size");
        }
        int sz_frag = count / 4;
        size_t stride = size_t(2 * HUF_CHUNK) / 4;  // :758-760, :818
        if (buffer.size() < 4 * stride + 16) buffer.assign(4 * stride + 16, 0);
        for (int j = 0; j < 4; j++) {
            size_t nbytes = (size_t(sz_bits[j]) + 7) >> 3;
            if (nbytes > buffer.size() - j * stride) throw Error(ERR_PROCESS_BLOCK, "slice bounds out of range");
            std::memset(&buffer[j * stride], 0, std::min(stride, nbytes + 16));
            br.read_array(&buffer[j * stride], sz_bits[j]);
        }
        for (int j = 0; j < 4; j++) {
            const uint8_t* p = &buffer[j * stride];
            uint8_t* out = block + size_t(j) * sz_frag;
            uint64_t bitpos = 0;
            uint64_t limit = (uint64_t(stride) << 3);
            for (int i = 0; i < sz_frag; i++) {
                // 12 bits at bitpos, zero extended
                uint32_t w = 0;
                if (bitpos + 24 <= limit) {
                    size_t b = size_t(bitpos >> 3);
                    w = (uint32_t(p[b]) << 16) | (uint32_t(p[b + 1]) << 8) | uint32_t(p[b + 2]);
                    w = (w >> (12 - (bitpos & 7))) & 0xFFF;
                }
                uint16_t v = table[w];
                out[i] = uint8_t(v >> 8);
                bitpos += (v & 0xFF);
            }
        }
        for (int i = 4 * sz_frag; i < count; i++) block[i] = uint8_t(br.read_bits(8));
    }

    // :757-805
    void read(uint8_t* block, size_t len) {
        size_t start = 0;
        while (start < len) {
            int sz = int(std::min<size_t>(HUF_CHUNK, len - start));
            if (sz < 32) {
                br.read_array(block + start, 8 * uint64_t(sz));
            } else {
                int asz = read_lengths();
                if (asz == 0) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: empty alphabet");  // reference: silent short read
                if (asz == 1) {
                    std::memset(block + start, alphabet[0], size_t(sz));
                } else {
                    if (!build_table(asz)) throw Error(ERR_PROCESS_BLOCK, "Invalid bitstream: incorrect symbol size");
                    decode_chunk(block + start, sz);
                }
            }
            start += size_t(sz);
        }
    }
};
}  // namespace

void huffman_encode(BitWriter& bw, const uint8_t* block, size_t n) {
    if (n == 0) return;
    HuffEncoder e(bw);
    e.write(block, n);
}
void huffman_decode(BitReader& br, uint8_t* block, size_t n) {
    if (n == 0) return;
    HuffDecoder d(br);
    d.read(block, n);
}

}  // namespace kzo
