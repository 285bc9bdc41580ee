// ORACLE — TEST INFRASTRUCTURE ONLY (see kzo.hpp).
// Block framing and whole-stream Writer/Reader restatement. Follows v2/io/CompressedStream.go:
//   writeHeader :429-519, processBlock :621-710, encodingTask.encode :729-977, Close :576-619,
//   readHeader :1316-1522, decodingTask.decode :1763-2012; transform sequence v2/transform/Sequence.go,
//   v2/transform/Factory.go:58-95; magic sniffing v2/internal/Magic.go:73-112.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <sstream>
#include <thread>

#include "kzo.hpp"

namespace kzo {

// ---- XXHash (v2/hash/XXHash32.go:51-102, XXHash64.go) — the standard XXH32/XXH64
static inline uint32_t rd32le(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
static inline uint64_t rd64le(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

uint32_t xxhash32(const uint8_t* p, size_t len, uint32_t seed) {
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    size_t n = 0;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        for (; n + 16 <= len; n += 16) {
            v1 = rotl32(v1 + rd32le(p + n) * P2, 13) * P1;
            v2 = rotl32(v2 + rd32le(p + n + 4) * P2, 13) * P1;
            v3 = rotl32(v3 + rd32le(p + n + 8) * P2, 13) * P1;
            v4 = rotl32(v4 + rd32le(p + n + 12) * P2, 13) * P1;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + P5;
    }
    h += uint32_t(len);
    for (; n + 4 <= len; n += 4) h = rotl32(h + rd32le(p + n) * P3, 17) * P4;
    for (; n < len; n++) h = rotl32(h + uint32_t(p[n]) * P5, 11) * P1;
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

uint64_t xxhash64(const uint8_t* p, size_t len, uint64_t seed) {
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    auto round = [&](uint64_t acc, uint64_t v) { return rotl64(acc + v * P2, 31) * P1; };
    auto merge = [&](uint64_t h, uint64_t v) { return (h ^ round(0, v)) * P1 + P4; };
    size_t n = 0;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        for (; n + 32 <= len; n += 32) {
            v1 = round(v1, rd64le(p + n));
            v2 = round(v2, rd64le(p + n + 8));
            v3 = round(v3, rd64le(p + n + 16));
            v4 = round(v4, rd64le(p + n + 24));
        }
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += uint64_t(len);
    for (; n + 8 <= len; n += 8) h = rotl64(h ^ round(0, rd64le(p + n)), 27) * P1 + P4;
    for (; n + 4 <= len; n += 4) h = rotl64(h ^ (uint64_t(rd32le(p + n)) * P1), 23) * P2 + P3;
    for (; n < len; n++) h = rotl64(h ^ (uint64_t(p[n]) * P5), 11) * P1;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// ---- magic (v2/internal/Magic.go:22-112, :116-206)
namespace {
const uint32_t BS_TYPE = 0x4B414E5A;  // "KANZ" CompressedStream.go:43
const int BS_VERSION = 6;             // :44
const uint8_t COPY_MASK = 0x80, TRANSFORMS_MASK = 0x10;  // :47-48
const size_t SMALL_BLOCK = 15;        // :51
const size_t MIN_BLOCK = 1024, MAX_BLOCK = size_t(1) << 30;  // :49-50
const int MAX_CONCURRENCY = 64;       // :52

uint32_t magic_type(const uint8_t* src, size_t n) {
    if (n < 4) return 0;
    uint32_t key = (uint32_t(src[0]) << 24) | (uint32_t(src[1]) << 16) | (uint32_t(src[2]) << 8) | src[3];
    if ((key & ~0x0Fu) == 0xFFD8FFE0u) return key;
    if ((key >> 8) == 0x425A68u || (key >> 8) == 0x494433u) return key >> 8;
    static const uint32_t k32[] = {0x47494638u, 0x25504446u, 0x504B0304u, 0x377ABCAFu, 0x89504E47u, 0x7F454C46u, 0xFEEDFACEu, 0xCEFAEDFEu,
                                   0xFEEDFACFu, 0xCFFAEDFEu, 0x28B52FFDu, 0x81CFB2CEu, 0x4D534346u, 0x52494646u, 0x664C6143u, 0xFD377A58u,
                                   0x4B414E5Au, 0x52617221u};
    for (uint32_t k : k32) if (key == k) return key;
    uint32_t key16 = key >> 16;
    if (key16 == 0x1F8B || key16 == 0x424D || key16 == 0x4D5A) return key16;
    if (key16 == 0x5034 || key16 == 0x5035 || key16 == 0x5036) {
        uint32_t sub = (key >> 8) & 0xFF;
        if (sub == 0x07 || sub == 0x0A || sub == 0x0D || sub == 0x20) return key16;
    }
    return 0;
}
bool magic_compressed(uint32_t m) {
    switch (m) {
        case 0xFFD8FFE0u: case 0x47494638u: case 0x89504E47u: case 0x377ABCAFu: case 0x28B52FFDu: case 0x81CFB2CEu: case 0x4D534346u:
        case 0x504B0304u: case 0x1F8Bu: case 0x425A68u: case 0x664C6143u: case 0x494433u: case 0xFD377A58u: case 0x4B414E5Au: case 0x52617221u:
            return true;
        default: return false;
    }
}
bool magic_multimedia(uint32_t m) {
    switch (m) {
        case 0xFFD8FFE0u: case 0x47494638u: case 0x89504E47u: case 0x52494646u: case 0x664C6143u: case 0x494433u: case 0x424Du:
        case 0x5034u: case 0x5035u: case 0x5036u:
            return true;
        default: return false;
    }
}
bool magic_executable(uint32_t m) {
    switch (m) {
        case 0x7F454C46u: case 0x4D5Au: case 0xFEEDFACEu: case 0xCEFAEDFEu: case 0xFEEDFACFu: case 0xCFFAEDFEu: return true;
        default: return false;
    }
}

}  // namespace
uint32_t get_magic_type(const uint8_t* src, size_t n) { return magic_type(src, n); }
namespace {
// Factory.New (:58-95): the list of transforms of a 48-bit type word. All-NONE keeps a single NONE.
std::vector<uint64_t> sequence_of(uint64_t t48) {
    std::vector<uint64_t> seq;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (t48 >> (42 - 6 * i)) & 63;
        if (t != T_NONE) seq.push_back(t);
    }
    if (seq.empty()) seq.push_back(T_NONE);
    return seq;
}
// NB: Factory.New fills slot i from the i-th 6-bit field (not the i-th non-NONE field): a word such as
// NONE+BWT would keep NONE at index 0 and leave a nil (dropped by NewByteTransformSequence). GetType()
// (Factory.go:330-395) never produces such words because NONE tokens are skipped while packing, so only the
// packed form is supported here.

size_t seq_max_encoded_len(const std::vector<uint64_t>& seq, size_t n) {  // Sequence.go:189-205
    size_t req = n;
    for (uint64_t t : seq) req = std::max(req, transform_max_encoded_len(t, req));
    return req;
}

// Sequence.Forward (:64-125). Returns post-transform length; result in `out`.
size_t seq_forward(const std::vector<uint64_t>& seq, Ctx& ctx, const uint8_t* src, size_t n, std::vector<uint8_t>& out, uint8_t& skip_flags) {
    skip_flags = 0xFF;
    if (n == 0) return 0;
    size_t required = seq_max_encoded_len(seq, n);
    std::vector<uint8_t> a(src, src + n), b;
    a.resize(std::max(required, n));
    b.assign(required, 0);
    size_t length = n;
    std::vector<uint8_t>*in = &a, *outp = &b;
    for (size_t i = 0; i < seq.size(); i++) {
        size_t new_len = 0;
        if (!transform_forward(seq[i], ctx, in->data(), length, outp->data(), outp->size(), &new_len)) continue;
        length = new_len;
        skip_flags &= uint8_t(~(1u << (7 - i)));
        std::swap(in, outp);
    }
    out.assign(in->begin(), in->begin() + length);
    return length;
}

// Sequence.Inverse (:131-186)
size_t seq_inverse(const std::vector<uint64_t>& seq, Ctx& ctx, uint8_t skip_flags, const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t cap) {
    if (n == 0) { out.clear(); return 0; }
    if (skip_flags == 0xFF) {
        out.assign(src, src + n);
        return n;
    }
    std::vector<uint8_t> a(src, src + n), b;
    a.resize(std::max(cap, n));
    b.assign(cap, 0);
    size_t length = n;
    std::vector<uint8_t>*in = &a, *outp = &b;
    for (int i = int(seq.size()) - 1; i >= 0; i--) {
        if (skip_flags & (1u << (7 - i))) continue;
        size_t new_len = 0;
        if (!transform_inverse(seq[size_t(i)], ctx, in->data(), length, outp->data(), outp->size(), &new_len))
            throw Error(ERR_PROCESS_BLOCK, "Inverse transform failed");
        length = new_len;
        std::swap(in, outp);
    }
    out.assign(in->begin(), in->begin() + length);
    return length;
}
}  // namespace

// encodingTask.encode :729-933 (everything up to the ordered commit)
void encode_block(const uint8_t* data, size_t n, uint64_t t48, uint32_t etype, int checksum_bits, bool skip_blocks, BlockResult& res,
                  size_t stream_block_size) {
    uint8_t mode = 0;
    uint64_t checksum = 0;
    if (checksum_bits == 32) checksum = xxhash32(data, n, BS_TYPE);
    else if (checksum_bits == 64) checksum = xxhash64(data, n, BS_TYPE);

    if (n <= SMALL_BLOCK) {
        t48 = 0;
        etype = E_NONE;
        mode |= COPY_MASK;
    } else if (skip_blocks) {
        bool skip = false;
        if (n >= 8) skip = magic_compressed(magic_type(data, n));
        if (!skip) {
            int histo[256];
            std::memset(histo, 0, sizeof(histo));
            histogram_order0(data, n, histo);
            skip = first_order_entropy_1024(n, histo) >= 973;  // INCOMPRESSIBLE_THRESHOLD EntropyUtils.go:26
        }
        if (skip) {
            t48 = 0;
            etype = E_NONE;
            mode |= COPY_MASK;
        }
    }
    Ctx ctx;
    ctx.entropy_type = etype;
    ctx.block_size = n;
    ctx.stream_block_size = stream_block_size ? stream_block_size : n;
    std::vector<uint64_t> seq = sequence_of(t48);
    uint32_t magic = magic_type(data, n);
    if (magic_compressed(magic)) ctx.data_type = DT_BIN;
    else if (magic_multimedia(magic)) ctx.data_type = DT_MULTIMEDIA;
    else if (magic_executable(magic)) ctx.data_type = DT_EXE;

    std::vector<uint8_t> buffer;
    uint8_t skip_flags;
    size_t post = seq_forward(seq, ctx, data, n, buffer, skip_flags);
    unsigned data_size = 1;
    if (post >= 256) {
        data_size = (log2_floor(uint32_t(post)) >> 3) + 1;
        if (data_size > 4) throw Error(ERR_WRITE_FILE, "Invalid block data length");
    }
    mode |= uint8_t(((data_size - 1) & 3) << 5);
    BitWriter bw;
    bw.reserve(std::max(post, n + (n >> 3)) + 64);
    if ((mode & COPY_MASK) != 0 || seq.size() <= 4) {
        mode |= uint8_t(skip_flags >> 4);
        bw.write_bits(mode, 8);
    } else {
        mode |= TRANSFORMS_MASK;
        bw.write_bits(mode, 8);
        bw.write_bits(skip_flags, 8);
    }
    bw.write_bits(uint64_t(post), 8 * data_size);
    if (checksum_bits == 32) bw.write_bits(checksum, 32);
    else if (checksum_bits == 64) bw.write_bits(checksum, 64);
    entropy_encode(etype, bw, buffer.data(), post);
    res.bits = bw.written();
    res.bytes = std::move(bw.close());
}

// decodingTask.decode :1873-2011 (after the serial extraction of the block's bits)
size_t decode_block(const uint8_t* payload, uint64_t bits, uint64_t t48, uint32_t etype, int checksum_bits, size_t block_size,
                    std::vector<uint8_t>& out) {
    size_t r = size_t((bits + 7) >> 3);
    BitReader br = BitReader::over_bytes(payload, r);
    uint8_t mode = uint8_t(br.read_bits(8));
    uint8_t skip_flags = 0;
    if (mode & COPY_MASK) {
        t48 = 0;
        etype = E_NONE;
    } else if (mode & TRANSFORMS_MASK) {
        skip_flags = uint8_t(br.read_bits(8));
    } else {
        skip_flags = uint8_t((mode << 4) | 0x0F);
    }
    unsigned data_size = 1 + ((mode >> 5) & 3);
    size_t pre = size_t(br.read_bits(8 * data_size));
    size_t max_len = std::min(std::max(block_size + block_size / 2, size_t(2048)), MAX_BLOCK);
    if (pre == 0 || pre > max_len) throw Error(ERR_BLOCK_SIZE, "Invalid compressed block size");
    uint64_t checksum1 = 0;
    if (checksum_bits == 32) checksum1 = br.read_bits(32);
    else if (checksum_bits == 64) checksum1 = br.read_bits(64);
    std::vector<uint8_t> buffer(pre + 512);
    entropy_decode(etype, br, buffer.data(), pre);
    Ctx ctx;
    ctx.entropy_type = etype;
    ctx.block_size = pre;
    ctx.stream_block_size = block_size;
    std::vector<uint64_t> seq = sequence_of(t48);
    // Reader.processBlock :1649-1653: the task buffers hold blockSize + max(EXTRA_BUFFER_SIZE = 512, blockSize >> 4) bytes
    size_t cap = std::max(r, block_size + std::max<size_t>(512, block_size >> 4));
    size_t decoded = seq_inverse(seq, ctx, skip_flags, buffer.data(), pre, out, cap);
    if (checksum_bits == 32) {
        if (xxhash32(out.data(), decoded, BS_TYPE) != uint32_t(checksum1)) throw Error(ERR_CRC_CHECK, "Corrupted bitstream: checksum mismatch");
    } else if (checksum_bits == 64) {
        if (xxhash64(out.data(), decoded, BS_TYPE) != checksum1) throw Error(ERR_CRC_CHECK, "Corrupted bitstream: checksum mismatch");
    }
    return decoded;
}

// ---- stream
namespace {
uint32_t header_checksum(int ck_size, uint32_t etype, uint64_t t48, uint32_t block_size, unsigned sz_mask, int64_t input_size) {  // :498-512
    const uint32_t HASH = 0x1E35A7BDu;
    uint32_t seed = uint32_t(0x01030507u * uint32_t(BS_VERSION));
    uint32_t ck = HASH * seed;
    ck ^= HASH * uint32_t(~ck_size);
    ck ^= HASH * uint32_t(~etype);
    ck ^= HASH * uint32_t((~t48) >> 32);
    ck ^= HASH * uint32_t(~t48);
    ck ^= HASH * uint32_t(~block_size);
    if (sz_mask > 0) {
        ck ^= HASH * uint32_t(uint64_t(~input_size) >> 32);
        ck ^= HASH * uint32_t(~input_size);
    }
    return ((ck >> 23) ^ (ck >> 3)) & 0xFFFFFFu;
}

template <class F>
void parallel_for(size_t count, int jobs, F f) {
    jobs = std::max(1, std::min<int>(jobs, int(count)));
    if (jobs == 1) {
        for (size_t i = 0; i < count; i++) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    std::vector<std::string> errs{size_t(jobs), std::string()};
    std::vector<int> codes(static_cast<size_t>(jobs), 0);
    for (int t = 0; t < jobs; t++)
        th.emplace_back([&, t] {
            try {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= count) break;
                    f(i);
                }
            } catch (const Error& e) {
                errs[size_t(t)] = e.what();
                codes[size_t(t)] = e.code;
            } catch (const std::exception& e) {
                errs[size_t(t)] = e.what();
                codes[size_t(t)] = ERR_UNKNOWN;
            }
        });
    for (auto& t : th) t.join();
    for (int t = 0; t < jobs; t++)
        if (codes[size_t(t)]) throw Error(codes[size_t(t)], errs[size_t(t)]);
}
}  // namespace

std::vector<uint8_t> compress_stream(const uint8_t* src, size_t n, const StreamParams& p) {
    if (p.block_size > MAX_BLOCK || p.block_size < MIN_BLOCK || (p.block_size & 15)) throw Error(ERR_INVALID_PARAM, "Invalid block size");
    if (p.jobs < 1 || p.jobs > MAX_CONCURRENCY) throw Error(ERR_INVALID_PARAM, "The number of jobs must be in [1..64]");
    BitWriter obs;
    obs.reserve(n / 2 + 1024);
    if (!p.headerless) {  // writeHeader :429-519
        int ck_size = p.checksum_bits == 32 ? 1 : (p.checksum_bits == 64 ? 2 : 0);
        obs.write_bits(BS_TYPE, 32);
        obs.write_bits(BS_VERSION, 4);
        obs.write_bits(uint64_t(ck_size), 2);
        obs.write_bits(p.entropy, 5);
        obs.write_bits(p.transform48, 48);
        obs.write_bits(p.block_size >> 4, 28);
        unsigned sz_mask;
        if (p.input_size == 0 || p.input_size >= (int64_t(1) << 48)) sz_mask = 0;
        else if (p.input_size >= (int64_t(1) << 32)) sz_mask = 3;
        else if (p.input_size >= (int64_t(1) << 16)) sz_mask = 2;
        else sz_mask = 1;
        obs.write_bits(sz_mask, 2);
        if (sz_mask > 0) obs.write_bits(uint64_t(p.input_size), 16 * sz_mask);
        obs.write_bits(0, 15);
        obs.write_bits(header_checksum(ck_size, p.entropy, p.transform48, p.block_size, sz_mask, p.input_size), 24);
    }
    size_t nblocks = (n + p.block_size - 1) / p.block_size;
    std::vector<BlockResult> results(nblocks);
    // processBlock :621-710 — batches of `jobs` blocks, one worker per block; output committed in block order
    parallel_for(nblocks, p.jobs, [&](size_t b) {
        size_t off = b * size_t(p.block_size);
        size_t len = std::min<size_t>(p.block_size, n - off);
        encode_block(src + off, len, p.transform48, p.entropy, p.checksum_bits, p.skip_blocks, results[b], p.block_size);
    });
    for (size_t b = 0; b < nblocks; b++) {  // :951-976
        uint64_t written = results[b].bits;
        unsigned lw = 3;
        if (written >= 8) lw = log2_floor(uint32_t(written >> 3)) + 4;
        obs.write_bits(lw - 3, 5);
        obs.write_bits(written, lw);
        obs.write_array(results[b].bytes.data(), written);
        results[b].bytes = std::vector<uint8_t>();
    }
    obs.write_bits(0, 5);  // Close :593-594
    obs.write_bits(0, 3);
    return std::move(obs.close());
}

std::vector<uint8_t> decompress_stream(const uint8_t* src, size_t n, int jobs, StreamParams* header_out, const StreamParams* headerless) {
    BitReader ibs = BitReader::over_bytes(src, n);
    StreamParams p;
    if (headerless) {
        p = *headerless;
    } else {  // readHeader :1316-1460
        if (ibs.read_bits(32) != BS_TYPE) throw Error(ERR_INVALID_FILE, "Invalid stream type");
        unsigned ver = unsigned(ibs.read_bits(4));
        if (ver != unsigned(BS_VERSION)) throw Error(ERR_STREAM_VERSION, "Only bitstream version 6 is restated");
        unsigned ck_size = unsigned(ibs.read_bits(2));
        if (ck_size == 3) throw Error(ERR_INVALID_CODEC, "Invalid bitstream, incorrect checksum size");
        p.checksum_bits = int(ck_size) * 32;
        p.entropy = uint32_t(ibs.read_bits(5));
        p.transform48 = ibs.read_bits(48);
        p.block_size = uint32_t(ibs.read_bits(28)) << 4;
        if (p.block_size < MIN_BLOCK || p.block_size > MAX_BLOCK) throw Error(ERR_BLOCK_SIZE, "Invalid bitstream, incorrect block size");
        unsigned sz_mask = unsigned(ibs.read_bits(2));
        if (sz_mask) p.input_size = int64_t(ibs.read_bits(16 * sz_mask));
        ibs.read_bits(15);
        uint32_t ck1 = uint32_t(ibs.read_bits(24));
        if (ck1 != header_checksum(int(ck_size), p.entropy, p.transform48, p.block_size, sz_mask, p.input_size))
            throw Error(ERR_CRC_CHECK, "Invalid bitstream: checksum mismatch");
    }
    if (header_out) *header_out = p;
    // serial extraction of block payloads (:1816-1852), then concurrent decode
    struct Blk { std::vector<uint8_t> payload; uint64_t bits; };
    std::vector<Blk> blks;
    for (;;) {
        unsigned lr = unsigned(ibs.read_bits(5)) + 3;
        uint64_t read = ibs.read_bits(lr);
        if (read == 0) break;
        if (read > (uint64_t(1) << 34)) throw Error(ERR_BLOCK_SIZE, "Invalid block size");
        Blk b;
        b.bits = read;
        b.payload.assign(size_t((read + 7) >> 3) + 16, 0);
        ibs.read_array(b.payload.data(), read);
        blks.push_back(std::move(b));
    }
    std::vector<std::vector<uint8_t>> outs(blks.size());
    parallel_for(blks.size(), jobs, [&](size_t i) {
        decode_block(blks[i].payload.data(), blks[i].bits, p.transform48, p.entropy, p.checksum_bits, p.block_size, outs[i]);
        blks[i].payload = std::vector<uint8_t>();
    });
    size_t total = 0;
    for (auto& o : outs) total += o.size();
    std::vector<uint8_t> res;
    res.reserve(total);
    for (auto& o : outs) res.insert(res.end(), o.begin(), o.end());
    return res;
}

// transform.GetType (Factory.go:330-395) / entropy.GetType (EntropyCodecFactory.go:170-206)
uint64_t parse_transform_names(const std::string& names) {
    static const struct { const char* name; uint64_t id; } tab[] = {
        {"NONE", T_NONE}, {"BWT", T_BWT}, {"BWTS", T_BWTS}, {"LZ", T_LZ}, {"RLT", T_RLT}, {"ZRLT", T_ZRLT}, {"MTFT", T_MTFT}, {"RANK", T_RANK},
        {"EXE", T_EXE}, {"TEXT", T_DICT}, {"ROLZ", T_ROLZ}, {"ROLZX", T_ROLZX}, {"SRT", T_SRT}, {"LZP", T_LZP}, {"MM", T_MM}, {"LZX", T_LZX},
        {"UTF", T_UTF}, {"PACK", T_PACK}, {"DNA", T_DNA}};
    uint64_t res = 0;
    int shift = 42, count = 0;
    std::stringstream ss(names);
    std::string tok;
    while (std::getline(ss, tok, '+')) {
        for (auto& c : tok) c = char(std::toupper(c));
        bool found = false;
        for (auto& e : tab)
            if (tok == e.name) {
                found = true;
                if (e.id != T_NONE) {
                    if (++count > 8) throw Error(ERR_INVALID_PARAM, "Only 8 transforms allowed");
                    res |= e.id << shift;
                    shift -= 6;
                }
            }
        if (!found) throw Error(ERR_INVALID_PARAM, "Unknown transform type: " + tok);
    }
    return res;
}
uint32_t parse_entropy_name(const std::string& name) {
    std::string s = name;
    for (auto& c : s) c = char(std::toupper(c));
    if (s == "NONE") return E_NONE;
    if (s == "HUFFMAN") return E_HUFFMAN;
    if (s == "ANS0") return E_ANS0;
    if (s == "ANS1") return E_ANS1;
    if (s == "RANGE") return E_RANGE;
    if (s == "FPAQ") return E_FPAQ;
    if (s == "CM") return E_CM;
    if (s == "TPAQ") return E_TPAQ;
    if (s == "TPAQX") return E_TPAQX;
    throw Error(ERR_INVALID_PARAM, "Unsupported entropy codec type: " + name);
}

}  // namespace kzo
