// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// C++ restatement ("port", the Go toolchain is absent from this image) of the per-block encode/decode
// path of flanglet/kanzi-go @ ba60b1f (entropy codecs, BWT / BWTS / LZ / LZX / ROLZ / RANK / MTFT / ZRLT, the pre-transforms TEXT (encoding 2) /
// UTF / PACK / DNA / MM / EXE, transform sequences, block framing, stream, XXHash). Each function cites the reference file:line it follows.
// PARITY PINNING: the reference ships no byte-level golden vectors for this path and
// cannot be built here (no Go toolchain) -> "parity unpinned" except for the pins listed in
// tests/test_oracle_pins.py (varint sizes, BWT "mississippi" doc example, format constants).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "kzo_bitio.hpp"

namespace kzo {

// ---- ids: v2/entropy/EntropyCodecFactory.go:26-35, v2/transform/Factory.go:31-50
enum EntropyType : uint32_t { E_NONE = 0, E_HUFFMAN = 1, E_FPAQ = 2, E_PAQ = 3, E_RANGE = 4, E_ANS0 = 5, E_CM = 6, E_TPAQ = 7, E_ANS1 = 8, E_TPAQX = 9 };
enum TransformType : uint64_t {
    T_NONE = 0, T_BWT = 1, T_BWTS = 2, T_LZ = 3, T_SNAPPY = 4, T_RLT = 5, T_ZRLT = 6, T_MTFT = 7, T_RANK = 8, T_EXE = 9,
    T_DICT = 10, T_ROLZ = 11, T_ROLZX = 12, T_SRT = 13, T_LZP = 14, T_MM = 15, T_LZX = 16, T_UTF = 17, T_PACK = 18, T_DNA = 19
};

// data types, v2/internal/Global.go (DataType)
enum DataType { DT_UNDEFINED = 0, DT_TEXT, DT_MULTIMEDIA, DT_EXE, DT_NUMERIC, DT_BASE64, DT_DNA, DT_BIN, DT_UTF8, DT_SMALL_ALPHABET };

// ---- entropy utils (v2/entropy/EntropyUtils.go)
int encode_alphabet(BitWriter& bw, const int* alphabet, int count);
int decode_alphabet(BitReader& br, int* alphabet /*[256]*/);
int normalize_frequencies(int* freqs /*[256]*/, int* alphabet /*[256]*/, int alphabet_len, int total, int scale);
int write_varint(BitWriter& bw, uint32_t v);
uint32_t read_varint(BitReader& br);
void histogram_order0(const uint8_t* p, size_t n, int* freqs /*[256]*/);
// order-1 with totals: freqs[257*ctx + sym], freqs[257*ctx + 256] = total of ctx; ctx of first byte is 0
void histogram_order1(const uint8_t* p, size_t n, int* freqs /*[256*257]*/);
uint32_t log2_floor(uint32_t x);
int first_order_entropy_1024(size_t len, const int* histo);

// ---- entropy codecs: append to / read from a bit stream exactly like EntropyEncoder.Write / EntropyDecoder.Read
void ans_encode(BitWriter& bw, const uint8_t* block, size_t n, int order, int chunk_size = 0, int log_range = 12);
void ans_decode(BitReader& br, uint8_t* block, size_t n, int order, int chunk_size = 0);
void huffman_encode(BitWriter& bw, const uint8_t* block, size_t n);
void huffman_decode(BitReader& br, uint8_t* block, size_t n);
void range_encode(BitWriter& bw, const uint8_t* block, size_t n);
void range_decode(BitReader& br, uint8_t* block, size_t n);
void null_encode(BitWriter& bw, const uint8_t* block, size_t n);
void null_decode(BitReader& br, uint8_t* block, size_t n);
void entropy_encode(uint32_t type, BitWriter& bw, const uint8_t* block, size_t n);
void entropy_decode(uint32_t type, BitReader& br, uint8_t* block, size_t n);

// ---- transforms. forward returns false when the reference would return an error ("skip me").
struct Ctx {
    int data_type = DT_UNDEFINED;
    uint32_t entropy_type = E_NONE;
    int jobs = 1;
    size_t block_size = 0;
    size_t stream_block_size = 0;  // ctx["blockSize"] of the reference (CompressedStream.go:220,1406): TEXT sizes its hash map from it
};
size_t transform_max_encoded_len(uint64_t type, size_t n);
bool transform_forward(uint64_t type, Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
bool transform_inverse(uint64_t type, Ctx& ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);

// BWT pieces exposed for tests
void bwt_forward_raw(const uint8_t* src, uint8_t* dst, size_t n, uint32_t* primary /*[8]*/);
bool bwt_inverse_raw(const uint8_t* src, uint8_t* dst, size_t n, const uint32_t* primary /*[8]*/);
int bwt_chunks(size_t n);

// ---- block framing (v2/io/CompressedStream.go encodingTask.encode / decodingTask.decode)
struct BlockResult {
    std::vector<uint8_t> bytes;  // block-local bitstring, zero padded
    uint64_t bits = 0;           // exact bit length ("written")
};
// transform48: 8 x 6-bit ids, first transform in the top 6 bits. checksum_bits: 0/32/64.
void encode_block(const uint8_t* data, size_t n, uint64_t transform48, uint32_t entropy, int checksum_bits, bool skip_blocks, BlockResult& out,
                  size_t stream_block_size = 0 /* 0: the block's own length */);
// returns decoded length. block_size = stream block size (sanity bound :1896)
size_t decode_block(const uint8_t* payload, uint64_t bits, uint64_t transform48, uint32_t entropy, int checksum_bits, size_t block_size,
                    std::vector<uint8_t>& out);

// ---- whole stream (Writer / Reader), in memory, `jobs` worker threads over blocks
struct StreamParams {
    uint64_t transform48 = 0;
    uint32_t entropy = E_NONE;
    uint32_t block_size = 4 << 20;
    int checksum_bits = 0;
    int jobs = 1;
    bool headerless = false;
    bool skip_blocks = false;
    int64_t input_size = 0;  // "fileSize" ctx; 0 = not provided
};
std::vector<uint8_t> compress_stream(const uint8_t* src, size_t n, const StreamParams& p);
std::vector<uint8_t> decompress_stream(const uint8_t* src, size_t n, int jobs, StreamParams* header_out = nullptr,
                                       const StreamParams* headerless_params = nullptr);
uint64_t parse_transform_names(const std::string& names);  // "BWT+RANK+ZRLT" -> 48-bit id word
uint32_t parse_entropy_name(const std::string& name);

uint32_t xxhash32(const uint8_t* p, size_t n, uint32_t seed);
uint64_t xxhash64(const uint8_t* p, size_t n, uint64_t seed);

}  // namespace kzo
