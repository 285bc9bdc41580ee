// Declarations of the Huffman kernels (kz_huffman.cu).
#pragma once
#include "kz_ans.cuh"

namespace kz {

static const int HUF_STATS_WARPS = 2;
static const int HUF_ENC_WARPS = 4;
static const int HUF_HDR_STRIDE = 768;                 // alphabet (<= 262 bits) + 256 Exp-Golomb deltas (<= 16 bits) + 4 varints
static const int HUF_FRAG_WORDS = 1540;                // 4096 symbols x 12 bits = 1536 words
static const int HUF_FRAG_STRIDE = HUF_FRAG_WORDS * 4; // bytes of scratch per fragment
static const int HUF_DEC_CTA_CHUNKS = 16;
static const int HUF_DEC_CHUNK_BYTES = 2944;           // 1024 x u16 table + 256 symbols + 256 sizes + 64 info + 256 scratch + pad
static const size_t HUF_DEC_SMEM = (size_t)HUF_DEC_CTA_CHUNKS * HUF_DEC_CHUNK_BYTES;

__global__ void huf_stats_kernel(const uint8_t* in, const ChunkIn* chunks, int nchunks, uint16_t* codes_out, uint8_t* hdr, uint32_t* hdr_bits,
                                 uint32_t* asz_out);
__global__ void huf_encode_kernel(const uint8_t* in, const ChunkIn* chunks, int nchunks, const uint16_t* codes_in, uint8_t* hdr,
                                  const uint32_t* hdr_bits, const uint32_t* asz_in, uint8_t* frag, uint64_t* seg_src, uint64_t* seg_bits,
                                  const uint32_t* seg_index);
__global__ void huf_walk_kernel(const uint32_t* words, uint64_t words_len, const DecBlock* blocks, int nblocks, uint64_t* chunk_hdr_bit,
                                uint64_t* blk_end, int32_t* status);
__global__ void huf_decode_kernel(const uint32_t* words, uint64_t words_len, const DecChunk* chunks, int nchunks, const uint64_t* chunk_hdr_bit,
                                  uint8_t* out, int32_t* status);

}  // namespace kz
