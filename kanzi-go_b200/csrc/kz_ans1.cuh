// Order-1 rANS kernels (kz_ans1.cu)
#pragma once
#include "kz_ans.cuh"

namespace kz {

static const int ANS1_CHUNK = ANS0_CHUNK << 8;          // 4 MiB (entropy/ANSRangeCodec.go:98-100)
static const int ANS1_HIST_WORDS = 256 * 257;
static const int ANS1_CTX_HDR_STRIDE = 416;              // <= 262 alphabet bits + 32 * 4 + 255 * 11 frequency bits
static const int ANS1_SEGS = 259;                        // lr | 256 context tables | size + states | payload

__global__ void ans1_hist_kernel(const uint8_t* in, const ChunkIn* chunks, int nchunks, uint32_t* hist);
__global__ void ans1_stats_kernel(const uint32_t* hist, int nchunks, uint32_t* enc_tab, uint8_t* ctx_hdr, uint32_t* ctx_bits);
__global__ void ans1_encode_kernel(const uint8_t* in, const ChunkIn* chunks, int nchunks, const uint32_t* enc_tab, const uint32_t* rcp, uint8_t* ctx_hdr,
                                   const uint32_t* ctx_bits, uint8_t* tail_hdr, uint8_t* pay, uint64_t pay_stride, uint64_t* seg_src, uint64_t* seg_bits,
                                   const uint32_t* seg_index, int32_t* status);
__global__ void ans1_decode_kernel(const uint32_t* words, uint64_t words_len, const DecChunk* chunks, int nchunks, const uint64_t* chunk_hdr_bit,
                                   uint8_t* f2s_all, uint32_t* sym_all, uint8_t* out, int32_t* status);

}  // namespace kz
