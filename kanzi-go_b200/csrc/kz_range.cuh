// Declarations of the range coder kernels (kz_range.cu).
#pragma once
#include "kz_ans.cuh"

namespace kz {

static const int RC_ENC_WARPS = 4;
static const int RC_HDR_WORDS = 120;                   // 2+6+256 alphabet bits + 3 + 32 groups x (4 + 8 x 12) bits = 3467 bits
static const int RC_OUT_STRIDE = 3 * RANGE_CHUNK + 1024;  // per chunk scratch: header + 3 bytes per symbol

__global__ void range_encode_kernel(const uint8_t* in, const ChunkIn* chunks, int nchunks, uint8_t* outbuf, uint64_t* seg_src, uint64_t* seg_bits,
                                    const uint32_t* seg_index, int32_t* status);
__global__ void range_decode_kernel(const uint32_t* words, const DecBlock* blocks, const DecChunk* chunks, int nblocks, uint8_t* out,
                                    uint64_t* blk_end, int32_t* status);

}  // namespace kz
