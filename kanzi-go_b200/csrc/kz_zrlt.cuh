// Declarations of the zero run length kernels (kz_zrlt.cu).
#pragma once
#include "kz_common.cuh"

namespace kz {

struct ZrltBlock {
    uint64_t src_off;  // byte offsets of the block inside the source / destination buffers (16-byte aligned)
    uint64_t dst_off;
    uint32_t len;
    uint32_t cap;      // inverse: capacity of the (zero filled) destination block
    uint32_t active;
    uint32_t pad;
};

__global__ void zrlt_forward_kernel(const uint8_t* data, const ZrltBlock* blocks, int nblocks, uint8_t* out, uint32_t* out_len, int32_t* status);
__global__ void zrlt_inverse_kernel(const uint8_t* data, const ZrltBlock* blocks, int nblocks, uint8_t* out, uint32_t* out_len, int32_t* status);

}  // namespace kz
