// Declarations of the block checksum kernel (kz_hash.cu).
#pragma once
#include "kz_common.cuh"

namespace kz {

struct HashJob {
    uint64_t off;  // byte offset of the block inside the data buffer
    uint32_t len;
    uint32_t pad;
};

__global__ void xxhash_blocks_kernel(const uint8_t* data, const HashJob* jobs, int njobs, int bits, uint64_t* out);

}  // namespace kz
