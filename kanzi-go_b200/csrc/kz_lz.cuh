// LZ / LZX kernels (kz_lz.cu)
#pragma once
#include <vector>

#include "kz_common.cuh"

namespace kz {

struct LzBlock {
    uint64_t src_off;      // input offset (forward: block data, inverse: LZ stream), 16-byte aligned
    uint64_t dst_off;      // output offset, 16-byte aligned
    uint64_t scratch_off;  // forward only
    uint32_t len;          // input length
    uint32_t cap;          // inverse only: size of the destination (len(dst) of the reference call)
    int32_t data_type;     // ctx["dataType"] (internal/Global.go DataType)
    uint32_t pad;
};
struct LzRun {
    uint32_t src, dst, len;
};
struct LzResult {
    int32_t status;  // 0 = transformed, 1 = skip (the reference returns an error: block too small / no compression / small alphabet)
    uint32_t out_len, n_runs, lit_end, tk_n, m_n, mlen_n;
};

static inline size_t lz_scratch_bytes(size_t count) {
    size_t s = (count / 4 + 64) + (3 * (count / 4) + 64) + (count + 64);
    s = (s + 15) & ~size_t(15);
    return s + (count / 4 + 8) * sizeof(LzRun) + 256;
}

// ---- data-parallel forward path (kz_lz_par.cu) ----
struct LzWorkspace {  // device scratch provider: returns a base pointer with at least `bytes` bytes (contents need not survive a call)
    virtual uint8_t* ensure(size_t bytes) = 0;
    virtual ~LzWorkspace() {}
};
struct LzHook {  // optional per pass instrumentation (kz_api.cu times the passes with CUDA events)
    virtual void begin(const char* name) = 0;
    virtual void end() = 0;
    virtual ~LzHook() {}
};
struct LzHookScope {
    LzHook* h;
    LzHookScope(LzHook* hook, const char* name) : h(hook) {
        if (h) h->begin(name);
    }
    ~LzHookScope() {
        if (h) h->end();
    }
};
// LZ / LZX forward of a batch: block b = lb[b] (len == 0: skipped). res[b].status 0 = transformed (out_len bytes at dst_off), 1 = the transform
// declines (block too small, small alphabet, no compression). Synchronises the stream before it returns.
cudaError_t lz_forward_parallel(const uint8_t* d_in, uint8_t* d_out, const std::vector<LzBlock>& lb, bool extra, LzWorkspace& W, cudaStream_t stream,
                                std::vector<LzResult>& res, LzHook* hook);
cudaError_t lz_forward_parallel_sub(const uint8_t* d_in, uint8_t* d_out, const LzBlock* lb, uint32_t nblocks, bool extra, LzWorkspace& W, cudaStream_t stream,
                                    LzResult* res, LzHook* hook);

// LZ / LZX inverse of a batch (kz_lz_inv.cu): parse by prefix scans, pointer jumping instead of ordered copies; blocks that trip a check of the
// reference go through lz_inverse_kernel. out_len[b] / status[b] as lz_inverse_kernel leaves them. Synchronises the stream.
cudaError_t lz_inverse_parallel(const uint8_t* d_in, uint8_t* d_out, const std::vector<LzBlock>& lb, LzWorkspace& W, cudaStream_t stream, std::vector<uint32_t>& out_len,
                                std::vector<int32_t>& status, LzHook* hook, uint64_t* launches);

__global__ void lz_parse_kernel(const uint8_t* in, const LzBlock* blocks, int nblocks, int extra, int32_t* hashes_all, uint8_t* scratch_all, uint8_t* out,
                                LzResult* res);
__global__ void lz_gather_kernel(const uint8_t* in, const LzBlock* blocks, int nblocks, const uint8_t* scratch_all, const LzResult* res, uint8_t* out);
__global__ void lz_inverse_kernel(const uint8_t* in, const LzBlock* blocks, int nblocks, uint8_t* out, uint32_t* out_len, int32_t* status);

}  // namespace kz
