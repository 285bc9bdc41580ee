// Burrows-Wheeler transform of a whole block on one B200: suffix array by prefix doubling (forward) and sampled list
// ranking (inverse).
//
// Reference semantics reproduced bit-exactly (paths relative to the reference's v2/ directory):
//   transform/BWT.go:132-173 Forward + transform/DivSufSort.go:179-311 ComputeBWT: suffix order without sentinel (a suffix
//     that is a proper prefix of another sorts first), output dst[0] = src[n-1] followed by src[SA[r]-1] for every rank
//     r in order except the rank of suffix 0; primary index k = rank(k * ceil(n/8)) + 1 (1 index when n < 256).
//     The BWT is canonical, so the suffix array is built with a GPU algorithm instead of DivSufSort's induced sorting:
//     rank arrays are doubled (h = 7, 14, 28, ...) with one 64-bit radix sort of (rank[i], rank[i+h]) per round
//     (cub::DeviceRadixSort for the sort itself; everything else is hand written).
//   transform/BWT.go:178-358 Inverse (inverseMergeTPSI) / :361-628 (biPSIv2): both walk the permutation
//     "rank of suffix i -> rank of suffix i+1" from primaryIndex(0)-1. The format only offers 8 independent walks,
//     far too few for a GPU, so the permutation list is cut at every rank that is a multiple of 64 (plus the 8 primary
//     ranks): all sub-lists are walked in parallel once to measure them, the few thousand sub-list heads are ranked with
//     pointer doubling, and a second parallel walk writes the bytes at their final offsets.
//   transform/BWTBlockCodec.go:78-136, :141-225: block header (mode byte + big-endian primaryIndex-1 per chunk).
#include <cub/device/device_radix_sort.cuh>
#include <thrust/iterator/reverse_iterator.h>
#include <cub/device/dispatch/dispatch_three_way_partition.cuh>
#include <cub/device/dispatch/dispatch_segmented_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <vector>

#include "kz_bwt.cuh"

namespace kz {

namespace {

KZ_HD uint32_t bwt_chunks(uint32_t n) { return n < 256 ? 1u : 8u; }

// key of suffix i from its first 7 bytes, 9 bits per symbol (byte + 1, 0 past the end): shorter suffix sorts first
__global__ void bwt_init_keys_kernel(const uint8_t* __restrict__ src, uint32_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const uint32_t p = i + j;
        const uint64_t s = p < n ? (uint64_t)src[p] + 1 : 0;
        k = (k << 9) | s;
    }
    keys[i] = k;
    idx[i] = i;
}

// flags[j] = 1 when sorted position j starts a new group of equal keys
__global__ void bwt_flag_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ head) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    head[j] = (j == 0 || keys[j] != keys[j - 1]) ? j : 0u;
}

// inclusive max-scan over `head` in place (group start index for every sorted position), three-phase block scan
__global__ void scan_max_local_kernel(uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ block_max) {
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base = blockIdx.x * (blockDim.x * 8);
    uint32_t x[8];
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t j = base + tid * 8 + q;
        x[q] = j < n ? v[j] : 0u;
        run = max(run, x[q]);
        x[q] = run;
    }
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl = max(incl, t);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w = max(w, t);
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) prev = 0;
    if (warp > 0) prev = max(prev, s_warp[warp - 1]);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t j = base + tid * 8 + q;
        if (j < n) v[j] = max(x[q], prev);
    }
    if (tid == blockDim.x - 1) block_max[blockIdx.x] = max(incl, prev);
}
__global__ void scan_max_blocks_kernel(uint32_t* __restrict__ block_max, uint32_t nblocks) {  // one CTA, serial over tiles of 1024
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_warp[32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t start = 0; start < nblocks; start += 1024) {
        const uint32_t i = start + threadIdx.x;
        uint32_t x = i < nblocks ? block_max[i] : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl = max(incl, t);
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w = max(w, t);
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t r = max(incl, s_carry);
        if (warp > 0) r = max(r, s_warp[warp - 1]);
        if (i < nblocks) block_max[i] = r;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = r;
        __syncthreads();
    }
}
// v[j] = max(v[j], prefix of previous blocks); rank[sa[j]] = v[j]; counts positions whose group has more than one member
__global__ void bwt_apply_rank_kernel(uint32_t* __restrict__ head, const uint32_t* __restrict__ block_max, const uint32_t* __restrict__ sa, uint32_t n,
                                      uint32_t* __restrict__ rank, uint32_t* __restrict__ unsorted_flag) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t blk = j / (256 * 8);
    uint32_t g = head[j];
    if (blk > 0) g = max(g, block_max[blk - 1]);
    head[j] = g;
    rank[sa[j]] = g;
    if (g != j) *unsorted_flag = 1;  // position j is not the head of its group -> some group has >= 2 members
}

__global__ void bwt_double_keys_kernel(const uint32_t* __restrict__ rank, uint32_t n, uint32_t h, uint32_t log_n, uint64_t* __restrict__ keys,
                                       uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t hi = rank[i];
    const uint64_t lo = (uint64_t)i + h < n ? (uint64_t)rank[i + h] + 1 : 0;  // <= n < 2^log_n
    keys[i] = (hi << log_n) | lo;
    idx[i] = i;
}

// BWT bytes + primary indexes from the final suffix array (ComputeBWT layout)
__global__ void bwt_emit_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint32_t n,
                                uint8_t* __restrict__ dst /*n bytes*/, uint32_t* __restrict__ primary /*8*/) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) dst[0] = src[n - 1];
    if (r < 8) {
        const uint32_t chunks = bwt_chunks(n);
        uint32_t step = n / chunks;
        if (step * chunks != n) step++;
        primary[r] = (r < chunks && (uint64_t)r * step < n) ? rank[r * step] + 1 : 0;
    }
    if (r >= n) return;
    const uint32_t s = sa[r];
    if (s == 0) return;
    const uint32_t rank0 = rank[0];
    dst[r < rank0 ? r + 1 : r] = src[s - 1];
}

// BWTBlockCodec.Forward header (:113-133). One thread.
__global__ void bwt_header_kernel(const uint32_t* __restrict__ primary, uint32_t n, uint8_t* __restrict__ dst, uint32_t* __restrict__ out_len) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t log_bs = log2_floor(n);
    if (n & (n - 1)) log_bs++;
    const int pidx = (int)(log_bs + 7) >> 3;
    const uint32_t chunks = bwt_chunks(n);
    dst[0] = (uint8_t)((log2_floor(chunks) << 2) | (uint32_t)(pidx - 1));
    uint32_t idx = 1;
    for (uint32_t i = 0; i < chunks; i++) {
        const uint32_t p = primary[i] - 1;
        for (int shift = (pidx - 1) << 3; shift >= 0; shift -= 8) dst[idx++] = (uint8_t)(p >> shift);
    }
    *out_len = n + chunks * pidx + 1;
}

// ---------------- inverse ----------------
// key = BWT byte, value = index of the next suffix' rank for the suffix that starts with this byte (BWT.go:246-258):
// position 0 -> END, positions 1..pidx-1 -> i-1, positions >= pidx -> i
__global__ void ibwt_init_kernel(const uint8_t* __restrict__ src, uint32_t n, uint32_t pidx, uint8_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = src[i];
    vals[i] = i == 0 ? 0xFFFFFFFFu : (i < pidx ? i - 1 : i);
}

static const uint32_t IBWT_STRIDE = 64;  // a sub-list starts at every rank that is a multiple of this (1024 left < 1 warp per SM for a 4 MiB block)

KZ_D bool ibwt_is_head(uint32_t t, const uint32_t* prim, uint32_t nprim) {
    if ((t & (IBWT_STRIDE - 1)) == 0) return true;
    for (uint32_t k = 0; k < nprim; k++)
        if (t == prim[k]) return true;
    return false;
}

// head id of rank t: multiples of the stride first, then the primary ranks
KZ_D uint32_t ibwt_head_id(uint32_t t, uint32_t nstride, const uint32_t* prim, uint32_t nprim) {
    if ((t & (IBWT_STRIDE - 1)) == 0) return t / IBWT_STRIDE;
    for (uint32_t k = 0; k < nprim; k++)
        if (t == prim[k]) return nstride + k;
    return 0xFFFFFFFFu;
}

// walk 1: one thread per sub-list head: length of the sub-list and id of the next head (END = 0xFFFFFFFF)
__global__ void ibwt_measure_kernel(const uint32_t* __restrict__ next, uint32_t n, const uint32_t* __restrict__ prim_rank, uint32_t nprim,
                                    uint32_t nstride, uint32_t* __restrict__ succ, uint32_t* __restrict__ len) {
    // the primary ranks are staged by the first 8 threads of EVERY CTA before any thread leaves: the last CTA holds the primary heads and
    // may have fewer than 8 live threads (a block whose head count is 1..7 modulo the CTA size used to read stale shared memory here)
    __shared__ uint32_t s_prim[8];
    if (threadIdx.x < 8) s_prim[threadIdx.x] = threadIdx.x < nprim ? prim_rank[threadIdx.x] : 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nstride + nprim) return;
    uint32_t t = j < nstride ? j * IBWT_STRIDE : s_prim[j - nstride];
    if (t >= n || (j >= nstride && (t & (IBWT_STRIDE - 1)) == 0)) {  // duplicate head (primary rank on a stride multiple) or out of range
        succ[j] = 0xFFFFFFFEu;                                        // marks "unused"
        len[j] = 0;
        return;
    }
    uint32_t l = 0;
    for (;;) {
        t = next[t];
        l++;
        if (t == 0xFFFFFFFFu) {
            succ[j] = 0xFFFFFFFFu;
            break;
        }
        if (t >= n || l > n) {  // corrupt permutation
            succ[j] = 0xFFFFFFFDu;
            break;
        }
        if (ibwt_is_head(t, s_prim, nprim)) {
            succ[j] = ibwt_head_id(t, nstride, s_prim, nprim);
            break;
        }
    }
    len[j] = l;
}

// Wyllie pointer doubling over the heads: after the rounds dist[j] = number of bytes from head j to the end of the list
__global__ void ibwt_jump_kernel(const uint32_t* __restrict__ succ_in, const uint64_t* __restrict__ dist_in, uint32_t nheads, uint32_t* __restrict__ succ_out,
                                 uint64_t* __restrict__ dist_out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nheads) return;
    const uint32_t s = succ_in[j];
    uint64_t d = dist_in[j];
    uint32_t s2 = s;
    if (s < nheads) {
        d += dist_in[s];
        s2 = succ_in[s];
    }
    succ_out[j] = s2;
    dist_out[j] = d;
}

// walk 2: writes the first byte of every visited rank at its final position
__global__ void ibwt_write_kernel(const uint32_t* __restrict__ next, const uint8_t* __restrict__ first, uint32_t n, const uint32_t* __restrict__ prim_rank,
                                  uint32_t nprim, uint32_t nstride, const uint32_t* __restrict__ succ0, const uint32_t* __restrict__ len,
                                  const uint64_t* __restrict__ dist, uint64_t total, uint8_t* __restrict__ dst, int32_t* __restrict__ status) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nstride + nprim) return;
    if (succ0[j] == 0xFFFFFFFEu) return;
    uint32_t t = j < nstride ? j * IBWT_STRIDE : prim_rank[j - nstride];
    const uint64_t d = dist[j];
    if (d > total) {  // head not on the main list (corrupt input)
        return;
    }
    uint64_t off = total - d;
    const uint32_t l = len[j];
    for (uint32_t i = 0; i < l; i++) {
        if (off + i < n) dst[off + i] = first[t];
        t = next[t];
        if (t >= n) break;
    }
    (void)status;
}

__global__ void ibwt_first_kernel(const uint8_t* __restrict__ sorted_keys, uint32_t n, uint8_t* __restrict__ first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) first[i] = sorted_keys[i];
}

}  // namespace

size_t bwt_forward_workspace(uint32_t n) {
    size_t temp = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)n, 0, 64);
    // keys x2, idx x2, rank, head, block_max, flags, cub temp
    return (size_t)n * (8 * 2 + 4 * 2 + 4 + 4) + 4 * ((size_t)n / 2048 + 16) + 256 + temp + 4096;
}

// d_src: n bytes; d_dst: n + 33 bytes (header + BWT); ws: workspace of bwt_forward_workspace(n) bytes (256-byte aligned)
cudaError_t bwt_forward_device(const uint8_t* d_src, uint32_t n, uint8_t* d_dst, uint32_t* d_out_len, uint8_t* ws, size_t ws_bytes, cudaStream_t stream,
                               uint64_t* launches) {
    if (n < 2) return cudaErrorInvalidValue;
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint64_t* keys_a = (uint64_t*)(ws + off); off = align(off + (size_t)n * 8);
    uint64_t* keys_b = (uint64_t*)(ws + off); off = align(off + (size_t)n * 8);
    uint32_t* idx_a = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* idx_b = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* rank = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* head = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    const uint32_t nscan = (n + 2047) / 2048;
    uint32_t* block_max = (uint32_t*)(ws + off); off = align(off + (size_t)nscan * 4);
    uint32_t* flags = (uint32_t*)(ws + off); off = align(off + 64);
    uint32_t* primary = flags + 4;
    void* temp = ws + off;
    size_t temp_bytes = ws_bytes > off ? ws_bytes - off : 0;
    const uint32_t T = 256, G = (n + T - 1) / T;
    uint32_t log_n = 1;
    while ((1ull << log_n) <= n) log_n++;
    uint32_t header;
    {
        uint32_t log_bs = log2_floor(n);
        if (n & (n - 1)) log_bs++;
        const uint32_t pidx = (log_bs + 7) >> 3;
        if (pidx == 0 || pidx >= 5) return cudaErrorInvalidValue;
        header = (n < 256 ? 1u : 8u) * pidx + 1;
    }
    bwt_init_keys_kernel<<<G, T, 0, stream>>>(d_src, n, keys_a, idx_a);
    (*launches)++;
    uint32_t h = 7;
    int end_bit = 63;
    for (int round = 0; round < 40; round++) {
        cub::DoubleBuffer<uint64_t> k(keys_a, keys_b);
        cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
        size_t tb = temp_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int)n, 0, end_bit, stream);
        if (e != cudaSuccess) return e;
        (*launches) += 8;
        const uint64_t* sk = k.Current();
        const uint32_t* sa = v.Current();
        cudaMemsetAsync(flags, 0, 4, stream);
        bwt_flag_kernel<<<G, T, 0, stream>>>(sk, n, head);
        scan_max_local_kernel<<<nscan, 256, 0, stream>>>(head, n, block_max);
        scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan);
        bwt_apply_rank_kernel<<<G, T, 0, stream>>>(head, block_max, sa, n, rank, flags);
        (*launches) += 4;
        uint32_t unsorted = 0;
        e = cudaMemcpyAsync(&unsorted, flags, 4, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (!unsorted || h >= n) {
            bwt_emit_kernel<<<G, T, 0, stream>>>(d_src, sa, rank, n, d_dst + header, primary);
            bwt_header_kernel<<<1, 32, 0, stream>>>(primary, n, d_dst, d_out_len);
            (*launches) += 2;
            return cudaGetLastError();
        }
        // the sorted buffers become scratch; build the doubled keys into the "a" buffers
        bwt_double_keys_kernel<<<G, T, 0, stream>>>(rank, n, h, log_n, keys_a, idx_a);
        (*launches)++;
        h *= 2;
        end_bit = 2 * (int)log_n;
    }
    return cudaErrorUnknown;
}

// ---- forward, all blocks of a batch in one position space ------------------------------------------------------------------------------------
// A 4 MiB sort cannot fill a B200 and 48 of them one after another pay 48 x (rounds x (sort + 6 small kernels + a host sync)). Here the blocks
// are laid end to end (block b = positions [base[b], base[b] + n[b])) and sorted together: the block index is the top of the first key
// (block | 6 symbols x 9 bits), so ranks taken as global group heads fall into [base[b], base[b] + n[b]) and the doubled keys
// (rank[i] << lg | rank[i + h] + 1, 0 past the END OF THE BLOCK) never mix blocks. One sort, one flag / scan / apply pass and one host
// sync per round for the whole batch; the number of rounds is the maximum over the blocks.
namespace {
struct BwtBatchBlock {
    uint64_t src_off, dst_off;
    uint32_t base, n;
    uint32_t header, pad;
};
const int BWT_BATCH_MAX = 1024;  // 10 bits of block index above 54 bits of symbols
__device__ __forceinline__ uint32_t bwtb_find(const uint32_t* __restrict__ s_base, uint32_t nb, uint32_t i) {  // last block with base <= i
    uint32_t lo = 0, hi = nb;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_base[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256) bwtb_init_keys_kernel(const uint8_t* __restrict__ in, const BwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t total,
                                                              uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    __shared__ uint32_t s_base[BWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base;
    __syncthreads();
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const uint32_t b = bwtb_find(s_base, nb, i);
    const BwtBatchBlock B = blocks[b];
    const uint8_t* src = in + B.src_off;
    const uint32_t l = i - B.base;
    uint64_t k = b;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const uint32_t p = l + j;
        const uint64_t sym = p < B.n ? (uint64_t)src[p] + 1 : 0;
        k = (k << 9) | sym;
    }
    keys[i] = k;
    idx[i] = i;
}
__global__ void __launch_bounds__(256) bwtb_double_keys_kernel(const uint32_t* __restrict__ rank, const BwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t total,
                                                                uint32_t h, uint32_t lg, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    __shared__ uint32_t s_base[BWT_BATCH_MAX];
    __shared__ uint32_t s_n[BWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base, s_n[k] = blocks[k].n;
    __syncthreads();
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const uint32_t b = bwtb_find(s_base, nb, i);
    const uint64_t end = (uint64_t)s_base[b] + s_n[b];
    const uint64_t hi = rank[i];
    const uint64_t lo = (uint64_t)i + h < end ? (uint64_t)rank[i + h] + 1 : 0;  // <= total < 2^lg
    keys[i] = (hi << lg) | lo;
    idx[i] = i;
}
__global__ void __launch_bounds__(256) bwtb_emit_kernel(const uint8_t* __restrict__ in, const BwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t total,
                                                         const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint8_t* __restrict__ out) {
    __shared__ uint32_t s_base[BWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base;
    __syncthreads();
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    if (r >= total) return;
    const uint32_t b = bwtb_find(s_base, nb, r);  // ranks of block b are [base, base + n) as well
    const BwtBatchBlock B = blocks[b];
    const uint8_t* src = in + B.src_off;
    uint8_t* dst = out + B.dst_off + B.header;
    const uint32_t rl = r - B.base;
    if (rl == 0) dst[0] = src[B.n - 1];
    const uint32_t s = sa[r] - B.base;
    if (s == 0) return;
    const uint32_t rank0 = rank[B.base] - B.base;
    dst[rl < rank0 ? rl + 1 : rl] = src[s - 1];
}
// primary indexes + BWTBlockCodec header of every block (BWTBlockCodec.go:113-133); one thread per block
__global__ void bwtb_header_kernel(const BwtBatchBlock* __restrict__ blocks, uint32_t nb, const uint32_t* __restrict__ rank, uint8_t* __restrict__ out,
                                   uint32_t* __restrict__ out_len /* indexed by pad = caller's block number */) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const BwtBatchBlock B = blocks[b];
    const uint32_t n = B.n;
    uint8_t* dst = out + B.dst_off;
    uint32_t log_bs = log2_floor(n);
    if (n & (n - 1)) log_bs++;
    const int pidx = (int)(log_bs + 7) >> 3;
    const uint32_t chunks = bwt_chunks(n);
    uint32_t step = n / chunks;
    if (step * chunks != n) step++;
    dst[0] = (uint8_t)((log2_floor(chunks) << 2) | (uint32_t)(pidx - 1));
    uint32_t idx = 1;
    for (uint32_t i = 0; i < chunks; i++) {
        const uint32_t prim = ((uint64_t)i * step < n) ? rank[B.base + i * step] - B.base + 1 : 0;
        const uint32_t p = prim - 1;
        for (int shift = (pidx - 1) << 3; shift >= 0; shift -= 8) dst[idx++] = (uint8_t)(p >> shift);
    }
    out_len[B.pad] = n + chunks * pidx + 1;
}
// ---- rounds over the unresolved part only -----------------------------------------------------------------------------------------------------
// After a round every group of equal h-prefixes is a contiguous range of sorted positions and only groups with more than one member can
// change. The members of those groups are compacted (sorted position, suffix), sorted by (group head, rank of the suffix h further), and
// written back into the same positions: the work of a round is proportional to what is still unresolved, not to the batch.
__global__ void __launch_bounds__(256) bwtu_flags_full_kernel(const uint32_t* __restrict__ head, uint32_t total, uint8_t* __restrict__ flags) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= total) return;
    const uint32_t hj = head[j];
    flags[j] = (hj != j || (j + 1 < total && head[j + 1] == hj)) ? 1 : 0;
}
__global__ void __launch_bounds__(256) bwtu_flags_kernel(const uint32_t* __restrict__ nhead, const uint32_t* __restrict__ cpos, uint32_t m, uint8_t* __restrict__ flags) {
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    const uint32_t hu = nhead[u];
    flags[u] = (hu != cpos[u] || (u + 1 < m && nhead[u + 1] == hu)) ? 1 : 0;
}
__global__ void __launch_bounds__(256) bwtu_keys_kernel(const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cval, uint32_t m,
                                                         const BwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t h, uint32_t lg, uint64_t* __restrict__ keys) {
    __shared__ uint32_t s_base[BWT_BATCH_MAX];
    __shared__ uint32_t s_n[BWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base, s_n[k] = blocks[k].n;
    __syncthreads();
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    const uint32_t i = cval[u];
    const uint32_t b = bwtb_find(s_base, nb, i);
    const uint64_t end = (uint64_t)s_base[b] + s_n[b];
    const uint64_t hi = rank[i];
    const uint64_t lo = (uint64_t)i + h < end ? (uint64_t)rank[i + h] + 1 : 0;
    keys[u] = (hi << lg) | lo;
}
// candidate heads of the new groups: the sorted position of the first member (0 elsewhere; a max-scan spreads them)
__global__ void __launch_bounds__(256) bwtu_heads_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cpos, uint32_t m, uint32_t* __restrict__ f) {
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    f[u] = (u == 0 || keys[u] != keys[u - 1]) ? cpos[u] : 0u;
}
__global__ void __launch_bounds__(256) bwtu_apply_kernel(uint32_t* __restrict__ f, const uint32_t* __restrict__ block_max, const uint32_t* __restrict__ cpos,
                                                          const uint32_t* __restrict__ cval, uint32_t m, uint32_t* __restrict__ sa, uint32_t* __restrict__ rank) {
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    const uint32_t blk = u / (256 * 8);
    uint32_t g = f[u];
    if (blk > 0) g = max(g, block_max[blk - 1]);
    f[u] = g;
    const uint32_t v = cval[u];
    sa[cpos[u]] = v;
    rank[v] = g;
}
// segmented variant of a later round: the compacted members are already grouped (a group = a contiguous range), so each group is a segment
// sorted by the 32-bit second key alone
__global__ void __launch_bounds__(256) bwtu_keys32_kernel(const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cval, uint32_t m,
                                                           const BwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t h, uint32_t* __restrict__ lo_keys,
                                                           uint8_t* __restrict__ segflag) {
    __shared__ uint32_t s_base[BWT_BATCH_MAX];
    __shared__ uint32_t s_n[BWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base, s_n[k] = blocks[k].n;
    __syncthreads();
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    const uint32_t i = cval[u];
    const uint32_t b = bwtb_find(s_base, nb, i);
    const uint64_t end = (uint64_t)s_base[b] + s_n[b];
    lo_keys[u] = (uint64_t)i + h < end ? rank[i + h] + 1 : 0;
    segflag[u] = (u == 0 || rank[cval[u - 1]] != rank[i]) ? 1 : 0;
}
__global__ void __launch_bounds__(256) bwtu_heads32_kernel(const uint32_t* __restrict__ lo_sorted, const uint8_t* __restrict__ segflag, const uint32_t* __restrict__ cpos,
                                                            uint32_t m, uint32_t* __restrict__ f) {
    const uint32_t u = blockIdx.x * 256 + threadIdx.x;
    if (u >= m) return;
    f[u] = (segflag[u] || lo_sorted[u] != lo_sorted[u - 1]) ? cpos[u] : 0u;
}
}  // namespace

size_t bwt_forward_batch_workspace(uint64_t total, uint32_t nblocks) {
    size_t temp = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int64_t)total, 0, 64);
    size_t temp2 = 0;
    cub::DeviceSelect::Flagged(nullptr, temp2, (const uint32_t*)nullptr, (const uint8_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)total);
    size_t temp3 = 0;
    cub::DeviceSegmentedSort::SortPairs(nullptr, temp3, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)total,
                                        (int64_t)total, (const uint32_t*)nullptr, (const uint32_t*)nullptr);
    return (size_t)total * (8 * 2 + 4 * 2 + 4 + 4 + 4 * 4 + 1) + 4 * ((size_t)total / 2048 + 16) + (size_t)nblocks * sizeof(BwtBatchBlock) + 16 * 256 +
           std::max(std::max(temp, temp2), temp3) + 4096;
}
// at most this many positions / blocks in one call (keys: 2 * lg bits with total < 2^lg; 10 bits of block index in the first key)
uint64_t bwt_forward_batch_max_total() { return (1ull << 30) - 1; }
uint32_t bwt_forward_batch_max_blocks() { return BWT_BATCH_MAX; }

// BWTBlockCodec.Forward of blocks[k] = (src offset, dst offset, n >= 2, caller's block number): d_out_len[number] = n + header
cudaError_t bwt_forward_batch(const uint8_t* d_in, uint8_t* d_out, const uint64_t* src_off, const uint64_t* dst_off, const uint32_t* n_of, const uint32_t* number,
                              uint32_t nb, uint32_t* d_out_len, uint8_t* ws, size_t ws_bytes, cudaStream_t stream, uint64_t* launches) {
    if (nb == 0) return cudaSuccess;
    if (nb > (uint32_t)BWT_BATCH_MAX) return cudaErrorInvalidValue;
    std::vector<BwtBatchBlock> hb(nb);
    uint64_t total64 = 0;
    for (uint32_t k = 0; k < nb; k++) {
        const uint32_t n = n_of[k];
        if (n < 2) return cudaErrorInvalidValue;
        uint32_t log_bs = log2_floor(n);
        if (n & (n - 1)) log_bs++;
        const uint32_t pidx = (log_bs + 7) >> 3;
        if (pidx == 0 || pidx >= 5) return cudaErrorInvalidValue;
        hb[k].src_off = src_off[k], hb[k].dst_off = dst_off[k], hb[k].base = (uint32_t)total64, hb[k].n = n;
        hb[k].header = (n < 256 ? 1u : 8u) * pidx + 1;
        hb[k].pad = number[k];
        total64 += n;
    }
    if (total64 > bwt_forward_batch_max_total()) return cudaErrorInvalidValue;
    const uint32_t total = (uint32_t)total64;
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint64_t* keys_a = (uint64_t*)(ws + off); off = align(off + (size_t)total * 8);
    uint64_t* keys_b = (uint64_t*)(ws + off); off = align(off + (size_t)total * 8);
    uint32_t* idx_a = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* idx_b = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* rank = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* head = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    const uint32_t nscan = (total + 2047) / 2048;
    uint32_t* block_max = (uint32_t*)(ws + off); off = align(off + (size_t)nscan * 4);
    uint32_t* flags = (uint32_t*)(ws + off); off = align(off + 64);
    BwtBatchBlock* d_blocks = (BwtBatchBlock*)(ws + off); off = align(off + (size_t)nb * sizeof(BwtBatchBlock));
    uint32_t* cval_a = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* cval_b = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* cpos_a = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* cpos_b = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint8_t* uflags = ws + off; off = align(off + (size_t)total);
    uint32_t* d_count = flags + 8;
    void* temp = ws + off;
    if (ws_bytes < off) return cudaErrorInvalidValue;
    const size_t temp_bytes = ws_bytes - off;
    cudaError_t e = cudaMemcpyAsync(d_blocks, hb.data(), (size_t)nb * sizeof(BwtBatchBlock), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    const uint32_t T = 256, G = (total + T - 1) / T;
    uint32_t lg = 1;
    while ((1ull << lg) <= total) lg++;
    bwtb_init_keys_kernel<<<G, T, 0, stream>>>(d_in, d_blocks, nb, total, keys_a, idx_a);
    (*launches)++;
    uint32_t h = 6;
    int end_bit = 64;
    uint32_t max_n = 0;
    for (uint32_t k = 0; k < nb; k++) max_n = std::max(max_n, hb[k].n);
    static const bool full_rounds = [] {  // KZ_BWT_FWD=full: every round sorts the whole batch (first batched version)
        const char* ev = getenv("KZ_BWT_FWD");
        return ev && ev[0] == 'f';
    }();
    uint32_t* sa = nullptr;
    // round 0 (and every round under KZ_BWT_FWD=full): the whole batch
    for (int round = 0; round < 40; round++) {
        cub::DoubleBuffer<uint64_t> k(keys_a, keys_b);
        cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
        size_t tb = temp_bytes;
        e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int64_t)total, 0, end_bit, stream);
        if (e != cudaSuccess) return e;
        (*launches) += 8;
        const uint64_t* sk = k.Current();
        sa = v.Current();
        cudaMemsetAsync(flags, 0, 4, stream);
        bwt_flag_kernel<<<G, T, 0, stream>>>(sk, total, head);
        scan_max_local_kernel<<<nscan, 256, 0, stream>>>(head, total, block_max);
        scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan);
        bwt_apply_rank_kernel<<<G, T, 0, stream>>>(head, block_max, sa, total, rank, flags);
        (*launches) += 4;
        uint32_t unsorted = 0;
        e = cudaMemcpyAsync(&unsorted, flags, 4, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (!unsorted || h >= max_n) {
            bwtb_emit_kernel<<<G, T, 0, stream>>>(d_in, d_blocks, nb, total, sa, rank, d_out);
            bwtb_header_kernel<<<(nb + 63) / 64, 64, 0, stream>>>(d_blocks, nb, rank, d_out, d_out_len);
            (*launches) += 2;
            return cudaGetLastError();
        }
        if (!full_rounds) break;
        bwtb_double_keys_kernel<<<G, T, 0, stream>>>(rank, d_blocks, nb, total, h, lg, keys_a, idx_a);
        (*launches)++;
        h *= 2;
        end_bit = 2 * (int)lg;
    }
    if (full_rounds) return cudaErrorUnknown;
    // later rounds: the members of groups with more than one suffix only
    uint32_t m = 0;
    {
        bwtu_flags_full_kernel<<<G, T, 0, stream>>>(head, total, uflags);
        size_t tb = temp_bytes;
        e = cub::DeviceSelect::Flagged(temp, tb, cub::CountingInputIterator<uint32_t>(0), uflags, cpos_a, d_count, (int64_t)total, stream);
        if (e != cudaSuccess) return e;
        tb = temp_bytes;
        e = cub::DeviceSelect::Flagged(temp, tb, (const uint32_t*)sa, uflags, cval_a, d_count, (int64_t)total, stream);
        if (e != cudaSuccess) return e;
        (*launches) += 5;
        e = cudaMemcpyAsync(&m, d_count, 4, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
    }
    end_bit = 2 * (int)lg;
    uint32_t* nhead = head;  // reused: new group heads of the compacted members
    // later rounds: 64-bit radix sorts of (group head, second key) over the compacted members. KZ_BWT_FWD=segmented: one segment per group
    // sorted by the 32-bit second key (cub::DeviceSegmentedSort) — measured no faster on the metric's slab (66.7 vs 65.5 ms per batch)
    static const bool radix_rounds = [] {
        const char* ev = getenv("KZ_BWT_FWD");
        return !(ev && ev[0] == 's' && ev[1] == 'e');
    }();
    uint32_t* lo_in = reinterpret_cast<uint32_t*>(keys_a);
    uint32_t* lo_out = reinterpret_cast<uint32_t*>(keys_b);
    uint32_t* seg_off = reinterpret_cast<uint32_t*>(keys_b) + total;  // second half of keys_b: at most m + 1 offsets
    for (int round = 1; round < 48 && m > 0; round++) {
        const uint32_t GM = (m + T - 1) / T, nscan_m = (m + 2047) / 2048;
        const uint32_t* sorted_val = nullptr;
        uint32_t* other = nullptr;
        if (radix_rounds) {
            bwtu_keys_kernel<<<GM, T, 0, stream>>>(rank, cval_a, m, d_blocks, nb, h, lg, keys_a);
            cub::DoubleBuffer<uint64_t> k(keys_a, keys_b);
            cub::DoubleBuffer<uint32_t> v(cval_a, cval_b);
            size_t tb = temp_bytes;
            e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int64_t)m, 0, end_bit, stream);
            if (e != cudaSuccess) return e;
            bwtu_heads_kernel<<<GM, T, 0, stream>>>(k.Current(), cpos_a, m, nhead);
            sorted_val = v.Current();
            other = v.Current() == cval_a ? cval_b : cval_a;
        } else {
            // the groups are contiguous ranges of the compacted order: one segment each, sorted by the second key alone
            bwtu_keys32_kernel<<<GM, T, 0, stream>>>(rank, cval_a, m, d_blocks, nb, h, lo_in, uflags);
            size_t tb = temp_bytes;
            e = cub::DeviceSelect::Flagged(temp, tb, cub::CountingInputIterator<uint32_t>(0), uflags, seg_off, d_count, (int64_t)m, stream);
            if (e != cudaSuccess) return e;
            uint32_t nseg = 0;
            if ((e = cudaMemcpyAsync(&nseg, d_count, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(seg_off + nseg, &m, 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
            tb = temp_bytes;
            e = cub::DeviceSegmentedSort::SortPairs(temp, tb, (const uint32_t*)lo_in, lo_out, (const uint32_t*)cval_a, cval_b, (int64_t)m, (int64_t)nseg, seg_off,
                                                    seg_off + 1, stream);
            if (e != cudaSuccess) return e;
            bwtu_heads32_kernel<<<GM, T, 0, stream>>>(lo_out, uflags, cpos_a, m, nhead);
            sorted_val = cval_b;
            other = cval_a;
        }
        scan_max_local_kernel<<<nscan_m, 256, 0, stream>>>(nhead, m, block_max);
        scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan_m);
        bwtu_apply_kernel<<<GM, T, 0, stream>>>(nhead, block_max, cpos_a, sorted_val, m, sa, rank);
        (*launches) += 13;
        h *= 2;
        if (h >= 2 * (uint64_t)max_n) break;  // every suffix pair has been compared past the end of its block
        bwtu_flags_kernel<<<GM, T, 0, stream>>>(nhead, cpos_a, m, uflags);
        size_t tb = temp_bytes;
        e = cub::DeviceSelect::Flagged(temp, tb, (const uint32_t*)cpos_a, uflags, cpos_b, d_count, (int64_t)m, stream);
        if (e != cudaSuccess) return e;
        tb = temp_bytes;
        e = cub::DeviceSelect::Flagged(temp, tb, sorted_val, uflags, other, d_count, (int64_t)m, stream);
        if (e != cudaSuccess) return e;
        (*launches) += 5;
        e = cudaMemcpyAsync(&m, d_count, 4, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        std::swap(cpos_a, cpos_b);
        if (other != cval_a) std::swap(cval_a, cval_b);  // the compacted suffixes are the next round's input
    }
    bwtb_emit_kernel<<<G, T, 0, stream>>>(d_in, d_blocks, nb, total, sa, rank, d_out);
    bwtb_header_kernel<<<(nb + 63) / 64, 64, 0, stream>>>(d_blocks, nb, rank, d_out, d_out_len);
    (*launches) += 2;
    return cudaGetLastError();
}

size_t bwt_inverse_workspace(uint32_t n) {
    size_t temp = 0;
    cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)n, 0, 8);
    const size_t nheads = (size_t)n / IBWT_STRIDE + 16;
    return (size_t)n * (1 * 2 + 4 * 2 + 1) + nheads * (4 * 3 + 8 * 2) + 4096 + temp + 4096;
}

// d_src: header + BWT bytes (len bytes); d_dst: receives the original block. Returns cudaErrorInvalidValue for malformed headers.
cudaError_t bwt_inverse_device(const uint8_t* d_src, const uint8_t* h_header /*first 33 bytes on the host*/, uint32_t len, uint8_t* d_dst, uint32_t cap,
                               uint32_t* out_len, uint8_t* ws, size_t ws_bytes, cudaStream_t stream, uint64_t* launches) {
    if (len < 2) return cudaErrorInvalidValue;
    const uint8_t mode = h_header[0];
    const uint32_t log_chunks = (mode >> 2) & 7;
    const int pidx_size = (int)(mode & 3) + 1;
    const uint32_t chunks = 1u << log_chunks;
    const uint32_t header = chunks * pidx_size + 1;
    if (len < header || chunks > 8) return cudaErrorInvalidValue;
    const uint32_t n = len - header;
    if (chunks != (n < 256 ? 1u : 8u)) return cudaErrorInvalidValue;
    *out_len = n;
    if (n == 0) return cudaSuccess;
    if (n > cap) return cudaErrorInvalidValue;
    uint32_t prim[8];
    for (uint32_t i = 0, idx = 1; i < chunks; i++) {
        uint32_t p = 0;
        for (int k = 0; k < pidx_size; k++) p = (p << 8) | h_header[idx++];
        prim[i] = p + 1;
        if (prim[i] == 0 || prim[i] > n) return cudaErrorInvalidValue;
    }
    const uint8_t* bw = d_src + header;
    if (n == 1) return cudaMemcpyAsync(d_dst, bw, 1, cudaMemcpyDeviceToDevice, stream);
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint8_t* keys_a = ws + off; off = align(off + n);
    uint8_t* keys_b = ws + off; off = align(off + n);
    uint32_t* val_a = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* val_b = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    const uint32_t nstride = (n + IBWT_STRIDE - 1) / IBWT_STRIDE;
    const uint32_t nprim = chunks;
    const uint32_t nheads = nstride + nprim;
    uint32_t* succ_a = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* succ_b = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* succ0 = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* lenv = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint64_t* dist_a = (uint64_t*)(ws + off); off = align(off + (size_t)nheads * 8);
    uint64_t* dist_b = (uint64_t*)(ws + off); off = align(off + (size_t)nheads * 8);
    uint32_t* d_prim = (uint32_t*)(ws + off); off = align(off + 64);
    void* temp = ws + off;
    size_t temp_bytes = ws_bytes > off ? ws_bytes - off : 0;
    uint32_t prim_rank[8];
    for (uint32_t i = 0; i < 8; i++) prim_rank[i] = i < chunks ? prim[i] - 1 : 0xFFFFFFFFu;
    cudaError_t e = cudaMemcpyAsync(d_prim, prim_rank, sizeof(prim_rank), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    const uint32_t T = 256, G = (n + T - 1) / T;
    ibwt_init_kernel<<<G, T, 0, stream>>>(bw, n, prim[0], keys_a, val_a);
    cub::DoubleBuffer<uint8_t> k(keys_a, keys_b);
    cub::DoubleBuffer<uint32_t> v(val_a, val_b);
    size_t tb = temp_bytes;
    e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int)n, 0, 8, stream);  // stable: LF mapping
    if (e != cudaSuccess) return e;
    const uint8_t* first = k.Current();
    const uint32_t* next = v.Current();
    const uint32_t GH = (nheads + T - 1) / T;
    ibwt_measure_kernel<<<GH, T, 0, stream>>>(next, n, d_prim, nprim, nstride, succ0, lenv);
    // dist[j] = len[j] initially; doubling
    // (len -> 64-bit) reuse jump kernel with succ pointing nowhere for the first copy
    {
        // initialise dist_a from len, succ_a from succ0
        cudaMemcpyAsync(succ_a, succ0, (size_t)nheads * 4, cudaMemcpyDeviceToDevice, stream);
        // widen len into dist_a with a tiny kernel-free trick: use cudaMemset + 2D copy of the low words
        cudaMemsetAsync(dist_a, 0, (size_t)nheads * 8, stream);
        cudaMemcpy2DAsync(dist_a, 8, lenv, 4, 4, nheads, cudaMemcpyDeviceToDevice, stream);
    }
    (*launches) += 4;
    uint32_t rounds = 1;
    while ((1u << rounds) < nheads) rounds++;
    for (uint32_t r = 0; r < rounds + 1; r++) {
        ibwt_jump_kernel<<<GH, T, 0, stream>>>(succ_a, dist_a, nheads, succ_b, dist_b);
        std::swap(succ_a, succ_b);
        std::swap(dist_a, dist_b);
        (*launches)++;
    }
    ibwt_write_kernel<<<GH, T, 0, stream>>>(next, first, n, d_prim, nprim, nstride, succ0, lenv, dist_a, (uint64_t)n, d_dst, nullptr);
    (*launches)++;
    return cudaGetLastError();
}


// ---- inverse, all blocks of a batch together ---------------------------------------------------------------------------------------------------
// Same algorithm as bwt_inverse_device; one position space for the batch (block b = [base, base + n)), the block index above the byte in the
// key of the stable sort (so the LF mapping of block b lands in its own range), one head list for all blocks, one set of launches: a 4 MiB
// block is 17 pointer-doubling launches of a few hundred threads each, 48 blocks one after another were mostly launch latency.
namespace {
struct IbwtBatchBlock {
    uint64_t src_off, dst_off;  // BWT bytes (behind the header) / output
    uint32_t base, n;
    uint32_t hbase, nstride;    // first head id of the block, heads on stride multiples (followed by the nprim primary heads)
    uint32_t nprim, pad;
    uint32_t prim[8];           // primary ranks (local), 0xFFFFFFFF when absent
};
const int IBWT_BATCH_MAX = 1024;
__device__ __forceinline__ uint32_t ibwtb_find(const uint32_t* __restrict__ s_base, uint32_t nb, uint32_t i) {
    uint32_t lo = 0, hi = nb;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_base[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256) ibwtb_init_kernel(const uint8_t* __restrict__ in, const IbwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t total,
                                                          uint16_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint32_t s_base[IBWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_base[k] = blocks[k].base;
    __syncthreads();
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= total) return;
    const uint32_t b = ibwtb_find(s_base, nb, g);
    const IbwtBatchBlock& B = blocks[b];
    const uint32_t i = g - B.base;
    const uint32_t pidx = B.prim[0] + 1;
    keys[g] = (uint16_t)((b << 8) | in[B.src_off + i]);
    vals[g] = i == 0 ? 0xFFFFFFFFu : B.base + (i < pidx ? i - 1 : i);
}
__device__ __forceinline__ bool ibwtb_is_head(uint32_t t_local, const IbwtBatchBlock& B) {
    if ((t_local & (IBWT_STRIDE - 1)) == 0) return true;
    for (uint32_t k = 0; k < B.nprim; k++)
        if (t_local == B.prim[k]) return true;
    return false;
}
__device__ __forceinline__ uint32_t ibwtb_head_id(uint32_t t_local, const IbwtBatchBlock& B) {
    if ((t_local & (IBWT_STRIDE - 1)) == 0) return B.hbase + t_local / IBWT_STRIDE;
    for (uint32_t k = 0; k < B.nprim; k++)
        if (t_local == B.prim[k]) return B.hbase + B.nstride + k;
    return 0xFFFFFFFFu;
}
__global__ void __launch_bounds__(256) ibwtb_measure_kernel(const uint32_t* __restrict__ next, const IbwtBatchBlock* __restrict__ blocks, uint32_t nb, uint32_t nheads,
                                                             uint32_t* __restrict__ succ, uint32_t* __restrict__ len, uint64_t* __restrict__ dist0, uint32_t* __restrict__ bad) {
    __shared__ uint32_t s_hbase[IBWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_hbase[k] = blocks[k].hbase;
    __syncthreads();
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= nheads) return;
    const uint32_t b = ibwtb_find(s_hbase, nb, j);
    const IbwtBatchBlock B = blocks[b];
    const uint32_t jl = j - B.hbase;
    uint32_t t = jl < B.nstride ? jl * IBWT_STRIDE : B.prim[jl - B.nstride];
    if (t >= B.n || (jl >= B.nstride && (t & (IBWT_STRIDE - 1)) == 0)) {  // duplicate head or out of range
        succ[j] = 0xFFFFFFFEu;
        len[j] = 0;
        dist0[j] = 0;
        return;
    }
    uint32_t l = 0;
    uint32_t g = B.base + t;
    for (;;) {
        g = next[g];
        l++;
        if (g == 0xFFFFFFFFu) {
            succ[j] = 0xFFFFFFFFu;
            break;
        }
        const uint32_t tl = g - B.base;
        if (g < B.base || tl >= B.n || l > B.n) {  // corrupt permutation
            succ[j] = 0xFFFFFFFDu;
            *bad = 1;
            break;
        }
        if (ibwtb_is_head(tl, B)) {
            succ[j] = ibwtb_head_id(tl, B);
            break;
        }
    }
    len[j] = l;
    dist0[j] = l;
}
__global__ void __launch_bounds__(256) ibwtb_write_kernel(const uint32_t* __restrict__ next, const uint16_t* __restrict__ first, const IbwtBatchBlock* __restrict__ blocks,
                                                           uint32_t nb, uint32_t nheads, const uint32_t* __restrict__ succ0, const uint32_t* __restrict__ len,
                                                           const uint64_t* __restrict__ dist, uint8_t* __restrict__ out, uint32_t* __restrict__ bad) {
    __shared__ uint32_t s_hbase[IBWT_BATCH_MAX];
    for (uint32_t k = threadIdx.x; k < nb; k += 256) s_hbase[k] = blocks[k].hbase;
    __syncthreads();
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= nheads) return;
    if (succ0[j] == 0xFFFFFFFEu) return;
    const uint32_t b = ibwtb_find(s_hbase, nb, j);
    const IbwtBatchBlock B = blocks[b];
    const uint32_t jl = j - B.hbase;
    const uint32_t t = jl < B.nstride ? jl * IBWT_STRIDE : B.prim[jl - B.nstride];
    const uint64_t d = dist[j];
    if (d > B.n) {  // head not on the main list (corrupt input: the permutation has more than one cycle)
        *bad = 1;
        return;
    }
    uint8_t* dst = out + B.dst_off;
    const uint64_t off = (uint64_t)B.n - d;
    const uint32_t l = len[j];
    uint32_t g = B.base + t;
    for (uint32_t i = 0; i < l; i++) {
        if (off + i < B.n) dst[off + i] = (uint8_t)first[g];
        g = next[g];
        if (g - B.base >= B.n) break;
    }
}
}  // namespace

size_t bwt_inverse_batch_workspace(uint64_t total, uint32_t nblocks) {
    size_t temp = 0;
    cub::DoubleBuffer<uint16_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int64_t)total, 0, 16);
    const size_t nheads = (size_t)total / IBWT_STRIDE + (size_t)nblocks * 10 + 16;
    return (size_t)total * (2 * 2 + 4 * 2) + nheads * (4 * 4 + 8 * 2) + (size_t)nblocks * sizeof(IbwtBatchBlock) + 17 * 256 + temp + 4096;
}
uint64_t bwt_inverse_batch_max_total() { return (1ull << 31) - 1; }
uint32_t bwt_inverse_batch_max_blocks() { return 256; }  // 8 bits of block index above the byte in a 16-bit key

// BWTBlockCodec.Inverse of blocks k = 0..nb-1: stream at d_in + src_off[k] (len_of[k] bytes), h_headers + 64 k = its first 33 bytes on the host.
// out_len[k] = decoded length. cudaErrorInvalidValue: a malformed header (checked before anything is launched).
cudaError_t bwt_inverse_batch(const uint8_t* d_in, uint8_t* d_out, const uint64_t* src_off, const uint64_t* dst_off, const uint32_t* len_of, const uint8_t* h_headers,
                              uint32_t nb, uint32_t cap, uint32_t* out_len, uint8_t* ws, size_t ws_bytes, cudaStream_t stream, uint64_t* launches) {
    if (nb == 0) return cudaSuccess;
    if (nb > bwt_inverse_batch_max_blocks()) return cudaErrorInvalidValue;
    std::vector<IbwtBatchBlock> hb;
    hb.reserve(nb);
    uint64_t total64 = 0, heads64 = 0;
    struct Tiny {
        uint64_t src, dst;
    };
    std::vector<Tiny> singles;
    for (uint32_t k = 0; k < nb; k++) {
        const uint32_t len = len_of[k];
        const uint8_t* hh = h_headers + (size_t)k * 64;
        if (len < 2) return cudaErrorInvalidValue;
        const uint8_t mode = hh[0];
        const uint32_t log_chunks = (mode >> 2) & 7;
        const int pidx_size = (int)(mode & 3) + 1;
        const uint32_t chunks = 1u << log_chunks;
        const uint32_t header = chunks * pidx_size + 1;
        if (len < header || chunks > 8) return cudaErrorInvalidValue;
        const uint32_t n = len - header;
        if (chunks != (n < 256 ? 1u : 8u)) return cudaErrorInvalidValue;
        out_len[k] = n;
        if (n == 0) continue;
        if (n > cap) return cudaErrorInvalidValue;
        IbwtBatchBlock B;
        for (uint32_t i = 0; i < 8; i++) B.prim[i] = 0xFFFFFFFFu;
        for (uint32_t i = 0, idx = 1; i < chunks; i++) {
            uint32_t p = 0;
            for (int q = 0; q < pidx_size; q++) p = (p << 8) | hh[idx++];
            const uint32_t pr = p + 1;
            if (pr == 0 || pr > n) return cudaErrorInvalidValue;
            B.prim[i] = pr - 1;
        }
        if (n == 1) {
            singles.push_back(Tiny{src_off[k] + header, dst_off[k]});
            continue;
        }
        B.src_off = src_off[k] + header, B.dst_off = dst_off[k], B.base = (uint32_t)total64, B.n = n;
        B.hbase = (uint32_t)heads64, B.nstride = (n + IBWT_STRIDE - 1) / IBWT_STRIDE, B.nprim = chunks, B.pad = 0;
        total64 += n;
        heads64 += B.nstride + B.nprim;
        hb.push_back(B);
    }
    cudaError_t e;
    for (const Tiny& t : singles)
        if ((e = cudaMemcpyAsync(d_out + t.dst, d_in + t.src, 1, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    if (hb.empty()) return cudaSuccess;
    if (total64 > bwt_inverse_batch_max_total()) return cudaErrorInvalidValue;
    const uint32_t total = (uint32_t)total64, nheads = (uint32_t)heads64, nbb = (uint32_t)hb.size();
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint16_t* keys_a = (uint16_t*)(ws + off); off = align(off + (size_t)total * 2);
    uint16_t* keys_b = (uint16_t*)(ws + off); off = align(off + (size_t)total * 2);
    uint32_t* val_a = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* val_b = (uint32_t*)(ws + off); off = align(off + (size_t)total * 4);
    uint32_t* succ_a = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* succ_b = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* succ0 = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint32_t* lenv = (uint32_t*)(ws + off); off = align(off + (size_t)nheads * 4);
    uint64_t* dist_a = (uint64_t*)(ws + off); off = align(off + (size_t)nheads * 8);
    uint64_t* dist_b = (uint64_t*)(ws + off); off = align(off + (size_t)nheads * 8);
    IbwtBatchBlock* d_blocks = (IbwtBatchBlock*)(ws + off); off = align(off + (size_t)nbb * sizeof(IbwtBatchBlock));
    uint32_t* d_bad = (uint32_t*)(ws + off); off = align(off + 64);
    void* temp = ws + off;
    if (ws_bytes < off) return cudaErrorInvalidValue;
    const size_t temp_bytes = ws_bytes - off;
    if ((e = cudaMemcpyAsync(d_blocks, hb.data(), (size_t)nbb * sizeof(IbwtBatchBlock), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_bad, 0, 4, stream)) != cudaSuccess) return e;
    // a corrupt payload (several cycles, a bad primary index) must not leave parts of the output unwritten: clear it first, report it afterwards
    for (const IbwtBatchBlock& B : hb)
        if ((e = cudaMemsetAsync(d_out + B.dst_off, 0, B.n, stream)) != cudaSuccess) return e;
    const uint32_t T = 256, G = (total + T - 1) / T, GH = (nheads + T - 1) / T;
    ibwtb_init_kernel<<<G, T, 0, stream>>>(d_in, d_blocks, nbb, total, keys_a, val_a);
    cub::DoubleBuffer<uint16_t> k(keys_a, keys_b);
    cub::DoubleBuffer<uint32_t> v(val_a, val_b);
    size_t tb = temp_bytes;
    int key_bits = 8;
    while ((1u << (key_bits - 8)) < nbb) key_bits++;
    if ((e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int64_t)total, 0, key_bits, stream)) != cudaSuccess) return e;  // stable: LF mapping per block
    const uint16_t* first = k.Current();
    const uint32_t* next = v.Current();
    ibwtb_measure_kernel<<<GH, T, 0, stream>>>(next, d_blocks, nbb, nheads, succ0, lenv, dist_a, d_bad);
    if ((e = cudaMemcpyAsync(succ_a, succ0, (size_t)nheads * 4, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    (*launches) += 8;
    uint32_t max_heads = 1;
    for (const IbwtBatchBlock& B : hb) max_heads = std::max(max_heads, B.nstride + B.nprim);
    uint32_t rounds = 1;
    while ((1u << rounds) < max_heads) rounds++;
    for (uint32_t r = 0; r < rounds + 1; r++) {
        ibwt_jump_kernel<<<GH, T, 0, stream>>>(succ_a, dist_a, nheads, succ_b, dist_b);
        std::swap(succ_a, succ_b);
        std::swap(dist_a, dist_b);
        (*launches)++;
    }
    ibwtb_write_kernel<<<GH, T, 0, stream>>>(next, first, d_blocks, nbb, nheads, succ0, lenv, dist_a, d_out, d_bad);
    (*launches)++;
    uint32_t h_bad = 0;
    if ((e = cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    if (h_bad) return cudaErrorInvalidValue;  // the permutation of some block is not one list through its n ranks
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// BWTS (bijective BWT, transform/BWTS.go:58-170 Forward, :211-288 Inverse).
//
// Forward. The reference builds the suffix array, finds the Lyndon factors (a factor starts where the rank of the
// suffix is smaller than every earlier one, :120-125) and repairs the order of the rotations factor by factor. The
// result is canonical: all rotations of all Lyndon factors sorted by their infinite periodic repetition, output byte =
// cyclic predecessor inside the factor (equal rotations of repeated factors carry equal bytes, so their mutual order
// does not matter). On the GPU: (1) suffix ranks by the same prefix doubling as the BWT, (2) factor starts by a
// prefix-minimum over the ranks, (3) a second prefix doubling in which "the position h bytes further" follows the
// cyclic successor inside the factor (succ_2h = succ_h o succ_h) until no group splits any more, (4) scatter.
// Inverse. The reference builds the LF mapping and walks its cycles, smallest unvisited index first, filling the output
// from the end. On the GPU: LF by one stable radix sort, cycle leaders (minimum index of a cycle) by pointer doubling of
// minima, distance from the leader by list ranking of the cycles cut at their leader, cycle offsets by a prefix sum.
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void bwts_negrank_kernel(const uint32_t* __restrict__ rank, uint32_t n, uint32_t* __restrict__ v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = n - rank[i];  // >= 1; larger = smaller rank
}
// in: v (n - rank), m = inclusive prefix maximum of v (finished with block_max). out: head position for the next scan
__global__ void bwts_heads_kernel(const uint32_t* __restrict__ v, const uint32_t* __restrict__ m, const uint32_t* __restrict__ block_max, uint32_t n,
                                  uint32_t* __restrict__ hpos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = i == 0;
    if (i > 0) {
        const uint32_t j = i - 1;
        const uint32_t blk = j / 2048;
        uint32_t before = m[j];
        if (blk > 0) before = max(before, block_max[blk - 1]);
        head = v[i] > before;  // rank[i] smaller than every earlier rank
    }
    hpos[i] = head ? i : 0u;
}
// fstart = inclusive prefix maximum of hpos (local part in fs, block part in block_max): succ1 / pred inside the factor
__global__ void bwts_succ_kernel(const uint32_t* __restrict__ hpos, const uint32_t* __restrict__ fs, const uint32_t* __restrict__ block_max, uint32_t n,
                                 uint32_t* __restrict__ succ, uint32_t* __restrict__ pred) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t blk = i / 2048;
    uint32_t start = fs[i];
    if (blk > 0) start = max(start, block_max[blk - 1]);
    const bool last = i + 1 == n || hpos[i + 1] == i + 1;
    const uint32_t sx = last ? start : i + 1;
    succ[i] = sx;
    pred[sx] = i;
}
// first key: 7 bytes along the cyclic successor (8 bits each are enough: every rotation has all 7), succ7 for the doubling
__global__ void bwts_init_keys_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ succ1, uint32_t n, uint64_t* __restrict__ keys,
                                      uint32_t* __restrict__ idx, uint32_t* __restrict__ succ7) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
    uint32_t p = i;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        k = (k << 9) | ((uint64_t)src[p] + 1);
        p = succ1[p];
    }
    keys[i] = k;
    idx[i] = i;
    succ7[i] = p;
}
__global__ void bwts_double_keys_kernel(const uint32_t* __restrict__ rank, const uint32_t* __restrict__ succ_h, uint32_t n, uint32_t log_n,
                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t* __restrict__ succ_2h) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = succ_h[i];
    keys[i] = ((uint64_t)rank[i] << log_n) | (uint64_t)rank[s];
    idx[i] = i;
    succ_2h[i] = succ_h[s];
}
// counts the groups (positions that head their group) of the current order
__global__ void bwts_count_groups_kernel(const uint32_t* __restrict__ head, uint32_t n, uint32_t* __restrict__ count) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool is_head = j < n && head[j] == j;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, is_head);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (uint32_t)__popc(b));
}
__global__ void bwts_emit_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ pred, uint32_t n,
                                 uint8_t* __restrict__ dst) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[pred[sa[j]]];
}

// ---- inverse
__global__ void ibwts_init_kernel(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = src[i];
        vals[i] = i;
    }
}
// sorted position s holds original index v[s]  ->  lf[v[s]] = s ; min/jump initialisation
__global__ void ibwts_lf_kernel(const uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ lf) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) lf[v[s]] = s;
}
__global__ void ibwts_min_init_kernel(const uint32_t* __restrict__ lf, uint32_t n, uint32_t* __restrict__ mn, uint32_t* __restrict__ jp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t t = lf[i];
        mn[i] = min(i, t);
        jp[i] = t;
    }
}
__global__ void ibwts_min_jump_kernel(const uint32_t* __restrict__ mn_in, const uint32_t* __restrict__ jp_in, uint32_t n, uint32_t* __restrict__ mn_out,
                                      uint32_t* __restrict__ jp_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t j = jp_in[i];
        mn_out[i] = min(mn_in[i], mn_in[j]);
        jp_out[i] = jp_in[j];
    }
}
// lists = cycles cut in front of their leader: nx[p] = lf[p], END when lf[p] leads its cycle; cnt[p] = 1
__global__ void ibwts_rank_init_kernel(const uint32_t* __restrict__ lf, const uint32_t* __restrict__ mn, uint32_t n, uint32_t* __restrict__ nx,
                                       uint32_t* __restrict__ cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t t = lf[i];
        nx[i] = mn[t] == t ? 0xFFFFFFFFu : t;
        cnt[i] = 1;
    }
}
__global__ void ibwts_rank_jump_kernel(const uint32_t* __restrict__ nx_in, const uint32_t* __restrict__ cnt_in, uint32_t n, uint32_t* __restrict__ nx_out,
                                       uint32_t* __restrict__ cnt_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t j = nx_in[i];
        uint32_t c = cnt_in[i];
        uint32_t j2 = j;
        if (j != 0xFFFFFFFFu) {
            c += cnt_in[j];
            j2 = nx_in[j];
        }
        nx_out[i] = j2;
        cnt_out[i] = c;
    }
}
// cnt[p] = nodes from p to the end of its list; a leader's cnt is the cycle length
__global__ void ibwts_len_kernel(const uint32_t* __restrict__ mn, const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ len_at) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len_at[i] = mn[i] == i ? cnt[i] : 0u;
}
// exclusive sum over len_at: three-phase like scan_max (values fit 32 bits: total = n)
__global__ void scan_add_local_kernel(uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base = blockIdx.x * (blockDim.x * 8);
    uint32_t x[8];
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t j = base + tid * 8 + q;
        const uint32_t a = j < n ? v[j] : 0u;
        x[q] = run;  // exclusive inside the thread
        run += a;
    }
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t prev = incl - run;
    if (warp > 0) prev += s_warp[warp - 1];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t j = base + tid * 8 + q;
        if (j < n) v[j] = x[q] + prev;
    }
    if (tid == blockDim.x - 1) block_sum[blockIdx.x] = prev + run;
}
__global__ void scan_add_blocks_kernel(uint32_t* __restrict__ block_sum, uint32_t nblocks) {  // one thread: nblocks = n / 2048 is small
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < nblocks; i++) {
        const uint32_t t = block_sum[i];
        block_sum[i] = acc;
        acc += t;
    }
}
// dst[n - 1 - (offset of the cycle + steps from the leader)] = src[p]
__global__ void ibwts_write_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ mn, const uint32_t* __restrict__ cnt,
                                   const uint32_t* __restrict__ excl, const uint32_t* __restrict__ block_sum, uint32_t n, uint8_t* __restrict__ dst) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t leader = mn[p];
    const uint32_t len = cnt[leader];
    const uint32_t d = len - cnt[p];
    const uint32_t off = excl[leader] + block_sum[leader / 2048];
    const uint64_t pos = (uint64_t)off + d;
    if (pos < n) dst[n - 1 - (uint32_t)pos] = src[p];
}

}  // namespace

size_t bwts_forward_workspace(uint32_t n) { return bwt_forward_workspace(n) + (size_t)n * 4 * 4 + 4096; }
size_t bwts_inverse_workspace(uint32_t n) {
    size_t temp = 0;
    cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int)n, 0, 8);
    return (size_t)n * (1 * 2 + 4 * 7) + 4 * ((size_t)n / 2048 + 16) + 4096 + temp + 4096;
}

// sorts keys_a/idx_a (n pairs), ranks the groups; returns the sorted order in *sa_out, group ranks in `rank`; flags[0] = 1 when some
// group has several members, flags[1] = number of groups
static cudaError_t bwt_sort_round(uint64_t* keys_a, uint64_t* keys_b, uint32_t* idx_a, uint32_t* idx_b, uint32_t n, int end_bit, void* temp,
                                  size_t temp_bytes, uint32_t* head, uint32_t* block_max, uint32_t nscan, uint32_t* rank, uint32_t* flags,
                                  const uint32_t** sa_out, bool count_groups, cudaStream_t stream, uint64_t* launches) {
    const uint32_t T = 256, G = (n + T - 1) / T;
    cub::DoubleBuffer<uint64_t> k(keys_a, keys_b);
    cub::DoubleBuffer<uint32_t> v(idx_a, idx_b);
    size_t tb = temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int)n, 0, end_bit, stream);
    if (e != cudaSuccess) return e;
    (*launches) += 8;
    const uint64_t* sk = k.Current();
    const uint32_t* sa = v.Current();
    cudaMemsetAsync(flags, 0, 8, stream);
    bwt_flag_kernel<<<G, T, 0, stream>>>(sk, n, head);
    scan_max_local_kernel<<<nscan, 256, 0, stream>>>(head, n, block_max);
    scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan);
    bwt_apply_rank_kernel<<<G, T, 0, stream>>>(head, block_max, sa, n, rank, flags);
    (*launches) += 4;
    if (count_groups) {
        bwts_count_groups_kernel<<<G, T, 0, stream>>>(head, n, flags + 1);
        (*launches)++;
    }
    *sa_out = sa;
    return cudaGetLastError();
}

// BWTS.Forward of one block: d_dst receives n bytes
cudaError_t bwts_forward_device(const uint8_t* d_src, uint32_t n, uint8_t* d_dst, uint8_t* ws, size_t ws_bytes, cudaStream_t stream, uint64_t* launches) {
    if (n < 2) return n == 1 ? cudaMemcpyAsync(d_dst, d_src, 1, cudaMemcpyDeviceToDevice, stream) : cudaSuccess;
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint64_t* keys_a = (uint64_t*)(ws + off); off = align(off + (size_t)n * 8);
    uint64_t* keys_b = (uint64_t*)(ws + off); off = align(off + (size_t)n * 8);
    uint32_t* idx_a = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* idx_b = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* rank = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* head = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* succ_a = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* succ_b = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* pred = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* aux = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    const uint32_t nscan = (n + 2047) / 2048;
    uint32_t* block_max = (uint32_t*)(ws + off); off = align(off + (size_t)nscan * 4);
    uint32_t* flags = (uint32_t*)(ws + off); off = align(off + 64);
    void* temp = ws + off;
    const size_t temp_bytes = ws_bytes > off ? ws_bytes - off : 0;
    const uint32_t T = 256, G = (n + T - 1) / T;
    uint32_t log_n = 1;
    while ((1ull << log_n) <= n) log_n++;
    // ---- (1) suffix ranks
    bwt_init_keys_kernel<<<G, T, 0, stream>>>(d_src, n, keys_a, idx_a);
    (*launches)++;
    uint32_t h = 7;
    int end_bit = 63;
    const uint32_t* sa = nullptr;
    bool done = false;
    for (int round = 0; round < 40 && !done; round++) {
        cudaError_t e = bwt_sort_round(keys_a, keys_b, idx_a, idx_b, n, end_bit, temp, temp_bytes, head, block_max, nscan, rank, flags, &sa, false, stream, launches);
        if (e != cudaSuccess) return e;
        uint32_t unsorted = 0;
        e = cudaMemcpyAsync(&unsorted, flags, 4, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (!unsorted || h >= n) {
            done = true;
            break;
        }
        bwt_double_keys_kernel<<<G, T, 0, stream>>>(rank, n, h, log_n, keys_a, idx_a);
        (*launches)++;
        h *= 2;
        end_bit = 2 * (int)log_n;
    }
    if (!done) return cudaErrorUnknown;
    // ---- (2) Lyndon factor starts: rank smaller than every earlier rank; cyclic successor / predecessor inside the factor
    bwts_negrank_kernel<<<G, T, 0, stream>>>(rank, n, aux);
    cudaMemcpyAsync(head, aux, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream);
    scan_max_local_kernel<<<nscan, 256, 0, stream>>>(head, n, block_max);
    scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan);
    bwts_heads_kernel<<<G, T, 0, stream>>>(aux, head, block_max, n, succ_b /*hpos*/);
    cudaMemcpyAsync(head, succ_b, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream);
    scan_max_local_kernel<<<nscan, 256, 0, stream>>>(head, n, block_max);
    scan_max_blocks_kernel<<<1, 1024, 0, stream>>>(block_max, nscan);
    bwts_succ_kernel<<<G, T, 0, stream>>>(succ_b /*hpos*/, head, block_max, n, aux /*succ1*/, pred);
    (*launches) += 7;
    // ---- (3) order of the rotations: doubling along the cyclic successor until no group splits
    bwts_init_keys_kernel<<<G, T, 0, stream>>>(d_src, aux, n, keys_a, idx_a, succ_a);
    (*launches)++;
    end_bit = 63;
    uint32_t prev_groups = 0;
    for (int round = 0; round < 64; round++) {
        cudaError_t e = bwt_sort_round(keys_a, keys_b, idx_a, idx_b, n, end_bit, temp, temp_bytes, head, block_max, nscan, rank, flags, &sa, true, stream, launches);
        if (e != cudaSuccess) return e;
        uint32_t fl[2] = {0, 0};
        e = cudaMemcpyAsync(fl, flags, 8, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (!fl[0] || fl[1] == prev_groups) {  // all distinct, or only equal rotations of repeated factors are left
            bwts_emit_kernel<<<G, T, 0, stream>>>(d_src, sa, pred, n, d_dst);
            (*launches)++;
            return cudaGetLastError();
        }
        prev_groups = fl[1];
        bwts_double_keys_kernel<<<G, T, 0, stream>>>(rank, succ_a, n, log_n, keys_a, idx_a, succ_b);
        (*launches)++;
        std::swap(succ_a, succ_b);
        end_bit = 2 * (int)log_n;
    }
    return cudaErrorUnknown;
}

// BWTS.Inverse of one block
cudaError_t bwts_inverse_device(const uint8_t* d_src, uint32_t n, uint8_t* d_dst, uint8_t* ws, size_t ws_bytes, cudaStream_t stream, uint64_t* launches) {
    if (n < 2) return n == 1 ? cudaMemcpyAsync(d_dst, d_src, 1, cudaMemcpyDeviceToDevice, stream) : cudaSuccess;
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    uint8_t* keys_a = ws + off; off = align(off + n);
    uint8_t* keys_b = ws + off; off = align(off + n);
    uint32_t* val_a = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* val_b = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* lf = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* a0 = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* a1 = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* b0 = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    uint32_t* b1 = (uint32_t*)(ws + off); off = align(off + (size_t)n * 4);
    const uint32_t nscan = (n + 2047) / 2048;
    uint32_t* block_sum = (uint32_t*)(ws + off); off = align(off + (size_t)nscan * 4);
    void* temp = ws + off;
    const size_t temp_bytes = ws_bytes > off ? ws_bytes - off : 0;
    const uint32_t T = 256, G = (n + T - 1) / T;
    ibwts_init_kernel<<<G, T, 0, stream>>>(d_src, n, keys_a, val_a);
    cub::DoubleBuffer<uint8_t> k(keys_a, keys_b);
    cub::DoubleBuffer<uint32_t> v(val_a, val_b);
    size_t tb = temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tb, k, v, (int)n, 0, 8, stream);  // stable: rank of i among equal bytes keeps the order
    if (e != cudaSuccess) return e;
    ibwts_lf_kernel<<<G, T, 0, stream>>>(v.Current(), n, lf);
    (*launches) += 10;
    uint32_t rounds = 1;
    while ((1ull << rounds) < n) rounds++;
    // cycle leaders: minimum index over the whole cycle. val_a / val_b are free again after the scatter
    uint32_t *mn = a0, *jp = a1, *mn2 = b0, *jp2 = b1;
    ibwts_min_init_kernel<<<G, T, 0, stream>>>(lf, n, mn, jp);
    for (uint32_t r = 0; r < rounds; r++) {
        ibwts_min_jump_kernel<<<G, T, 0, stream>>>(mn, jp, n, mn2, jp2);
        std::swap(mn, mn2);
        std::swap(jp, jp2);
    }
    (*launches) += rounds + 1;
    // distance to the end of the cut cycle: mn must stay intact -> use val_a/val_b and the free pair
    uint32_t* nx = jp;       // jump pointers are no longer needed
    uint32_t* cnt = mn2;
    uint32_t* nx2 = jp2;
    uint32_t* cnt2 = val_a;
    ibwts_rank_init_kernel<<<G, T, 0, stream>>>(lf, mn, n, nx, cnt);
    for (uint32_t r = 0; r < rounds; r++) {
        ibwts_rank_jump_kernel<<<G, T, 0, stream>>>(nx, cnt, n, nx2, cnt2);
        std::swap(nx, nx2);
        std::swap(cnt, cnt2);
    }
    (*launches) += rounds + 1;
    uint32_t* len_at = val_b;
    ibwts_len_kernel<<<G, T, 0, stream>>>(mn, cnt, n, len_at);
    scan_add_local_kernel<<<nscan, 256, 0, stream>>>(len_at, n, block_sum);
    scan_add_blocks_kernel<<<1, 32, 0, stream>>>(block_sum, nscan);
    ibwts_write_kernel<<<G, T, 0, stream>>>(d_src, mn, cnt, len_at, block_sum, n, d_dst);
    (*launches) += 4;
    return cudaGetLastError();
}

}  // namespace kz
