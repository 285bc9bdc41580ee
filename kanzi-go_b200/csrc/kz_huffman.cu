// Canonical length-limited Huffman (kanzi "HUFFMAN", bitstream v6) on sm_100a.
//
// Reference semantics reproduced bit-exactly (paths relative to the reference's v2/ directory):
//   entropy/HuffmanCodec.go:390-433  Write (16 KiB chunks, chunks < 32 bytes raw, 1-symbol chunks header only)
//   entropy/HuffmanCodec.go:128-214  updateFrequencies (alphabet, code lengths, canonical codes, Exp-Golomb header)
//   entropy/HuffmanCodec.go:300-385  computeCodeLengths (sort keys (freq<<8)|sym, Moffat-Katajainen in-place phases)
//   entropy/HuffmanCodec.go:216-297  limitCodeLengths (12 bit limit: debt repayment, renormalise-and-retry fallback)
//   entropy/HuffmanCodec.go:37-77    generateCanonicalCodes
//   entropy/ExpGolombCodec.go:104-112, :159-187 signed Exp-Golomb of the code length deltas
//   entropy/HuffmanCodec.go:435-511  encodeChunk (4 fragments, bit lengths as varints, fragments back to back)
//   entropy/HuffmanCodec.go:620-697  readLengths / buildDecodingTable, :807-969 decodeChunkV6
//
// Parallel decomposition: statistics, code construction and bit packing use one warp per chunk (the code construction
// is a few hundred serial steps on <= 256 symbols and runs on lane 0 over shared memory; histogram, key ranking, header
// and the bit packing of the four fragments are warp parallel). Decoding uses four lanes per chunk, one per fragment
// (the fragments are independent bit strings), with a two-level table: 10-bit direct lookup + compare for 11/12-bit codes.
#include "kz_huffman.cuh"

#include "kz_warp.cuh"

namespace kz {

namespace {

// signed Exp-Golomb code of an int8 delta (ExpGolombCodec.go:45-62 table semantics): returns bits, *len = length
KZ_D uint32_t expg_signed(int v, uint32_t* len) {
    if (v == 0) {
        *len = 1;
        return 1;
    }
    const uint32_t m = (uint32_t)(v < 0 ? -v : v);
    const uint32_t k = 31u - (uint32_t)__clz((int)(m + 1));
    *len = 2 * k + 2;
    return ((1u << (k + 1)) | ((m + 1 - (1u << k)) << 1) | (v < 0 ? 1u : 0u)) & 0x1FFu;
}

// HuffmanCodec.go:326-385 on lane 0. data[0..n) = frequencies sorted ascending; on return code lengths. Returns max length.
__device__ int inplace_code_lengths(int* data, int n) {
    for (int s = 0, r = 0, t = 0; t < n - 1; t++) {  // phase 1
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
    if (n < 2) return 0;
    int level_top = n - 2, depth = 1, i = n, total_nodes = 2;  // phase 2
    while (i > 0) {
        int k = level_top;
        while (k > 0 && data[k - 1] >= level_top) k--;
        const int internal_nodes = level_top - k;
        const int leaves = total_nodes - internal_nodes;
        for (int j = 0; j < leaves; j++) data[--i] = depth;
        total_nodes = internal_nodes << 1;
        level_top = k;
        depth++;
    }
    return depth - 1;
}

}  // namespace

// one warp per chunk: histogram -> code lengths -> canonical codes -> header bits
__global__ void __launch_bounds__(HUF_STATS_WARPS * 32) huf_stats_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                                          uint16_t* __restrict__ codes_out, uint8_t* __restrict__ hdr,
                                                                          uint32_t* __restrict__ hdr_bits, uint32_t* __restrict__ asz_out) {
    __shared__ uint32_t s_hist[HUF_STATS_WARPS][8 * 256];
    __shared__ uint32_t s_hdr[HUF_STATS_WARPS][HUF_HDR_STRIDE / 4];
    __shared__ int s_key[HUF_STATS_WARPS][256];    // sort keys, then sorted keys
    __shared__ int s_sorted[HUF_STATS_WARPS][256];
    __shared__ int s_freq[HUF_STATS_WARPS][256];   // frequencies by symbol (fallback path)
    __shared__ uint8_t s_size[HUF_STATS_WARPS][256];
    __shared__ uint16_t s_code[HUF_STATS_WARPS][256];
    __shared__ int s_misc[HUF_STATS_WARPS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * HUF_STATS_WARPS + warp;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    const uint32_t len = ck.len;
    uint32_t* hw = s_hdr[warp];
    for (int i = lane; i < HUF_HDR_STRIDE / 4; i += 32) hw[i] = 0;
    uint32_t f[8];
    warp_histogram(src, len, s_hist[warp], lane, f);
    // alphabet ranks (symbol order)
    int count = 0;
    int rank[8];
    uint32_t last_sym = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        rank[j] = count + __popc(b & ((1u << lane) - 1u));
        if (b) last_sym = 32u * j + (31u - (uint32_t)__clz((int)b));
        count += __popc(b);
        s_size[warp][lane + 32 * j] = 0;
        s_code[warp][lane + 32 * j] = 0;
        s_freq[warp][lane + 32 * j] = (int)f[j];
        if (f[j] != 0) s_key[warp][rank[j]] = (int)((f[j] << 8) | (uint32_t)(lane + 32 * j));
    }
    __syncwarp();
    if (count == 1) {
        if (lane == 0) {  // :156-158
            s_size[warp][last_sym] = 1;
            s_code[warp][last_sym] = 0;
        }
    } else if (count > 1) {
        // sort.Ints on unique keys (:303): position = number of smaller keys
        for (int base = 0; base < count; base += 32) {
            const int i = base + lane;
            if (i < count) {
                const int key = s_key[warp][i];
                int pos = 0;
                for (int j = 0; j < count; j++) pos += s_key[warp][j] < key;
                s_sorted[warp][pos] = key;
            }
        }
        __syncwarp();
        if (lane == 0) {
            int* ranks = s_sorted[warp];   // keys sorted ascending -> symbols (ranks[i] &= 0xFF)
            int* work = s_key[warp];       // frequencies -> lengths
            uint8_t* sizes = s_size[warp];
            for (int i = 0; i < count; i++) {
                work[i] = ranks[i] >> 8;
                ranks[i] &= 0xFF;
            }
            int max_len = inplace_code_lengths(work, count);
            for (int i = 0; i < count; i++) sizes[ranks[i]] = (uint8_t)work[i];
            if (max_len > HUF_MAX_BITS) {  // limitCodeLengths :216-297
                int n = 0, debt = 0;
                while (sizes[ranks[n]] >= HUF_MAX_BITS) {
                    debt += (int)sizes[ranks[n]] - HUF_MAX_BITS;
                    sizes[ranks[n]] = HUF_MAX_BITS;
                    n++;
                }
                // six FIFO queues of symbols, one per size delta; ranks[] is ordered by decreasing size so every queue is a
                // contiguous run of ranks[]: keep (start, end) cursors instead of slices
                int qs[6], qe[6];
                for (int i = 0; i < 6; i++) qs[i] = qe[i] = 0;
                bool opened[6] = {false, false, false, false, false, false};
                while (n < count) {
                    const int idx = HUF_MAX_BITS - 1 - (int)sizes[ranks[n]];
                    if (idx > 5 || debt < (1 << idx)) break;
                    if (!opened[idx]) {
                        opened[idx] = true;
                        qs[idx] = n;
                    }
                    qe[idx] = n + 1;
                    n++;
                }
                int idx = 5;
                while (debt > 0 && idx >= 0) {
                    if (qs[idx] == qe[idx] || debt < (1 << idx)) {
                        idx--;
                        continue;
                    }
                    sizes[ranks[qs[idx]++]]++;
                    debt -= (1 << idx);
                }
                idx = 0;
                while (debt > 0 && idx < 6) {
                    if (qs[idx] == qe[idx]) {
                        idx++;
                        continue;
                    }
                    sizes[ranks[qs[idx]++]]++;
                    debt -= (1 << idx);
                }
                max_len = HUF_MAX_BITS;
                if (debt > 0) max_len = -1;  // slow path needed (handled below by the whole warp)
            }
            s_misc[warp][0] = max_len;
        }
        __syncwarp();
        if (s_misc[warp][0] == -1) {
            // :272-294 renormalise the frequencies to 2048 (NormalizeFrequencies over the count alphabet entries, in
            // alphabet order) and rebuild. Extremely rare (needs > 12 bit codes that the debt heuristic cannot fix);
            // done serially on lane 0 with the scalar restatement.
            if (lane == 0) {
                int* fr = s_key[warp];  // f[i] = freqs[symbols[i]] in alphabet order
                int n = 0, total = 0;
                for (int s = 0; s < 256; s++)
                    if (s_freq[warp][s] > 0) {
                        fr[n] = s_freq[warp][s];
                        total += fr[n];
                        n++;
                    }
                const int scale = HUF_CHUNK >> 3;
                // scalar NormalizeFrequencies (EntropyUtils.go:123-260) over fr[0..n)
                if (total != scale) {
                    int sum_scaled = 0, sum = 0, idx_max = 0, asz = 0;
                    int* alpha = s_sorted[warp];
                    for (int i = 0; i < n; i++) {
                        const int fv = fr[i];
                        if (fv == 0) continue;
                        const long long sf = (long long)fv * scale;
                        const int sc = sf <= total ? 1 : (int)((sf + (total >> 1)) / total);
                        alpha[asz++] = i;
                        sum_scaled += sc;
                        fr[i] = sc;
                        sum += fv;
                        if (sc > fr[idx_max]) idx_max = i;
                        if (sum >= total) break;
                    }
                    if (asz == 1) {
                        fr[alpha[0]] = scale;
                    } else if (asz > 1 && sum_scaled != scale) {
                        int delta = sum_scaled - scale;
                        const int err_thr = fr[idx_max] >> 4;
                        const int abs_delta = delta < 0 ? -delta : delta;
                        if (abs_delta <= err_thr) {
                            fr[idx_max] -= delta;
                        } else {
                            int inc;
                            if (delta < 0) {
                                delta += err_thr;
                                fr[idx_max] += err_thr;
                                inc = 1;
                                delta = -delta;
                            } else {
                                delta -= err_thr;
                                fr[idx_max] -= err_thr;
                                inc = -1;
                            }
                            for (int round = 1; round < 6 && delta > 0; round++) {
                                int adjustments = 0;
                                for (int k2 = 0; k2 < asz; k2++) {
                                    const int id = alpha[k2];
                                    if (fr[id] <= 2) continue;
                                    fr[id] += inc;
                                    adjustments++;
                                    delta--;
                                    if (delta == 0) break;
                                }
                                if (adjustments == 0) break;
                            }
                            fr[idx_max] = max(fr[idx_max] - delta, 1);
                        }
                    }
                }
                // ranks[i] = (f[i] << 8) | symbols[i]; sort; lengths
                int* ranks = s_sorted[warp];
                int m = 0;
                for (int s = 0; s < 256; s++)
                    if (s_freq[warp][s] > 0) {
                        ranks[m] = (fr[m] << 8) | s;
                        m++;
                    }
                for (int i = 1; i < m; i++) {  // insertion sort (rare path)
                    const int key = ranks[i];
                    int j = i - 1;
                    while (j >= 0 && ranks[j] > key) {
                        ranks[j + 1] = ranks[j];
                        j--;
                    }
                    ranks[j + 1] = key;
                }
                int* work = s_key[warp];
                for (int i = 0; i < m; i++) {
                    work[i] = ranks[i] >> 8;
                    ranks[i] &= 0xFF;
                }
                const int max_len = inplace_code_lengths(work, m);
                for (int i = 0; i < m; i++) s_size[warp][ranks[i]] = (uint8_t)work[i];
                s_misc[warp][0] = max_len;
            }
            __syncwarp();
        }
        if (lane == 0) {
            uint8_t* sizes = s_size[warp];
            uint16_t* codes = s_code[warp];
            if (s_misc[warp][0] > HUF_MAX_BITS) {  // :181-186 unlikely: 8-bit codes = alphabet index
                int i = 0;
                for (int s = 0; s < 256; s++)
                    if (s_freq[warp][s] > 0) {
                        codes[s] = (uint16_t)i++;
                        sizes[s] = 8;
                    }
            } else {  // generateCanonicalCodes :37-77: symbols ordered by (size, symbol)
                int cnt[HUF_MAX_BITS + 2];
                for (int l = 0; l <= HUF_MAX_BITS + 1; l++) cnt[l] = 0;
                int min_len = 99;
                for (int s = 0; s < 256; s++)
                    if (sizes[s]) {
                        cnt[sizes[s]]++;
                        if (sizes[s] < min_len) min_len = sizes[s];
                    }
                int next[HUF_MAX_BITS + 2];
                int code = 0, cur = min_len;
                for (int l = min_len; l <= HUF_MAX_BITS; l++) {
                    if (cnt[l] == 0) continue;
                    code <<= (l - cur);
                    cur = l;
                    next[l] = code;
                    code += cnt[l];
                }
                for (int s = 0; s < 256; s++)
                    if (sizes[s]) codes[s] = (uint16_t)next[sizes[s]]++;
            }
        }
    }
    __syncwarp();
    // ---- header: alphabet (EntropyUtils.go:38-67) + signed Exp-Golomb of size deltas in symbol order (:201-210)
    uint32_t pos = 0;
    if (count == 256) {
        pos = 2;
    } else if (count == 0) {
        if (lane == 0) smem_put_bits(hw, 1, 1, 1);
        pos = 2;
    } else {
        const uint32_t last_mask = last_sym >> 3;
        if (lane == 0) {
            smem_put_bits(hw, 0, 1, 1);
            smem_put_bits(hw, 1, last_mask, 5);
        }
        pos = 6;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (f[j] != 0) {
                const uint32_t s = (uint32_t)lane + 32u * j;
                smem_put_bits(hw, pos + 8 * (s >> 3) + (7 - (s & 7)), 1, 1);
            }
        pos += 8 * (last_mask + 1);
    }
    if (count > 0) {
        // previous present symbol's size for every present symbol: alphabet-order neighbour
        __shared__ uint8_t s_alpha_size[HUF_STATS_WARPS][256];
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (f[j] != 0) s_alpha_size[warp][rank[j]] = s_size[warp][lane + 32 * j];
        __syncwarp();
        uint32_t carry = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t bits = 0, blen = 0;
            if (f[j] != 0) {
                const int cur = s_alpha_size[warp][rank[j]];
                const int prev = rank[j] == 0 ? 2 : s_alpha_size[warp][rank[j] - 1];
                bits = expg_signed((int)(int8_t)(uint8_t)(cur - prev), &blen);
            }
            const uint32_t incl = warp_incl_scan(blen, lane);
            if (blen) smem_put_bits(hw, pos + carry + incl - blen, bits, blen);
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        pos += carry;
    }
    __syncwarp();
    uint32_t* gh = reinterpret_cast<uint32_t*>(hdr + (size_t)c * HUF_HDR_STRIDE);
    for (int i = lane; i < HUF_HDR_STRIDE / 4; i += 32) gh[i] = bswap32(hw[i]);
    uint16_t* gc = codes_out + (size_t)c * 256;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int s = lane + 32 * j;
        gc[s] = (uint16_t)(s_code[warp][s] | ((uint16_t)s_size[warp][s] << 12));
    }
    if (lane == 0) {
        hdr_bits[c] = pos;
        asz_out[c] = (uint32_t)count;
    }
}

// one warp per chunk: packs the 4 fragments (MSB-first concatenation of the codes, :442-491)
__global__ void __launch_bounds__(HUF_ENC_WARPS * 32) huf_encode_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                                         const uint16_t* __restrict__ codes_in, uint8_t* __restrict__ hdr,
                                                                         const uint32_t* __restrict__ hdr_bits, const uint32_t* __restrict__ asz_in,
                                                                         uint8_t* __restrict__ frag, uint64_t* __restrict__ seg_src,
                                                                         uint64_t* __restrict__ seg_bits, const uint32_t* __restrict__ seg_index) {
    __shared__ uint32_t s_buf[HUF_ENC_WARPS][HUF_FRAG_WORDS];
    __shared__ uint16_t s_codes[HUF_ENC_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * HUF_ENC_WARPS + warp;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    const uint32_t count = ck.len;
    const uint32_t si = seg_index[c];
    uint8_t* my_hdr = hdr + (size_t)c * HUF_HDR_STRIDE;
    const uint32_t hb = hdr_bits[c];
    if (asz_in[c] <= 1) {  // :423-426 header only
        if (lane == 0) {
            seg_src[si] = (uint64_t)(uintptr_t)my_hdr;
            seg_bits[si] = hb;
            for (int j = 1; j < 6; j++) {
                seg_src[si + j] = 0;
                seg_bits[si + j] = 0;
            }
        }
        return;
    }
    for (int i = lane; i < 256; i += 32) s_codes[warp][i] = codes_in[(size_t)c * 256 + i];
    __syncwarp();
    const uint32_t sz_frag = count / 4;
    uint32_t nb_bits[4];
    uint32_t* buf = s_buf[warp];
    uint8_t* my_frag = frag + (size_t)c * HUF_FRAG_STRIDE * 4;
    for (int j = 0; j < 4; j++) {
        for (int i = lane; i < HUF_FRAG_WORDS; i += 32) buf[i] = 0;
        __syncwarp();
        const uint8_t* p = src + (size_t)j * sz_frag;
        uint32_t base_bits = 0;
        // tiles of 32 lanes x 8 symbols
        for (uint32_t t0 = 0; t0 < sz_frag; t0 += 256) {
            const uint32_t i0 = t0 + 8u * lane;
            uint64_t acc = 0;  // up to 96 bits: keep two pieces
            uint32_t acc_hi = 0, nbits = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (i0 + q < sz_frag) {
                    const uint32_t cd = s_codes[warp][p[i0 + q]];
                    const uint32_t l = cd >> 12;
                    // (acc_hi:acc) = (acc_hi:acc) << l | code
                    acc_hi = (acc_hi << l) | (uint32_t)(acc >> (64 - l));
                    acc = (acc << l) | (cd & 0x0FFF);
                    nbits += l;
                }
            }
            const uint32_t incl = warp_incl_scan(nbits, lane);
            uint32_t at = base_bits + incl - nbits;
            // emit nbits (<= 96) bits: top (nbits - 64) from acc_hi, then 64 from acc, MSB first
            if (nbits > 64) {
                smem_put_bits(buf, at, acc_hi, nbits - 64);
                at += nbits - 64;
                smem_put_bits(buf, at, (uint32_t)(acc >> 32), 32);
                smem_put_bits(buf, at + 32, (uint32_t)acc, 32);
            } else if (nbits > 32) {
                smem_put_bits(buf, at, (uint32_t)(acc >> 32), nbits - 32);
                smem_put_bits(buf, at + nbits - 32, (uint32_t)acc, 32);
            } else if (nbits > 0) {
                smem_put_bits(buf, at, (uint32_t)acc, nbits);
            }
            base_bits += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        __syncwarp();
        nb_bits[j] = base_bits;
        uint32_t* gf = reinterpret_cast<uint32_t*>(my_frag + (size_t)j * HUF_FRAG_STRIDE);
        const uint32_t nwords = (base_bits + 31) >> 5;
        for (uint32_t i = lane; i < nwords; i += 32) gf[i] = bswap32(buf[i]);
        __syncwarp();
    }
    if (lane == 0) {
        BitWriter bw(my_hdr, hb);
        for (int j = 0; j < 4; j++) bw.write_varint(nb_bits[j]);  // :494-497
        seg_src[si] = (uint64_t)(uintptr_t)my_hdr;
        seg_bits[si] = bw.pos;
        for (int j = 0; j < 4; j++) {
            seg_src[si + 1 + j] = (uint64_t)(uintptr_t)(my_frag + (size_t)j * HUF_FRAG_STRIDE);
            seg_bits[si + 1 + j] = nb_bits[j];
        }
        seg_src[si + 5] = (uint64_t)(uintptr_t)(src + 4 * (size_t)sz_frag);  // :505-510 chunk last bytes
        seg_bits[si + 5] = 8ull * (count - 4 * sz_frag);
    }
}

// ------------------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------------------
namespace {

// readLengths (:620-658) from a bit reader: alphabet + code lengths. sizes[] indexed by alphabet rank. Returns count or -1.
// MSB-first reader over a window of the stream staged in shared memory (words already byte swapped): same interface as
// BitReader, no global memory traffic. `base` = absolute bit position of win[0]; reads past the window return zeros.
struct WinBitReader {
    const uint32_t* win;
    uint32_t nwords;
    uint64_t base, pos, end;
    bool overrun;
    KZ_D WinBitReader(const uint32_t* w, uint32_t n, uint64_t b, uint64_t p, uint64_t e) : win(w), nwords(n), base(b), pos(p), end(e), overrun(false) {}
    KZ_D uint32_t read(uint32_t n) {  // n in 1..32
        if (pos + n > end) {
            overrun = true;
            pos += n;
            return 0;
        }
        const uint64_t rel = pos - base;
        const uint32_t w = (uint32_t)(rel >> 5), o = (uint32_t)rel & 31u;
        pos += n;
        if (w + 1 >= nwords) {
            overrun = true;
            return 0;
        }
        return __funnelshift_l(win[w + 1], win[w], o) >> (32 - n);
    }
    KZ_D uint32_t read_varint() {  // entropy/EntropyUtils.go:278-296
        uint32_t res = 0, shift = 0;
        for (int i = 0; i < 4; i++) {
            const uint32_t v = read(8);
            res |= (v & 0x7F) << shift;
            if (v < 128) return res;
            shift += 7;
        }
        return res | ((read(8) & 0x0F) << 28);
    }
};

template <class Reader>
KZ_D int read_lengths(Reader& br, uint8_t* alphabet /*smem, count entries*/, uint8_t* sizes /*by rank*/) {
    int count = 0;
    if (br.read(1) == 0) {
        if (br.read(1) == 0) {
            count = 256;
            for (int i = 0; i < 256; i++) alphabet[i] = (uint8_t)i;
        }
    } else {
        const uint32_t last = br.read(5);
        for (uint32_t i = 0; i <= last; i++) {
            uint32_t m = br.read(8);
            while (m) {
                const int j = __ffs((int)m) - 1;
                m &= m - 1;
                alphabet[count++] = (uint8_t)(8 * i + j);
            }
        }
    }
    if (count == 0) return 0;
    int cur = 2;
    for (int i = 0; i < count; i++) {
        // ExpGolombCodec.go:159-187 (signed)
        int delta = 0;
        if (br.read(1) == 0) {
            uint32_t lg = 1;
            while (br.read(1) != 1) {
                lg++;
                if (br.overrun) return -1;
            }
            lg &= 7;
            const uint32_t val = br.read(lg + 1);
            int res = (int)((val >> 1) + (1u << lg) - 1u);
            if (val & 1) res = -res;
            delta = (int)(int8_t)(uint8_t)res;
        }
        cur = (int)(int8_t)(cur + delta);
        if (cur <= 0 || cur > HUF_MAX_BITS) return -1;
        sizes[i] = (uint8_t)cur;
    }
    return count;
}

}  // namespace

// One WARP per block: skims the chunk headers to find where every chunk starts (serial by format: a chunk's position is
// known only after the previous header and its four fragment sizes are parsed). v1 walked the <= 256 Exp-Golomb codes of
// every header with two dependent global loads per bit field (18.9 ms per 64 x 4 MiB); here the header window
// (<= 545 + 20 bytes) is staged in shared memory with one coalesced load, the alphabet is counted by popc + REDUX and the
// code lengths are skimmed with one count-leading-zeros per code (ExpGolombCodec.go:159-187: '1' = 0, otherwise z zeros,
// a one, (z & 7) + 1 bits), several codes per 32-bit window.
static const int HUF_WIN_WORDS = 176;

__global__ void __launch_bounds__(32) huf_walk_kernel(const uint32_t* __restrict__ words, uint64_t words_len, const DecBlock* __restrict__ blocks,
                                                      int nblocks, uint64_t* __restrict__ chunk_hdr_bit, uint64_t* __restrict__ blk_end,
                                                      int32_t* __restrict__ status) {
    __shared__ uint32_t s_win[HUF_WIN_WORDS + 2];
    __shared__ uint8_t s_lut[256];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const DecBlock blk = blocks[b];
    if (blk.nchunks == 0) return;
    for (int v = lane; v < 256; v += 32) {  // whole 1-bit / 4-bit codes at the top of the byte v
        uint32_t x = (uint32_t)v << 24, nb = 0, nc = 0;
        for (;;) {
            uint32_t len;
            if (x & 0x80000000u) len = 1;
            else if (x & 0x40000000u) len = 4;
            else break;
            if (nb + len > 8) break;
            nb += len;
            nc++;
            x <<= len;
        }
        s_lut[v] = (uint8_t)((nc << 4) | nb);
    }
    __syncwarp();
    uint64_t pos = blk.data_bit;
    uint32_t remaining = blk.pre_len;
    int32_t st = 0;
    auto rd = [&](uint64_t at, uint32_t n) -> uint32_t {  // n in 1..32 bits at window relative bit `at`; zeros past the window
        const uint64_t w = at >> 5;
        if (w >= (uint64_t)HUF_WIN_WORDS) return 0u;
        return __funnelshift_l(s_win[w + 1], s_win[w], (uint32_t)at & 31u) >> (32 - n);
    };
    for (uint32_t k = 0; k < blk.nchunks; k++) {
        if (lane == 0) chunk_hdr_bit[blk.chunk_base + k] = pos;
        if (st) continue;  // descriptors of the remaining chunks point at the failure position; decode is skipped
        const uint32_t clen = remaining < (uint32_t)HUF_CHUNK ? remaining : (uint32_t)HUF_CHUNK;
        remaining -= clen;
        if (clen < 32) {  // :768-770 raw
            pos += 8ull * clen;
        } else {
            const uint64_t w0 = pos >> 5;
            __syncwarp();
            for (int i = lane; i < HUF_WIN_WORDS + 2; i += 32) s_win[i] = w0 + i < words_len ? bswap32(__ldg(words + w0 + i)) : 0u;
            __syncwarp();
            uint64_t q = pos & 31;
            int count = 0;
            if (rd(q, 1) == 0) {
                count = rd(q + 1, 1) ? 0 : 256;
                q += 2;
            } else {
                const uint32_t last = rd(q + 1, 5);
                q += 6;
                const uint32_t cnt = (uint32_t)lane <= last ? (uint32_t)__popc(rd(q + 8 * lane, 8)) : 0u;
                count = (int)__reduce_add_sync(0xFFFFFFFFu, cnt);
                q += 8 * (last + 1);
            }
            if (count == 0) {
                st = -KZ_E_PROCESS_BLOCK;
            } else {
                // skim the Exp-Golomb coded size deltas. Codes are 1 bit ('1'), 4 bits ('01vv'), 7 bits ('001vvvv') ... long:
                // s_lut[byte] = (codes that end inside these 8 bits) << 4 | bits they occupy, so a typical header (mostly 1 and 4
                // bit codes) advances several codes per look-up; longer codes take the count-leading-zeros path.
                int i = 0;
                while (i < count) {
                    uint32_t win = rd(q, 32);
                    uint32_t used = 0;
                    do {
                        const uint32_t e = s_lut[win >> 24];
                        const uint32_t nc = e >> 4;
                        if (nc != 0 && i + (int)nc <= count) {
                            const uint32_t nb = e & 15u;
                            used += nb;
                            win <<= nb;
                            i += (int)nc;
                            continue;
                        }
                        uint32_t len = 1;
                        if ((int32_t)win >= 0) {  // leading bit 0
                            const uint32_t z = (uint32_t)__clz((int)win);  // 1..32
                            len = z + 1 + (z & 7) + 1;
                        }
                        if (len > 32 - used) {
                            if (used == 0) {  // longer than a window (corrupt input): advance and fetch again
                                used = len;
                                i++;
                            }
                            break;
                        }
                        used += len;
                        win = len < 32 ? win << len : 0u;
                        i++;
                    } while (i < count && used <= 14);
                    q += used;
                }
                if (count > 1) {
                    uint64_t total = 0;
                    for (int j = 0; j < 4; j++) {  // ReadVarInt (EntropyUtils.go:278-296)
                        uint32_t v = 0, shift = 0;
                        int n = 0;
                        for (; n < 4; n++) {
                            const uint32_t by = rd(q, 8);
                            q += 8;
                            v |= (by & 0x7F) << shift;
                            shift += 7;
                            if (by < 128) break;
                        }
                        if (n == 4) {
                            v |= (rd(q, 8) & 0x0F) << 28;
                            q += 8;
                        }
                        if ((int32_t)v < 0) st = -KZ_E_PROCESS_BLOCK;
                        total += v;
                    }
                    q += total + 8ull * (clen - 4 * (clen / 4));
                }
            }
            pos = (w0 << 5) + q;
        }
        if (pos > blk.end_bit && st == 0) st = -KZ_E_PROCESS_BLOCK;
    }
    if (lane == 0) {
        blk_end[b] = pos;
        if (st) status[b] = st;
    }
}

// Four lanes per chunk (lane j decodes fragment j), 8 chunks per warp, HUF_DEC_CTA_CHUNKS chunks per CTA.
// Table per chunk: 1024 x u16 on the top 10 bits of the 12-bit window: (sym << 8) | len for codes of <= 10 bits,
// 0xFFFF marks an 11/12-bit code which is resolved canonically (first code / count / symbol base per length).
__global__ void __launch_bounds__(HUF_DEC_CTA_CHUNKS * 4) huf_decode_kernel(const uint32_t* __restrict__ words, uint64_t words_len,
                                                                             const DecChunk* __restrict__ chunks, int nchunks,
                                                                             const uint64_t* __restrict__ chunk_hdr_bit, uint8_t* __restrict__ out,
                                                                             int32_t* __restrict__ status) {
    extern __shared__ uint8_t s_huf[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int slot = tid >> 2, k = tid & 3;
    uint8_t* base = s_huf + (size_t)slot * HUF_DEC_CHUNK_BYTES;
    uint16_t* table = reinterpret_cast<uint16_t*>(base);           // 1024 entries
    uint8_t* alphabet = base + 2048;                                // 256: symbols, then symbols sorted by (size, sym)
    uint8_t* sizes = base + 2048 + 256;                             // 256: size by rank
    uint32_t* info = reinterpret_cast<uint32_t*>(base + 2048 + 512);  // [0..3] fragment bit offsets (lo), [4] mode, [5..8] long-code data
    const int c = blockIdx.x * HUF_DEC_CTA_CHUNKS + slot;
    uint64_t frag_bit = 0;
    uint32_t frag_bits = 0;
    uint32_t sz_frag = 0;
    uint8_t* dst = nullptr;
    int mode = 0;
    // long codes: for len 11 and 12: first code (left aligned to 12 bits), count, base index into the sorted symbols
    uint32_t lim11 = 0, first11 = 0, base11 = 0, first12 = 0, base12 = 0;
    // the chunk header (<= 545 + 20 bytes) is staged by the four lanes of the chunk in the (not yet built) table area: the
    // serial parse below then runs on shared memory (v1 walked it with two dependent global loads per bit field)
    constexpr uint32_t HWIN = 148;
    uint32_t* hwin = reinterpret_cast<uint32_t*>(table);
    uint64_t hbit = 0;
    if (c < nchunks) {
        hbit = chunk_hdr_bit[c];
        const uint64_t w0 = hbit >> 5;
        for (uint32_t j = (uint32_t)k; j < HWIN; j += 4) hwin[j] = w0 + j < words_len ? bswap32(__ldg(words + w0 + j)) : 0u;
    }
    __syncwarp();
    if (c < nchunks) {
        const DecChunk ck = chunks[c];
        dst = out + ck.out_off;
        if (k == 0 && status[ck.block] == 0) {
            WinBitReader br(hwin, HWIN, (hbit >> 5) << 5, hbit, ck.end_bit);
            if (ck.out_len < 32) {  // raw chunk (:768-770)
                for (uint32_t i = 0; i < ck.out_len; i++) dst[i] = (uint8_t)br.read(8);
                if (br.overrun) atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
            } else {
                const int count = read_lengths(br, alphabet, sizes);
                if (count <= 0 || br.overrun) {
                    atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
                } else if (count == 1) {
                    const uint8_t v = alphabet[0];
                    for (uint32_t i = 0; i < ck.out_len; i++) dst[i] = v;
                } else {
                    // canonical codes (:37-77): order (size, symbol); the alphabet is in symbol order already
                    uint32_t cnt[HUF_MAX_BITS + 2];
                    for (int l = 0; l <= HUF_MAX_BITS + 1; l++) cnt[l] = 0;
                    for (int i = 0; i < count; i++) cnt[sizes[i]]++;
                    uint32_t first[HUF_MAX_BITS + 2], sbase[HUF_MAX_BITS + 2];
                    uint32_t code = 0, cur = 0, idx = 0;
                    bool started = false, bad = false;
                    for (int l = 1; l <= HUF_MAX_BITS; l++) {
                        first[l] = 0;
                        sbase[l] = idx;
                        if (cnt[l] == 0) continue;
                        if (!started) {
                            started = true;
                            cur = l;
                        }
                        code <<= (l - cur);
                        cur = l;
                        first[l] = code;
                        code += cnt[l];
                        idx += cnt[l];
                        if (code > (1u << l)) bad = true;  // buildDecodingTable :683-685
                    }
                    // the four fragment sizes follow the code lengths: read them before the table overwrites the staged header
                    uint32_t szb[4];
                    for (int j = 0; j < 4; j++) {
                        szb[j] = br.read_varint();
                        if ((int32_t)szb[j] < 0) bad = true;
                    }
                    const uint64_t fb = br.pos;
                    if (br.overrun) bad = true;
                    if (bad) {
                        atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
                    } else {
                        for (int i = 0; i < 1024; i++) table[i] = 0x0007;  // :665-667 default: consume 7 bits, symbol 0
                        // direct entries for lengths <= 10, sorted symbol list for all
                        uint32_t next[HUF_MAX_BITS + 2];
                        for (int l = 1; l <= HUF_MAX_BITS; l++) next[l] = first[l];
                        uint8_t* sorted = alphabet;  // rewritten in place below via a temporary pass over sizes
                        // place symbols: need (size, sym) order; the alphabet is ascending so a stable pass per symbol works
                        // with per-length cursors. Symbols are copied to a scratch area first.
                        uint8_t* tmp = reinterpret_cast<uint8_t*>(info) + 64;  // 256 bytes scratch
                        for (int i = 0; i < count; i++) tmp[i] = alphabet[i];
                        uint32_t place[HUF_MAX_BITS + 2];
                        for (int l = 1; l <= HUF_MAX_BITS; l++) place[l] = sbase[l];
                        for (int i = 0; i < count; i++) {
                            const uint32_t l = sizes[i];
                            const uint32_t sym = tmp[i];
                            sorted[place[l]++] = (uint8_t)sym;
                            const uint32_t cd = next[l]++;
                            if (l <= 10) {
                                const uint32_t lo = cd << (10 - l), n = 1u << (10 - l);
                                const uint16_t v = (uint16_t)((sym << 8) | l);
                                for (uint32_t q = 0; q < n; q++) table[lo + q] = v;
                            } else {
                                table[cd >> (l - 10)] = 0xFFFF;
                            }
                        }
                        info[5] = first[11] << 1;                       // left aligned to 12 bits
                        info[6] = (first[11] + cnt[11]) << 1;           // limit of the 11-bit codes (12-bit aligned)
                        info[7] = sbase[11];
                        info[8] = first[12];
                        info[9] = sbase[12];
                        uint64_t tot = (uint64_t)szb[0] + szb[1] + szb[2] + szb[3];
                        if (fb + tot + 8ull * (ck.out_len & 3) > ck.end_bit) {
                            atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
                        } else {
                            info[0] = szb[0];
                            info[1] = szb[1];
                            info[2] = szb[2];
                            info[3] = szb[3];
                            info[10] = (uint32_t)fb;
                            info[11] = (uint32_t)(fb >> 32);
                            info[4] = 1;
                            mode = 1;
                            // chunk last bytes (:961-966)
                            uint64_t tp = fb + tot;
                            const uint32_t c4 = 4 * (ck.out_len / 4);
                            for (uint32_t i = c4; i < ck.out_len; i++, tp += 8) dst[i] = (uint8_t)bits_at(words, tp, 8);
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    const int leader = lane & ~3;
    mode = __shfl_sync(0xFFFFFFFFu, mode, leader);
    if (mode) {
        const DecChunk ck = chunks[c];
        sz_frag = ck.out_len / 4;
        uint64_t fb = ((uint64_t)info[11] << 32) | info[10];
        for (int j = 0; j < k; j++) fb += info[j];
        frag_bit = fb;
        frag_bits = info[k];
        lim11 = info[6];
        first11 = info[5];
        base11 = info[7];
        first12 = info[8];
        base12 = info[9];
    }
    if (!mode) return;
    // ---- fragment decode: sequential table lookups on a zero-extended bit string (:872-959).
    // v1 fetched the 12-bit window of every symbol from global memory (two dependent loads per symbol, 4.2 ms per
    // 64 x 4 MiB). Now the lane keeps a 64-bit shift register of its fragment and pulls 32 bits at a time from a queue of
    // four words loaded well ahead of their use (the addresses do not depend on the decoded symbols): 2.3 ms. A variant
    // with two 16-byte quads in flight measured slower (4.0 ms) and was dropped.
    uint8_t* o = dst + (size_t)k * sz_frag;
    const uint32_t* wp = words + (frag_bit >> 5);
    const uint32_t sh0 = (uint32_t)frag_bit & 31u;
    uint32_t qa = __ldg(wp), qb = __ldg(wp + 1), qc = __ldg(wp + 2), qd = __ldg(wp + 3);
    wp += 4;
    int64_t avail = (int64_t)frag_bits;  // fragment bits not yet moved into `head`
    auto take = [&](uint32_t raw, uint32_t valid) -> uint32_t {  // the next `valid` (<= 32) stream bits, left aligned; zeros past the fragment
        uint32_t w = raw;
        if (avail < (int64_t)valid) w = avail <= 0 ? 0u : (w & ~(0xFFFFFFFFu >> (uint32_t)avail));
        avail -= valid;
        return w;
    };
    uint64_t head = (uint64_t)take(bswap32(qa) << sh0, 32u - sh0) << 32;
    uint32_t nbits = 32u - sh0;
    qa = qb, qb = qc, qc = qd, qd = __ldg(wp++);
    const bool al4 = ((uintptr_t)o & 3) == 0;
    uint32_t pack = 0;
    for (uint32_t i = 0; i < sz_frag; i++) {
        if (nbits <= 32) {
            head |= (uint64_t)take(bswap32(qa), 32) << (32u - nbits);
            nbits += 32;
            qa = qb, qb = qc, qc = qd;
            qd = avail > 96 ? __ldg(wp) : 0u;  // words past the fragment are never needed (and may lie past the buffer)
            wp++;
        }
        const uint32_t w = (uint32_t)(head >> 52);
        const uint32_t e = table[w >> 2];
        uint32_t sym, l;
        if (e == 0xFFFF) {
            if (w < lim11 && lim11 != first11) {
                l = 11;
                sym = alphabet[(base11 + ((w - first11) >> 1)) & 0xFF];
            } else {
                l = 12;
                sym = alphabet[(base12 + (w - first12)) & 0xFF];
            }
        } else {
            sym = e >> 8;
            l = e & 0xFF;
        }
        head <<= l;
        nbits -= l;
        if (al4) {
            pack |= sym << (8 * (i & 3));
            if ((i & 3) == 3) {
                *reinterpret_cast<uint32_t*>(o + i - 3) = pack;
                pack = 0;
            }
        } else {
            o[i] = (uint8_t)sym;
        }
    }
    if (al4)
        for (uint32_t i = sz_frag & ~3u; i < sz_frag; i++) o[i] = (uint8_t)(pack >> (8 * (i & 3)));
}

}  // namespace kz
