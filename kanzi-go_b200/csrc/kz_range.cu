// Order-0 range coder (kanzi "RANGE") on sm_100a.
//
// Reference semantics reproduced bit-exactly (paths relative to the reference's v2/ directory):
//   entropy/RangeCodec.go:223-272  RangeEncoder.Write (32 KiB chunks, lr lowered while 2^lr > chunk length)
//   entropy/RangeCodec.go:137-161  updateFrequencies -> NormalizeFrequencies (EntropyUtils.go:123-260) + cumFreqs
//   entropy/RangeCodec.go:163-217  encodeHeader (alphabet FIRST, then 3 bits lr-8, then the frequency groups)
//   entropy/RangeCodec.go:281-305  encodeByte (28-bit digits written straight into the bit stream, 60-bit flush)
//   entropy/RangeCodec.go:403-487  decodeHeader, :489-557 Read / decodeByte
//
// What the wire format allows. Encode: chunks are independent (rng/low are reset per chunk and every chunk yields
// its own bit string) -> one warp per chunk: histogram, normalisation and header are warp parallel, the arithmetic
// chain of the chunk (one 64-bit multiply chain per byte) is walked by lane 0; 128 chunks per 4 MiB block run
// concurrently. Decode: a chunk carries NO length field, chunk k+1 starts where the decoder of chunk k stops reading,
// so a block is one serial chain of pre_len divisions -> one warp per block (tables built by the warp, chain on
// lane 0). That is a format limit, not an implementation choice; only blocks run in parallel.
#include "kz_range.cuh"
#include "kz_warp.cuh"

namespace kz {

static const uint64_t RC_TOP = 0x0FFFFFFFFFFFFFFFull;     // RangeCodec.go:31
static const uint64_t RC_BOTTOM = 0x000000000000FFFFull;  // :32
static const uint64_t RC_MASK = 0x0FFFFFFF00000000ull;    // :33

// MSB-first bit appender into a big-endian byte string in global memory, 32 bits at a time (single writer thread)
struct WordBitWriter {
    uint32_t* w;     // destination as words (4-byte aligned)
    uint64_t acc;    // pending bits, left aligned in the low `n` bits
    uint32_t n;      // number of pending bits (< 32 after every put)
    uint64_t total;  // bits appended so far
    KZ_D void init(uint32_t* dst) {
        w = dst;
        acc = 0;
        n = 0;
        total = 0;
    }
    KZ_D void put(uint64_t v, uint32_t bits) {  // bits <= 32
        acc = (acc << bits) | (v & ((1ull << bits) - 1ull));
        n += bits;
        total += bits;
        if (n >= 32) {
            n -= 32;
            *w++ = bswap32((uint32_t)(acc >> n));
        }
    }
    KZ_D void flush() {  // zero padded
        if (n) *w++ = bswap32((uint32_t)(acc << (32 - n)));
        n = 0;
    }
};

// grid: one warp per chunk, RC_ENC_WARPS warps per CTA
__global__ void __launch_bounds__(RC_ENC_WARPS * 32) range_encode_kernel(const uint8_t* __restrict__ in, const ChunkIn* __restrict__ chunks, int nchunks,
                                                                         uint8_t* __restrict__ outbuf, uint64_t* __restrict__ seg_src,
                                                                         uint64_t* __restrict__ seg_bits, const uint32_t* __restrict__ seg_index,
                                                                         int32_t* __restrict__ status) {
    __shared__ uint32_t s_hist[RC_ENC_WARPS][8][256];
    __shared__ uint32_t s_hdr[RC_ENC_WARPS][RC_HDR_WORDS];
    __shared__ uint16_t s_rank[RC_ENC_WARPS][256];
    __shared__ uint16_t s_cum[RC_ENC_WARPS][260];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * RC_ENC_WARPS + warp;
    if (c >= nchunks) return;
    const ChunkIn ck = chunks[c];
    const uint8_t* src = in + ck.src_off;
    const uint32_t len = ck.len;
    uint32_t lr = 12;                          // _DEFAULT_RANGE_LOG_RANGE
    while (lr > 8 && (1u << lr) > len) lr--;   // Write :243-245
    for (int i = lane; i < RC_HDR_WORDS; i += 32) s_hdr[warp][i] = 0;
    uint32_t f[8];
    warp_histogram(src, len, &s_hist[warp][0][0], lane, f);
    const int asz = warp_normalize(f, len, lr, lane);
    // cumFreqs (updateFrequencies :150-156) and the frequencies by alphabet rank
    uint32_t carry = 0, last_sym = 0;
    int rank_base = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t incl = warp_incl_scan(f[j], lane);
        s_cum[warp][lane + 32 * j + 1] = (uint16_t)(carry + incl);  // <= 4096
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        const int rank = rank_base + __popc(b & ((1u << lane) - 1u));
        if (f[j] != 0) s_rank[warp][rank] = (uint16_t)(f[j] - 1);
        if (b) last_sym = 32u * j + (31u - (uint32_t)__clz((int)b));
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        rank_base += __popc(b);
    }
    if (lane == 0) s_cum[warp][0] = 0;
    __syncwarp();
    // ---- header: EncodeAlphabet (EntropyUtils.go:38-67), 3 bits lr - 8, frequency groups (encodeHeader :163-217)
    uint32_t* hw = s_hdr[warp];
    uint32_t pos = 0;
    if (asz == 256) {
        pos = 2;  // '0','0'
    } else if (asz == 0) {
        if (lane == 0) smem_put_bits(hw, 1, 1, 1);  // '0','1' (cannot happen for len > 0)
        pos = 2;
    } else {
        const uint32_t last_mask = last_sym >> 3;
        if (lane == 0) {
            smem_put_bits(hw, 0, 1, 1);
            smem_put_bits(hw, 1, last_mask, 5);
        }
        pos = 6;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (f[j] != 0) {
                const uint32_t s = (uint32_t)lane + 32u * j;
                smem_put_bits(hw, pos + 8 * (s >> 3) + (7 - (s & 7)), 1, 1);
            }
        }
        pos += 8 * (last_mask + 1);
    }
    if (asz > 0) {
        if (lane == 0) smem_put_bits(hw, pos, lr - 8, 3);
        pos += 3;
    }
    if (asz > 1) {
        const int gs = asz < 64 ? 6 : 8;
        uint32_t llr = 3;
        while ((1u << llr) <= lr) llr++;
        const int ngroups = (asz - 1 + gs - 1) / gs;  // <= 32
        uint32_t log_max = 0, glen = 0;
        const int first = 1 + lane * gs;
        int count = 0;
        if (lane < ngroups) {
            count = asz - first < gs ? asz - first : gs;
            uint32_t mx = 0;
            for (int k = 0; k < count; k++) mx = max(mx, (uint32_t)s_rank[warp][first + k]);
            log_max = mx ? 32u - (uint32_t)__clz((int)mx) : 0u;
            glen = llr + (uint32_t)count * log_max;
        }
        const uint32_t gincl = warp_incl_scan(glen, lane);
        const uint32_t goff = pos + gincl - glen;
        if (lane < ngroups) {
            smem_put_bits(hw, goff, log_max, llr);
            if (log_max) {
                uint32_t p = goff + llr;
                for (int k = 0; k < count; k++, p += log_max) smem_put_bits(hw, p, s_rank[warp][first + k], log_max);
            }
        }
        pos += __shfl_sync(0xFFFFFFFFu, gincl, 31);
    }
    __syncwarp();
    uint32_t* gout = reinterpret_cast<uint32_t*>(outbuf + (size_t)c * RC_OUT_STRIDE);
    // header words: when a payload follows, the partial last word is written by the bit appender below
    const uint32_t hwords = asz > 1 ? (pos >> 5) : ((pos + 31) >> 5);
    for (uint32_t i = lane; i < hwords; i += 32) gout[i] = bswap32(hw[i]);
    __syncwarp();
    const uint32_t si = seg_index[c];
    if (lane != 0) return;
    seg_src[si] = (uint64_t)(uintptr_t)gout;
    if (asz <= 1) {  // Write :256-260: chunk skipped, header only
        seg_bits[si] = pos;
        return;
    }
    // ---- the arithmetic chain (lane 0). The bit appender resumes inside the last header word.
    WordBitWriter bw;
    bw.init(gout + (pos >> 5));
    if (pos & 31) {
        bw.acc = hw[pos >> 5] >> (32 - (pos & 31));
        bw.n = pos & 31;
    }
    bw.total = pos;
    const uint16_t* cum = s_cum[warp];
    uint64_t rng = RC_TOP, low = 0;
    bool overflow = false;
    const uint64_t cap_bits = 8ull * (RC_OUT_STRIDE - 64);
    const uint4* src4 = reinterpret_cast<const uint4*>(src);  // chunk starts are 16-byte aligned
    auto encode_byte = [&](uint32_t s) {  // encodeByte :281-305
        const uint64_t cf = cum[s];
        rng >>= lr;
        low += cf * rng;
        rng *= (uint64_t)(cum[s + 1] - cf);
        for (;;) {
            if (((low ^ (low + rng)) & RC_MASK) != 0) {
                if (rng > RC_BOTTOM) break;
                rng = (0ull - low) & RC_BOTTOM;
            }
            if (bw.total + 128 > cap_bits) overflow = true;  // the block is failed below
            else bw.put(low >> 32, 28);
            rng <<= 28;
            low <<= 28;
        }
    };
    const uint32_t n16 = len >> 4;
    for (uint32_t i = 0; i < n16; i++) {
        const uint4 v = __ldg(&src4[i]);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            encode_byte(w4[q] & 0xFF);
            encode_byte((w4[q] >> 8) & 0xFF);
            encode_byte((w4[q] >> 16) & 0xFF);
            encode_byte(w4[q] >> 24);
        }
    }
    for (uint32_t i = n16 << 4; i < len; i++) encode_byte(src[i]);
    bw.put(low >> 32, 28);  // WriteBits(low, 60) (:266): the low 60 bits, most significant first
    bw.put(low, 32);
    const uint64_t total = bw.total;
    bw.flush();
    seg_bits[si] = total;
    // the scratch area holds 3 bytes per symbol; a chunk that needs more (never seen; would need a forced 28-bit
    // renormalisation on nearly every byte) is reported instead of overrunning the buffer
    if (overflow) atomicCAS(&status[ck.block], 0, -KZ_E_PROCESS_BLOCK);
}

// ------------------------------------------------------------------------------------------------------------
// decode: one warp per block
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) range_decode_kernel(const uint32_t* __restrict__ words, const DecBlock* __restrict__ blocks,
                                                          const DecChunk* __restrict__ chunks, int nblocks, uint8_t* __restrict__ out,
                                                          uint64_t* __restrict__ blk_end, int32_t* __restrict__ status) {
    __shared__ uint8_t s_f2s[4096];
    __shared__ uint16_t s_cum[260];
    __shared__ uint16_t s_freq[256];
    __shared__ uint8_t s_alpha[256];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nblocks) return;
    const DecBlock blk = blocks[b];
    if (blk.nchunks == 0 || status[b] != 0) return;
    uint8_t* dst = out + chunks[blk.chunk_base].out_off;
    uint64_t pos = blk.data_bit;
    int32_t st = 0;
    uint32_t start = 0;
    const uint32_t len = blk.pre_len;
    while (start < len && st == 0) {
        const uint32_t end = start + (uint32_t)RANGE_CHUNK < len ? start + (uint32_t)RANGE_CHUNK : len;
        // ---- header (decodeHeader :403-487); every lane parses the few scalar fields redundantly
        BitReader br(words, pos, blk.end_bit);
        int asz = 0;
        for (int i = lane; i < 256; i += 32) s_freq[i] = 0;
        __syncwarp();
        if (br.read(1) == 0) {
            if (br.read(1) == 0) {
                asz = 256;
                for (int i = lane; i < 256; i += 32) s_alpha[i] = (uint8_t)i;
            }
        } else {
            const uint32_t last = br.read(5);
            // mask byte i covers symbols 8i..8i+7, bit j of the byte = symbol 8i + j (EntropyUtils.go:96-112)
            const uint32_t m = (uint32_t)lane <= last ? bits_at(words, br.pos + 8 * lane, 8) : 0u;
            const uint32_t before = warp_incl_scan((uint32_t)__popc(m), lane) - (uint32_t)__popc(m);
            uint32_t mm = m, r = before;
            while (mm) {
                const int j = __ffs((int)mm) - 1;
                mm &= mm - 1;
                s_alpha[r++] = (uint8_t)(8 * lane + j);
            }
            asz = (int)__shfl_sync(0xFFFFFFFFu, before + (uint32_t)__popc(m), 31);
            br.skip(8ull * (last + 1));
        }
        __syncwarp();
        if (asz == 0) {
            st = -KZ_E_PROCESS_BLOCK;
            break;
        }
        const uint32_t lr = 8 + br.read(3);
        const uint32_t scale = 1u << lr;
        if (lr > 12) {  // tables are sized for the encoder's range (<= 12)
            st = -KZ_E_INVALID_CODEC;
            break;
        }
        uint32_t llr = 3;
        while ((1u << llr) <= lr) llr++;
        bool bad = false;
        uint32_t sum = 0;
        if (asz > 1) {  // frequency groups: serial chase of the group lengths, values extracted by lane 0 (tiny)
            const int gs = asz < 64 ? 6 : 8;
            if (lane == 0) {
                for (int i = 1; i < asz; i += gs) {
                    const uint32_t log_max = br.read(llr);
                    if ((1u << log_max) > scale) bad = true;
                    const int endj = i + gs < asz ? i + gs : asz;
                    for (int j = i; j < endj; j++) {
                        uint32_t fq = 1;
                        if (log_max > 0 && !bad) {
                            fq = 1 + br.read(log_max);
                            if (fq >= scale) bad = true;
                        }
                        s_freq[s_alpha[j]] = (uint16_t)fq;
                        sum += fq;
                    }
                }
                if (scale <= sum) bad = true;
                if (!bad) s_freq[s_alpha[0]] = (uint16_t)(scale - sum);
            }
            br.pos = __shfl_sync(0xFFFFFFFFu, br.pos, 0);
            br.overrun = __shfl_sync(0xFFFFFFFFu, (int)br.overrun, 0) != 0;
            bad = __shfl_sync(0xFFFFFFFFu, (int)bad, 0) != 0;
        } else if (lane == 0) {
            s_freq[s_alpha[0]] = (uint16_t)scale;
        }
        __syncwarp();
        if (bad || br.overrun) {
            st = -KZ_E_PROCESS_BLOCK;
            break;
        }
        if (asz == 1) {  // Read :511-515 (memset; nothing else is read)
            const uint8_t v = s_alpha[0];
            for (uint32_t i = start + lane; i < end; i += 32) dst[i] = v;
            pos = br.pos;
            start = end;
            __syncwarp();
            continue;
        }
        {  // cumulated frequencies and the slot -> symbol table (:470-483)
            uint32_t carry = 0;
            for (int j = 0; j < 8; j++) {
                const uint32_t fq = s_freq[lane + 32 * j];
                const uint32_t incl = warp_incl_scan(fq, lane);
                const uint32_t c0 = carry + incl - fq;
                s_cum[lane + 32 * j + 1] = (uint16_t)(c0 + fq);
                for (uint32_t q = c0; q < c0 + fq && q < 4096; q++) s_f2s[q] = (uint8_t)(lane + 32 * j);
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (lane == 0) s_cum[0] = 0;
        }
        __syncwarp();
        if (lane == 0) {  // ---- the arithmetic chain (decodeByte :530-557)
            uint64_t rng = RC_TOP, low = 0;
            uint64_t code = ((uint64_t)br.read(28) << 32) | br.read(32);  // ReadBits(60)
            for (uint32_t i = start; i < end; i++) {
                rng >>= lr;
                const uint64_t count = rng ? (code - low) / rng : ~0ull;  // division by zero panics in the reference
                if (count >= scale) {
                    st = -KZ_E_PROCESS_BLOCK;
                    break;
                }
                const uint32_t s = s_f2s[count];
                const uint64_t cf = s_cum[s];
                low += cf * rng;
                rng *= (uint64_t)(s_cum[s + 1] - cf);
                for (;;) {
                    if (((low ^ (low + rng)) & RC_MASK) != 0) {
                        if (rng > RC_BOTTOM) break;
                        rng = (0ull - low) & RC_BOTTOM;
                    }
                    code = (code << 28) | br.read(28);
                    rng <<= 28;
                    low <<= 28;
                }
                dst[i] = (uint8_t)s;
            }
            if (br.overrun && st == 0) st = -KZ_E_PROCESS_BLOCK;
        }
        pos = __shfl_sync(0xFFFFFFFFu, br.pos, 0);
        st = __shfl_sync(0xFFFFFFFFu, st, 0);
        start = end;
        __syncwarp();
    }
    if (lane == 0) {
        blk_end[b] = pos;
        if (st) status[b] = st;
    }
}

}  // namespace kz
