// C ABI of the B200 engine (include/kanzi_b200.h) and the host-side orchestration of the per-block path.
//
// Host-side restatements (tiny, serial, run once per call) cite the reference next to each function; all per-byte
// work happens in the CUDA kernels (kz_ans.cu, kz_huffman.cu, kz_concat.cu, ...). There is deliberately NO CPU code
// path for any codec here: if the CUDA runtime is unusable every entry point fails.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../../include/kanzi_b200.h"
#include "kz_ans1.cuh"
#include "kz_engine.cuh"
#include "kz_bwt.cuh"
#include "kz_huffman.cuh"
#include "kz_lz.cuh"
#include "kz_range.cuh"
#include "kz_hash.cuh"
#include "kz_sbrt.cuh"
#include "kz_zrlt.cuh"
#include "kz_rolz.cuh"
#include "kz_alias.cuh"
#include "kz_fsd.cuh"
#include "kz_text.cuh"
#include "kz_utf.cuh"
#include "kz_exe.cuh"

#include <memory>

using namespace kz;

#define CK(x)                                                 \
    do {                                                      \
        cudaError_t e_ = (x);                                 \
        if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #x); \
    } while (0)

namespace {

const uint32_t BS_TYPE = 0x4B414E5A;  // io/CompressedStream.go:43
const uint32_t BS_VERSION = 6;        // :44
const size_t SMALL_BLOCK = 15;        // :51
const size_t RAW_PIECE = 64 * 1024;   // raw segments are cut so that one CTA never copies more than this

cudaEvent_t get_event(kz_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

// host-side phase timer (only when profiling): accumulates wall milliseconds under "host:<name>"
struct HostScope {
    kz_ctx* ctx;
    const char* name;
    std::chrono::steady_clock::time_point t0;
    HostScope(kz_ctx* c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostScope() {
        if (ctx->profile) {
            auto& e = ctx->prof[std::string("host:") + name];
            e.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            e.launches++;
        }
    }
};

struct LaunchScope {
    kz_ctx* ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* name;
    LaunchScope(kz_ctx* c, const char* n) : ctx(c), name(n) {
        ctx->launches++;
        if (ctx->profile) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, ctx->stream);
        }
    }
    ~LaunchScope() {
        if (ctx->profile) {
            cudaEventRecord(b, ctx->stream);
            ctx->pending.push_back({name, a, b});
        }
    }
};

// algorithmic bytes of one transform stage over a batch (stage input + stage output), kept under "stage:fwd:<id>" / "stage:inv:<id>"
void note_stage_bytes(kz_ctx* ctx, const char* dir, uint64_t id, uint64_t bytes_in, uint64_t bytes_out) {
    if (!ctx->profile) return;
    auto& e = ctx->prof[std::string("stage:") + dir + ":" + std::to_string((unsigned long long)id)];
    e.bytes_in += bytes_in;
    e.bytes_out += bytes_out;
    e.launches++;
}

void drain_profile(kz_ctx* ctx) {
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& p : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto& e = ctx->prof[p.name];
            e.ms += ms;
            e.launches++;
        }
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// adaptors for the data-parallel LZ forward path (kz_lz_par.cu): scratch from ctx->d_ws, every pass timed under its own name
struct CtxLzWorkspace : LzWorkspace {
    kz_ctx* ctx;
    explicit CtxLzWorkspace(kz_ctx* c) : ctx(c) {}
    uint8_t* ensure(size_t bytes) override { return ctx->d_ws.ensure(bytes) == cudaSuccess ? ctx->d_ws.as<uint8_t>() : nullptr; }
};
struct CtxLzHook : LzHook {
    kz_ctx* ctx;
    LaunchScope* cur = nullptr;
    explicit CtxLzHook(kz_ctx* c) : ctx(c) {}
    void begin(const char* name) override { cur = new LaunchScope(ctx, name); }
    void end() override {
        delete cur;
        cur = nullptr;
    }
};
// KZ_LZ_PARSE=serial selects the round-1 parse (one lane per block replays the reference loop); default: the data-parallel path
bool lz_serial_parse() {
    static const bool v = [] {
        const char* e = getenv("KZ_LZ_PARSE");
        return e && e[0] == 's';
    }();
    return v;
}

// LZ / LZX forward of a batch: the data-parallel path (kz_lz_par.cu); blocks it hands back (status 3: no fixed point of the table
// membership within its round limit) and everything under KZ_LZ_PARSE=serial go through the serial parse of kz_lz.cu.
int upload(kz_ctx* ctx, Packer& pk);
int lz_forward_batch(kz_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, std::vector<LzBlock>& lb, bool extra, std::vector<LzResult>& hres) {
    const uint32_t nblocks = (uint32_t)lb.size();
    hres.assign(nblocks, LzResult{1, 0, 0, 0, 0, 0, 0});
    bool serial = lz_serial_parse();
    if (!serial) {
        CtxLzWorkspace W(ctx);
        CtxLzHook H(ctx);
        CK(lz_forward_parallel(d_in, d_out, lb, extra, W, ctx->stream, hres, &H));
        for (uint32_t b = 0; b < nblocks; b++) serial = serial || hres[b].status == 3;
        if (!serial) return 0;
    }
    uint32_t max_len = 0;
    std::vector<LzBlock> sb(lb);
    for (uint32_t b = 0; b < nblocks; b++) {
        if (!lz_serial_parse() && hres[b].status != 3) sb[b].len = 0;  // done by the parallel path
        max_len = std::max(max_len, sb[b].len);
    }
    const uint64_t sstride = (lz_scratch_bytes(max_len) + 255) & ~size_t(255);
    const size_t hbytes = ((size_t)nblocks << (extra ? 19 : 16)) * 4;
    for (uint32_t b = 0; b < nblocks; b++) sb[b].scratch_off = (uint64_t)b * sstride;
    CK(ctx->d_ws.ensure(sstride * nblocks + hbytes + 4096));
    CK(ctx->d_lens.ensure((size_t)nblocks * sizeof(LzResult) + 64));
    Packer pk;
    const size_t o_lb = pk.add(sb.data(), sb.size() * sizeof(LzBlock));
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* ws = ctx->d_ws.as<uint8_t>();
    int32_t* d_hash = (int32_t*)(ws + sstride * nblocks);
    CK(cudaMemsetAsync(d_hash, 0, hbytes, ctx->stream));
    LzBlock* d_lb = (LzBlock*)(ctx->d_tables.as<uint8_t>() + o_lb);
    LzResult* d_res = ctx->d_lens.as<LzResult>();
    {
        LaunchScope ls(ctx, "lz_parse");
        lz_parse_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, d_lb, (int)nblocks, extra ? 1 : 0, d_hash, ws, d_out, d_res);
    }
    {
        LaunchScope ls(ctx, "lz_gather");
        lz_gather_kernel<<<dim3(nblocks, 8), 256, 0, ctx->stream>>>(d_in, d_lb, (int)nblocks, ws, d_res, d_out);
    }
    CK(cudaGetLastError());
    std::vector<LzResult> sres(nblocks);
    CK(cudaMemcpyAsync(sres.data(), d_res, (size_t)nblocks * sizeof(LzResult), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t b = 0; b < nblocks; b++)
        if (lz_serial_parse() || hres[b].status == 3) hres[b] = sres[b];
    return 0;
}

int upload(kz_ctx* ctx, Packer& pk) {
    CK(ctx->h_stage.ensure(pk.bytes.size() + 256));
    CK(ctx->d_tables.ensure(pk.bytes.size() + 256));
    memcpy(ctx->h_stage.p, pk.bytes.data(), pk.bytes.size());
    CK(cudaMemcpyAsync(ctx->d_tables.p, ctx->h_stage.p, pk.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// XXHash32/64 of a set of blocks (hash/XXHash32.go, XXHash64.go; seed "KANZ", io/CompressedStream.go:354-356). Synchronous.
int hash_blocks(kz_ctx* ctx, const uint8_t* d_base, const std::vector<uint64_t>& off, const std::vector<uint32_t>& len, uint32_t bits,
                std::vector<uint64_t>& out) {
    const size_t nb = off.size();
    out.assign(nb, 0);
    if (nb == 0) return 0;
    std::vector<HashJob> hj(nb);
    for (size_t b = 0; b < nb; b++) {
        hj[b].off = off[b];
        hj[b].len = len[b];
        hj[b].pad = 0;
    }
    Packer pk;
    const size_t o_j = pk.add(hj.data(), hj.size() * sizeof(HashJob));
    const size_t o_h = pk.reserve(nb * 8);
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* T = ctx->d_tables.as<uint8_t>();
    {
        LaunchScope ls(ctx, "xxhash");
        xxhash_blocks_kernel<<<(int)nb, 32, 0, ctx->stream>>>(d_base, (const HashJob*)(T + o_j), (int)nb, (int)bits, (uint64_t*)(T + o_h));
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out.data(), T + o_h, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bits == 32)
        for (auto& h : out) h &= 0xFFFFFFFFull;
    return 0;
}

// decodingTask.decode :1990-2006: hash of the decoded block against the transmitted checksum
int verify_checksums(kz_ctx* ctx, const uint8_t* d_base, const std::vector<uint64_t>& off, const std::vector<uint32_t>& len, uint32_t bits,
                     std::vector<DecJob>& jobs) {
    std::vector<uint64_t> hv;
    int rc = hash_blocks(ctx, d_base, off, len, bits, hv);
    if (rc) return rc;
    for (size_t b = 0; b < jobs.size(); b++) {
        const uint64_t want = bits == 32 ? (jobs[b].checksum & 0xFFFFFFFFull) : jobs[b].checksum;
        if (hv[b] != want) {
            jobs[b].status = -KZ_ERR_CRC_CHECK;
            return ctx->fail(KZ_ERR_CRC_CHECK, "Corrupted bitstream: block checksum mismatch");
        }
    }
    return 0;
}

bool etype_supported(uint32_t e) { return e == KZ_E_NONE || e == KZ_E_ANS0 || e == KZ_E_HUFFMAN || e == KZ_E_ANS1 || e == KZ_E_RANGE; }

// ------------------------------------------------------------------------------------------------------------
// entropy encode of a batch of blocks + bit-granular assembly
// ------------------------------------------------------------------------------------------------------------
struct EncResult {
    uint64_t total_bits = 0;          // LAYOUT_STREAM: bits of the complete stream
    std::vector<uint64_t> blk_bits;   // per block "written"
    std::vector<int32_t> status;      // per block 0 / -(error code)
};

int encode_batch(kz_ctx* ctx, EncLayout layout, uint32_t etype, uint32_t checksum_bits, const uint8_t* d_data, const std::vector<EncJob>& jobs,
                 const std::vector<uint8_t>& stream_hdr, uint64_t stream_hdr_bits, uint8_t* d_dst, size_t dst_cap,
                 const std::vector<uint64_t>& blk_base_bits, EncResult& res) {
    const int nblocks = (int)jobs.size();
    if (!etype_supported(etype)) return ctx->fail(KZ_ERR_CREATE_CODEC, "entropy codec not available on the GPU path");
    HostScope hs_all(ctx, "encode_batch");
    std::unique_ptr<HostScope> hs_tab(new HostScope(ctx, "enc_tables"));
    // ---- host: segment and chunk tables. They are a pure function of the batch description (layout, codec, block
    // lengths and flags, buffer addresses), and a stream of equal-sized blocks repeats the same description batch after
    // batch (Writer.processBlock is called once per `jobs` blocks): the packed tables of the last batch are kept and
    // reused when the description is byte-identical (0.6 ms of host work per 64 x 4 MiB batch otherwise).
    const size_t small_hdr_off = 0;
    const size_t small_blk_off = 64;
    const size_t small_prefix_off = small_blk_off + 16 * (size_t)nblocks;
    const size_t small_zero_off = small_prefix_off + 8 * (size_t)nblocks;
    const size_t small_lr1_off = small_zero_off + 8;  // 3 bits "lr - 8" of the order-1 chunks: 11 - 8 = 0b011
    const size_t small_size = small_zero_off + 16;
    CK(ctx->d_small.ensure(small_size));
    uint8_t* d_small = ctx->d_small.as<uint8_t>();
    std::vector<uint8_t> key;
    {
        auto put = [&](const void* p, size_t n) {
            const uint8_t* q = (const uint8_t*)p;
            key.insert(key.end(), q, q + n);
        };
        const uint64_t scal[8] = {(uint64_t)layout | (ctx->fragment ? 0x100u : 0u), etype, checksum_bits, (uint64_t)(uintptr_t)d_data, (uint64_t)(uintptr_t)d_small, (uint64_t)nblocks,
                                  stream_hdr_bits, (uint64_t)stream_hdr.size()};
        key.reserve(sizeof(scal) + stream_hdr.size() + jobs.size() * 40 + blk_base_bits.size() * 8);
        put(scal, sizeof(scal));
        put(stream_hdr.data(), stream_hdr.size());
        for (const EncJob& j : jobs) {
            const uint64_t f[4] = {j.data_off, ((uint64_t)j.post_len << 32) | j.ntransforms, ((uint64_t)j.skip_flags << 8) | (j.copy ? 1u : 0u), j.checksum};
            put(f, sizeof(f));
        }
        put(blk_base_bits.data(), blk_base_bits.size() * 8);
    }
    EncPlan& plan = ctx->enc_plan;
    if (!(plan.valid && plan.key == key)) {
        plan.valid = false;
        plan.build(key, [&](EncPlan& P) {
            std::vector<uint8_t> small(small_size, 0);
            if (!stream_hdr.empty()) memcpy(&small[small_hdr_off], stream_hdr.data(), stream_hdr.size());
            small[small_lr1_off] = 0x60;

            std::vector<uint64_t> seg_src, seg_bits;
            std::vector<uint32_t> blk_first(nblocks + 1);
            std::vector<ChunkIn> chunks;
            std::vector<uint32_t> chunk_seg;
            auto add_seg = [&](uint64_t src, uint64_t bits) {
                seg_src.push_back(src);
                seg_bits.push_back(bits);
            };
            auto add_raw = [&](const uint8_t* p, uint64_t nbytes) {
                for (uint64_t o = 0; o < nbytes; o += RAW_PIECE) add_seg((uint64_t)(uintptr_t)(p + o), 8 * std::min<uint64_t>(RAW_PIECE, nbytes - o));
            };
            if (layout == LAYOUT_STREAM) add_seg((uint64_t)(uintptr_t)(d_small + small_hdr_off), stream_hdr_bits);
            for (int b = 0; b < nblocks; b++) {
                const EncJob& j = jobs[b];
                if (layout == LAYOUT_STREAM) add_seg(0, 0);  // block prefix, filled by block_prefix_kernel
                blk_first[b] = (uint32_t)seg_src.size();
                const uint32_t et = j.copy ? (uint32_t)KZ_E_NONE : etype;
                if (layout != LAYOUT_RAW) {
                    uint8_t* hb = &small[small_blk_off + 16 * (size_t)b];
                    uint32_t hbits = make_block_header(hb, j.copy, j.ntransforms, j.skip_flags, j.post_len, checksum_bits, j.checksum);
                    add_seg((uint64_t)(uintptr_t)(d_small + small_blk_off + 16 * (size_t)b), hbits);
                }
                const uint8_t* data = d_data + j.data_off;
                if (et == KZ_E_NONE || ((et == KZ_E_ANS0 || et == KZ_E_ANS1) && j.post_len <= 32)) {  // NullEntropyCodec.go:43-61 / ANSRangeCodec.go:279-282
                    add_raw(data, j.post_len);
                } else if (et == KZ_E_ANS1) {
                    for (uint32_t o = 0; o < j.post_len; o += ANS1_CHUNK) {
                        ChunkIn c;
                        c.src_off = j.data_off + o;
                        c.len = std::min<uint32_t>(ANS1_CHUNK, j.post_len - o);
                        c.block = (uint32_t)b;
                        chunks.push_back(c);
                        add_seg((uint64_t)(uintptr_t)(d_small + small_lr1_off), 3);  // lr - 8 (updateFrequencies :174)
                        chunk_seg.push_back((uint32_t)seg_src.size());
                        for (int q = 0; q < ANS1_SEGS - 1; q++) add_seg(0, 0);  // 256 context tables, size + states, payload
                    }
                } else if (et == KZ_E_ANS0) {
                    for (uint32_t o = 0; o < j.post_len; o += ANS0_CHUNK) {
                        ChunkIn c;
                        c.src_off = j.data_off + o;
                        c.len = std::min<uint32_t>(ANS0_CHUNK, j.post_len - o);
                        c.block = (uint32_t)b;
                        chunks.push_back(c);
                        chunk_seg.push_back((uint32_t)seg_src.size());
                        add_seg(0, 0);  // A: chunk header + varint + states
                        add_seg(0, 0);  // B: rANS bytes
                    }
                } else if (et == KZ_E_RANGE) {  // RangeCodec.go:233-272: one bit string per 32 KiB chunk
                    for (uint32_t o = 0; o < j.post_len; o += RANGE_CHUNK) {
                        ChunkIn c;
                        c.src_off = j.data_off + o;
                        c.len = std::min<uint32_t>(RANGE_CHUNK, j.post_len - o);
                        c.block = (uint32_t)b;
                        chunks.push_back(c);
                        chunk_seg.push_back((uint32_t)seg_src.size());
                        add_seg(0, 0);
                    }
                } else if (et == KZ_E_HUFFMAN) {  // HuffmanCodec.go:408-430
                    for (uint32_t o = 0; o < j.post_len; o += HUF_CHUNK) {
                        const uint32_t clen = std::min<uint32_t>(HUF_CHUNK, j.post_len - o);
                        if (clen < 32) {
                            add_raw(data + o, clen);
                            continue;
                        }
                        ChunkIn c;
                        c.src_off = j.data_off + o;
                        c.len = clen;
                        c.block = (uint32_t)b;
                        chunks.push_back(c);
                        chunk_seg.push_back((uint32_t)seg_src.size());
                        for (int q = 0; q < 6; q++) add_seg(0, 0);  // header+sizes, 4 fragments, chunk last bytes
                    }
                }
            }
            if (layout == LAYOUT_STREAM) {
                add_seg((uint64_t)(uintptr_t)(d_small + small_zero_off), ctx->fragment ? 0 : 8);  // end marker: 5+3 zero bits (Close :593-594); none behind a fragment
                blk_first[nblocks] = (uint32_t)seg_src.size();
            } else {
                blk_first[nblocks] = (uint32_t)seg_src.size();
            }
            P.nsegs = (int)seg_src.size();
            add_seg(0, 0);  // sentinel for the scan
            P.nchunks = (int)chunks.size();
            P.max_chunk = 0;
            for (auto& c : chunks) P.max_chunk = std::max(P.max_chunk, c.len);

            Packer& pk = P.pk;
            pk.bytes.clear();
            P.o_small = pk.add(small.data(), small.size());
            P.o_src = pk.add(seg_src.data(), seg_src.size() * 8);
            P.o_bits = pk.add(seg_bits.data(), seg_bits.size() * 8);
            P.o_first = pk.add(blk_first.data(), blk_first.size() * 4);
            P.o_chunks = pk.add(chunks.data(), chunks.size() * sizeof(ChunkIn));
            P.o_cseg = pk.add(chunk_seg.data(), chunk_seg.size() * 4);
            P.o_base = pk.add(blk_base_bits.data(), blk_base_bits.size() * 8);
            P.o_off = pk.reserve(seg_src.size() * 8);
            P.o_dst = pk.reserve(seg_src.size() * 8);
            P.o_blkbits = pk.reserve((size_t)(nblocks + 1) * 8);
            P.o_total = pk.reserve(16);
            P.o_hbits = pk.reserve((size_t)P.nchunks * 4 + 4);
            P.o_asz = pk.reserve((size_t)P.nchunks * 4 + 4);
            P.o_estatus = pk.reserve((size_t)nblocks * 4 + 4);
        });
    }
    Packer& pk = plan.pk;
    const int nsegs = plan.nsegs, nchunks = plan.nchunks;
    const size_t o_small = plan.o_small, o_src = plan.o_src, o_bits = plan.o_bits, o_first = plan.o_first, o_chunks = plan.o_chunks, o_cseg = plan.o_cseg,
                 o_base = plan.o_base, o_off = plan.o_off, o_dst = plan.o_dst, o_blkbits = plan.o_blkbits, o_total = plan.o_total, o_hbits = plan.o_hbits,
                 o_asz = plan.o_asz, o_estatus = plan.o_estatus;
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* T = ctx->d_tables.as<uint8_t>();
    // the small byte strings must live at the addresses baked into the segment table
    CK(cudaMemcpyAsync(d_small, T + o_small, small_size, cudaMemcpyDeviceToDevice, ctx->stream));
    uint64_t* d_seg_src = (uint64_t*)(T + o_src);
    uint64_t* d_seg_bits = (uint64_t*)(T + o_bits);
    uint32_t* d_blk_first = (uint32_t*)(T + o_first);
    ChunkIn* d_chunks = (ChunkIn*)(T + o_chunks);
    uint32_t* d_cseg = (uint32_t*)(T + o_cseg);
    uint64_t* d_base = (uint64_t*)(T + o_base);
    uint64_t* d_seg_off = (uint64_t*)(T + o_off);
    uint64_t* d_seg_dst = (uint64_t*)(T + o_dst);
    uint64_t* d_blk_bits = (uint64_t*)(T + o_blkbits);
    uint64_t* d_total = (uint64_t*)(T + o_total);
    uint32_t* d_hbits = (uint32_t*)(T + o_hbits);
    uint32_t* d_asz = (uint32_t*)(T + o_asz);
    int32_t* d_estatus = (int32_t*)(T + o_estatus);
    hs_tab.reset();

    // ---- entropy kernels
    if (nchunks > 0 && etype == KZ_E_ANS1) {
        const uint32_t max_chunk = plan.max_chunk;
        const uint64_t pay_stride = (2ull * max_chunk + 96 + 15) & ~15ull;
        CK(ctx->d_enc_tab.ensure((size_t)nchunks * 65536 * sizeof(uint32_t)));
        CK(ctx->d_hdr.ensure((size_t)nchunks * 256 * ANS1_CTX_HDR_STRIDE + (size_t)nchunks * 32 + 256));
        CK(ctx->d_pay.ensure((size_t)nchunks * pay_stride));
        CK(ctx->d_hist.ensure((size_t)nchunks * ANS1_HIST_WORDS * 4 + (size_t)nchunks * 256 * 4 + 256));
        if (!ctx->d_rcp.p) {
            CK(ctx->d_rcp.ensure(4096 * sizeof(uint32_t)));
            LaunchScope ls(ctx, "rcp_table");
            rcp_table_kernel<<<16, 256, 0, ctx->stream>>>(ctx->d_rcp.as<uint32_t>());
        }
        uint32_t* d_hist = ctx->d_hist.as<uint32_t>();
        uint32_t* d_ctx_bits = d_hist + (size_t)nchunks * ANS1_HIST_WORDS;
        uint8_t* d_ctx_hdr = ctx->d_hdr.as<uint8_t>();
        uint8_t* d_tail_hdr = d_ctx_hdr + (size_t)nchunks * 256 * ANS1_CTX_HDR_STRIDE;
        CK(cudaMemsetAsync(d_hist, 0, (size_t)nchunks * ANS1_HIST_WORDS * 4, ctx->stream));
        {
            LaunchScope ls(ctx, "ans1_hist");
            ans1_hist_kernel<<<dim3(nchunks, 32), 256, 0, ctx->stream>>>(d_data, d_chunks, nchunks, d_hist);
        }
        {
            LaunchScope ls(ctx, "ans1_stats");
            ans1_stats_kernel<<<(nchunks * 256 + 3) / 4, 128, 0, ctx->stream>>>(d_hist, nchunks, ctx->d_enc_tab.as<uint32_t>(), d_ctx_hdr, d_ctx_bits);
        }
        {
            LaunchScope ls(ctx, "ans1_encode");
            ans1_encode_kernel<<<nchunks, 32, 0, ctx->stream>>>(d_data, d_chunks, nchunks, ctx->d_enc_tab.as<uint32_t>(), ctx->d_rcp.as<uint32_t>(), d_ctx_hdr,
                                                                d_ctx_bits, d_tail_hdr, ctx->d_pay.as<uint8_t>(), pay_stride, d_seg_src, d_seg_bits, d_cseg,
                                                                d_estatus);
        }
    } else if (nchunks > 0 && etype == KZ_E_RANGE) {
        CK(ctx->d_pay.ensure((size_t)nchunks * RC_OUT_STRIDE));
        LaunchScope ls(ctx, "range_encode");
        range_encode_kernel<<<(nchunks + RC_ENC_WARPS - 1) / RC_ENC_WARPS, RC_ENC_WARPS * 32, 0, ctx->stream>>>(
            d_data, d_chunks, nchunks, ctx->d_pay.as<uint8_t>(), d_seg_src, d_seg_bits, d_cseg, d_estatus);
    } else if (nchunks > 0 && etype == KZ_E_HUFFMAN) {
        CK(ctx->d_enc_tab.ensure((size_t)nchunks * 256 * sizeof(uint16_t)));
        CK(ctx->d_hdr.ensure((size_t)nchunks * HUF_HDR_STRIDE));
        CK(ctx->d_pay.ensure((size_t)nchunks * HUF_FRAG_STRIDE * 4));
        {
            LaunchScope ls(ctx, "huf_stats");
            huf_stats_kernel<<<(nchunks + HUF_STATS_WARPS - 1) / HUF_STATS_WARPS, HUF_STATS_WARPS * 32, 0, ctx->stream>>>(
                d_data, d_chunks, nchunks, ctx->d_enc_tab.as<uint16_t>(), ctx->d_hdr.as<uint8_t>(), d_hbits, d_asz);
        }
        {
            LaunchScope ls(ctx, "huf_encode");
            huf_encode_kernel<<<(nchunks + HUF_ENC_WARPS - 1) / HUF_ENC_WARPS, HUF_ENC_WARPS * 32, 0, ctx->stream>>>(
                d_data, d_chunks, nchunks, ctx->d_enc_tab.as<uint16_t>(), ctx->d_hdr.as<uint8_t>(), d_hbits, d_asz, ctx->d_pay.as<uint8_t>(),
                d_seg_src, d_seg_bits, d_cseg);
        }
    } else if (nchunks > 0) {
        CK(ctx->d_enc_tab.ensure((size_t)nchunks * 256 * sizeof(uint32_t)));
        if (!ctx->d_rcp.p) {
            CK(ctx->d_rcp.ensure(4096 * sizeof(uint32_t)));
            LaunchScope ls(ctx, "rcp_table");
            rcp_table_kernel<<<16, 256, 0, ctx->stream>>>(ctx->d_rcp.as<uint32_t>());
        }
        CK(ctx->d_hdr.ensure((size_t)nchunks * HDR_STRIDE));
        CK(ctx->d_pay.ensure((size_t)nchunks * PAY_STRIDE));
        {
            LaunchScope ls(ctx, "ans0_stats");
            ans0_stats_kernel<<<(nchunks + STATS_WARPS - 1) / STATS_WARPS, STATS_WARPS * 32, 0, ctx->stream>>>(
                d_data, d_chunks, nchunks, ctx->d_enc_tab.as<uint32_t>(), ctx->d_hdr.as<uint8_t>(), d_hbits, d_asz);
        }
        {
            LaunchScope ls(ctx, "ans0_encode");
            ans0_encode_kernel<<<(nchunks + ENC_CTA_CHUNKS - 1) / ENC_CTA_CHUNKS, 256, ENC_SMEM, ctx->stream>>>(
                d_data, d_chunks, nchunks, ctx->d_enc_tab.as<uint32_t>(), ctx->d_rcp.as<uint32_t>(), ctx->d_hdr.as<uint8_t>(), d_hbits, d_asz,
                ctx->d_pay.as<uint8_t>(), d_seg_src, d_seg_bits, d_cseg);
        }
    }
    // ---- layout + gather
    {
        LaunchScope ls(ctx, "scan");
        scan_bits_kernel<<<1, 1024, 0, ctx->stream>>>(d_seg_bits, d_seg_off, nsegs + 1, 0, d_total);
    }
    if (layout == LAYOUT_STREAM) {
        if (nblocks > 0) {
            LaunchScope ls(ctx, "block_prefix");
            block_prefix_kernel<<<(nblocks + 127) / 128, 128, 0, ctx->stream>>>(d_seg_off, d_blk_first, nblocks, d_small + small_prefix_off, d_seg_src,
                                                                                d_seg_bits, d_blk_bits);
        }
        LaunchScope ls(ctx, "scan");
        scan_bits_kernel<<<1, 1024, 0, ctx->stream>>>(d_seg_bits, d_seg_dst, nsegs + 1, 0, d_total);
    } else if (nblocks > 0) {
        LaunchScope ls(ctx, "block_layout");
        block_layout_kernel<<<nblocks, 128, 0, ctx->stream>>>(d_seg_off, d_blk_first, nblocks, d_base, d_seg_dst, d_blk_bits);
    }
    // results needed on the host before the gather can be bounds-checked
    CK(ctx->h_dl.ensure((size_t)(nblocks + 4) * 8 + (size_t)nblocks * 4 + 64));
    uint64_t* h = (uint64_t*)ctx->h_dl.p;
    int32_t* hst = (int32_t*)(h + nblocks + 4);
    CK(cudaMemcpyAsync(h, d_blk_bits, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + nblocks, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hst, d_estatus, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    {
        HostScope hs_sync(ctx, "enc_sync");
        CK(cudaStreamSynchronize(ctx->stream));
    }
    res.blk_bits.assign(h, h + nblocks);
    res.total_bits = h[nblocks];
    res.status.assign(hst, hst + nblocks);
    for (int b = 0; b < nblocks; b++)
        if (res.status[b]) return ctx->fail(-res.status[b], "entropy encoding failed (the reference fails with ERR_PROCESS_BLOCK on this input)");
    if (layout == LAYOUT_STREAM) {
        if ((res.total_bits + 7) / 8 + 8 > dst_cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
    } else {
        for (int b = 0; b < nblocks; b++) {
            uint64_t lim = (b + 1 < nblocks ? blk_base_bits[b + 1] : 8 * (uint64_t)dst_cap);
            if (blk_base_bits[b] + res.blk_bits[b] + 64 > lim) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
        }
    }
    if (nsegs > 0) {
        {
            LaunchScope ls(ctx, "concat_zero");
            concat_zero_kernel<<<(nsegs + 255) / 256, 256, 0, ctx->stream>>>(d_seg_bits, d_seg_dst, nsegs, (uint32_t*)d_dst);
        }
        LaunchScope ls(ctx, "concat");
        concat_copy_kernel<<<nsegs, 128, 0, ctx->stream>>>(d_seg_src, d_seg_bits, d_seg_dst, nsegs, (uint32_t*)d_dst);
    }
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// entropy decode of a batch of blocks
// ------------------------------------------------------------------------------------------------------------
__global__ void block_header_kernel(const uint32_t* __restrict__ words, const uint64_t* __restrict__ bit_off, const uint64_t* __restrict__ bits, int nblocks,
                                    uint32_t checksum_bits, uint32_t block_size, BlockHeader* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    out[b] = parse_block_header(words, bit_off[b], bits[b], checksum_bits, block_size);
}

// Serial walk over the block length prefixes of a complete stream (decodingTask.decode :1816-1852). One thread.
// stop_after != 0: a fragment of that many block records (no end marker behind them)
__global__ void stream_walk_kernel(const uint32_t* __restrict__ words, uint64_t start_bit, uint64_t end_bit, uint32_t max_blocks,
                                   uint64_t* __restrict__ bit_off, uint64_t* __restrict__ bits, uint32_t* __restrict__ count_status, uint32_t stop_after = 0) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    BitReader br(words, start_bit, end_bit);
    uint32_t n = 0;
    int32_t st = 0;
    for (;;) {
        if (stop_after && n == stop_after) break;
        const uint32_t lr = br.read(5) + 3;
        uint64_t len;
        if (lr > 32) {
            uint64_t hi = br.read(lr - 32);
            len = (hi << 32) | br.read(32);
        } else {
            len = br.read(lr);
        }
        if (br.overrun) {
            st = -KZ_E_PROCESS_BLOCK;
            break;
        }
        if (len == 0) break;
        if (len > (1ull << 34)) {
            st = -KZ_E_BLOCK_SIZE;
            break;
        }
        if (n >= max_blocks) {
            st = -KZ_E_BLOCK_SIZE;
            break;
        }
        bit_off[n] = br.pos;
        bits[n] = len;
        n++;
        br.skip(len);
        if (br.overrun) {
            st = -KZ_E_PROCESS_BLOCK;
            break;
        }
    }
    count_status[0] = n;
    count_status[1] = (uint32_t)st;
}

// jobs[b].bit_off / bits / out_off must be set; fills pre_len, mode, ..., status. d_words = device stream buffer.
int decode_parse_headers(kz_ctx* ctx, const uint32_t* d_words, std::vector<DecJob>& jobs, uint32_t checksum_bits, uint32_t block_size) {
    const int nblocks = (int)jobs.size();
    if (nblocks == 0) return 0;
    std::vector<uint64_t> off(nblocks), bits(nblocks);
    for (int b = 0; b < nblocks; b++) {
        off[b] = jobs[b].bit_off;
        bits[b] = jobs[b].bits;
    }
    Packer pk;
    const size_t o_off = pk.add(off.data(), off.size() * 8);
    const size_t o_bits = pk.add(bits.data(), bits.size() * 8);
    const size_t o_out = pk.reserve(sizeof(BlockHeader) * (size_t)nblocks);
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* T = ctx->d_tables.as<uint8_t>();
    {
        LaunchScope ls(ctx, "block_header");
        block_header_kernel<<<(nblocks + 127) / 128, 128, 0, ctx->stream>>>(d_words, (uint64_t*)(T + o_off), (uint64_t*)(T + o_bits), nblocks, checksum_bits,
                                                                            block_size, (BlockHeader*)(T + o_out));
    }
    CK(ctx->h_dl.ensure(sizeof(BlockHeader) * (size_t)nblocks));
    CK(cudaMemcpyAsync(ctx->h_dl.p, T + o_out, sizeof(BlockHeader) * (size_t)nblocks, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const BlockHeader* h = (const BlockHeader*)ctx->h_dl.p;
    for (int b = 0; b < nblocks; b++) {
        jobs[b].pre_len = h[b].pre_len;
        jobs[b].mode = h[b].mode;
        jobs[b].skip_flags = h[b].skip_flags;
        jobs[b].copy = h[b].copy;
        jobs[b].checksum = h[b].checksum;
        jobs[b].data_bit = h[b].data_bit;
        jobs[b].status = h[b].status;
    }
    return 0;
}

// Decodes the entropy data of every job (data_bit, pre_len must be valid) to d_out + out_off. with_end: fetch end_pos.
int decode_batch(kz_ctx* ctx, uint32_t etype, const uint32_t* d_words, uint64_t words_len, std::vector<DecJob>& jobs, uint8_t* d_out) {
    const int nblocks = (int)jobs.size();
    if (nblocks == 0) return 0;
    if (!etype_supported(etype)) return ctx->fail(KZ_ERR_INVALID_CODEC, "entropy codec not available on the GPU path");
    std::vector<DecBlock> dblocks(nblocks);
    std::vector<ExtractJob> xjobs;
    std::vector<int32_t> status(nblocks, 0);
    uint32_t nchunks_total = 0, max_chunks = 0;
    for (int b = 0; b < nblocks; b++) {
        DecJob& j = jobs[b];
        DecBlock& db = dblocks[b];
        db.data_bit = j.data_bit;
        db.end_bit = j.bit_off + j.bits;
        db.pre_len = j.pre_len;
        db.nchunks = 0;
        db.chunk_base = nchunks_total;
        db.chunk_size = 0;
        db.out_off = j.out_off;
        status[b] = j.status;
        if (j.status) continue;
        const uint32_t et = j.copy ? (uint32_t)KZ_E_NONE : etype;
        if (et == KZ_E_NONE || ((et == KZ_E_ANS0 || et == KZ_E_ANS1) && j.pre_len <= 32)) {
            if (j.data_bit + 8ull * j.pre_len > db.end_bit) {
                status[b] = j.status = -KZ_ERR_PROCESS_BLOCK;
                continue;
            }
            for (uint64_t o = 0; o < j.pre_len; o += RAW_PIECE) {
                ExtractJob x;
                x.src_bit = j.data_bit + 8 * o;
                x.dst_off = j.out_off + o;
                x.nbytes = (uint32_t)std::min<uint64_t>(RAW_PIECE, j.pre_len - o);
                x.pad = 0;
                xjobs.push_back(x);
            }
            j.end_pos = j.data_bit + 8ull * j.pre_len;
        } else {
            // the chunk descriptors are expanded on the device (expand_chunks_kernel); Range blocks are one serial chain: one entry
            const uint32_t cs = et == KZ_E_ANS0 ? (uint32_t)ANS0_CHUNK : et == KZ_E_ANS1 ? (uint32_t)ANS1_CHUNK : et == KZ_E_HUFFMAN ? (uint32_t)HUF_CHUNK
                                                                                                                                     : 0xFFFFFFFFu;
            db.chunk_size = cs;
            db.nchunks = cs == 0xFFFFFFFFu ? 1u : (uint32_t)(((uint64_t)j.pre_len + cs - 1) / cs);
            nchunks_total += db.nchunks;
            max_chunks = std::max(max_chunks, db.nchunks);
        }
    }
    const int nchunks = (int)nchunks_total;
    Packer pk;
    const size_t o_blocks = pk.add(dblocks.data(), dblocks.size() * sizeof(DecBlock));
    const size_t o_x = pk.add(xjobs.data(), xjobs.size() * sizeof(ExtractJob));
    const size_t o_status = pk.add(status.data(), status.size() * 4);
    const size_t o_hbit = pk.reserve((size_t)nchunks * 8 + 8);
    const size_t o_end = pk.reserve((size_t)nblocks * 8 + 8);
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* T = ctx->d_tables.as<uint8_t>();
    int32_t* d_status = (int32_t*)(T + o_status);
    CK(ctx->d_chunks.ensure((size_t)nchunks * sizeof(DecChunk) + 64));
    DecChunk* d_dchunks = ctx->d_chunks.as<DecChunk>();
    if (nchunks > 0) {
        LaunchScope ls(ctx, "expand_chunks");
        expand_chunks_kernel<<<dim3((max_chunks + 255) / 256, nblocks), 256, 0, ctx->stream>>>((DecBlock*)(T + o_blocks), nblocks, d_dchunks);
    }
    if (nchunks > 0 && etype == KZ_E_ANS1) {
        CK(ctx->d_enc_tab.ensure((size_t)nchunks * 65536 * sizeof(uint32_t)));
        CK(ctx->d_hist.ensure((size_t)nchunks * 256 * 2048 + 256));
        {
            LaunchScope ls(ctx, "ans_walk");
            ans_walk_kernel<<<(nblocks + 31) / 32, 32, 0, ctx->stream>>>(d_words, (DecBlock*)(T + o_blocks), nblocks, 1, (uint64_t*)(T + o_hbit),
                                                                         (uint64_t*)(T + o_end), d_status);
        }
        LaunchScope ls(ctx, "ans1_decode");
        ans1_decode_kernel<<<nchunks, 32, 0, ctx->stream>>>(d_words, words_len, d_dchunks, nchunks, (uint64_t*)(T + o_hbit),
                                                            ctx->d_hist.as<uint8_t>(), ctx->d_enc_tab.as<uint32_t>(), d_out, d_status);
    } else if (nchunks > 0 && etype == KZ_E_RANGE) {
        LaunchScope ls(ctx, "range_decode");
        range_decode_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_words, (DecBlock*)(T + o_blocks), d_dchunks, nblocks, d_out,
                                                             (uint64_t*)(T + o_end), d_status);
    } else if (nchunks > 0 && etype == KZ_E_HUFFMAN) {
        {
            LaunchScope ls(ctx, "huf_walk");
            huf_walk_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_words, words_len, (DecBlock*)(T + o_blocks), nblocks, (uint64_t*)(T + o_hbit),
                                                             (uint64_t*)(T + o_end), d_status);
        }
        LaunchScope ls(ctx, "huf_decode");
        huf_decode_kernel<<<(nchunks + HUF_DEC_CTA_CHUNKS - 1) / HUF_DEC_CTA_CHUNKS, HUF_DEC_CTA_CHUNKS * 4, HUF_DEC_SMEM, ctx->stream>>>(
            d_words, words_len, d_dchunks, nchunks, (uint64_t*)(T + o_hbit), d_out, d_status);
    } else if (nchunks > 0) {
        {
            LaunchScope ls(ctx, "ans_walk");
            ans0_walk_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_words, words_len, (DecBlock*)(T + o_blocks), nblocks, (uint64_t*)(T + o_hbit),
                                                              (uint64_t*)(T + o_end), d_status);
        }
        if (ctx->ans0_dec_version >= 5) {
            CK(ctx->d_enc_tab.ensure((size_t)nchunks * DEC_TAB_WORDS * sizeof(uint32_t) + 256));
            LaunchScope ls(ctx, "ans0_tables");
            ans0_tables_kernel<<<(nchunks + 3) / 4, 128, 0, ctx->stream>>>(d_words, words_len, d_dchunks, nchunks, (uint64_t*)(T + o_hbit),
                                                                        ctx->d_enc_tab.as<uint32_t>(), d_out, d_status);
        }
        LaunchScope ls(ctx, "ans0_decode");
        if (ctx->ans0_dec_version >= 6)
            ans0_decode6_kernel<<<(nchunks + DEC6_CHUNKS - 1) / DEC6_CHUNKS, 128, DEC6_SMEM, ctx->stream>>>(
                d_words, words_len, ctx->d_enc_tab.as<uint32_t>(), d_dchunks, nchunks, d_out);
        else if (ctx->ans0_dec_version >= 5)
            ans0_decode5_kernel<<<(nchunks + DEC_CTA_CHUNKS - 1) / DEC_CTA_CHUNKS, 64, DEC4_SMEM, ctx->stream>>>(
                d_words, words_len, ctx->d_enc_tab.as<uint32_t>(), d_dchunks, nchunks, (uint64_t*)(T + o_hbit), d_out, d_status);
        else if (ctx->ans0_dec_version == 3)
            ans0_decode_kernel<<<(nchunks + DEC_CTA_CHUNKS - 1) / DEC_CTA_CHUNKS, 64, DEC_SMEM, ctx->stream>>>(d_words, words_len, d_dchunks,
                                                                                                   nchunks, (uint64_t*)(T + o_hbit), d_out, d_status);
        else
            ans0_decode4_kernel<<<(nchunks + DEC_CTA_CHUNKS - 1) / DEC_CTA_CHUNKS, 64, DEC4_SMEM, ctx->stream>>>(d_words, words_len, d_dchunks,
                                                                                                     nchunks, (uint64_t*)(T + o_hbit), d_out, d_status);
    }
    if (!xjobs.empty()) {
        LaunchScope ls(ctx, "extract");
        extract_copy_kernel<<<(int)xjobs.size(), 256, 0, ctx->stream>>>(d_words, (ExtractJob*)(T + o_x), (int)xjobs.size(), d_out);
    }
    CK(cudaGetLastError());
    CK(ctx->h_dl.ensure((size_t)nblocks * 16 + 64));
    int32_t* hs = (int32_t*)ctx->h_dl.p;
    uint64_t* he = (uint64_t*)((uint8_t*)ctx->h_dl.p + (((size_t)nblocks * 4 + 15) & ~size_t(15)));
    CK(cudaMemcpyAsync(hs, d_status, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(he, T + o_end, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nblocks; b++) {
        jobs[b].status = hs[b];
        if (dblocks[b].nchunks) jobs[b].end_pos = he[b];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// stream header (Writer.writeHeader io/CompressedStream.go:429-519, Reader.readHeader :1316-1460)
// ------------------------------------------------------------------------------------------------------------
uint32_t header_checksum(uint32_t ck_size, uint32_t etype, uint64_t t48, uint32_t block_size, uint32_t sz_mask, int64_t input_size) {
    const uint32_t HASH = 0x1E35A7BDu;
    uint32_t ck = HASH * (0x01030507u * BS_VERSION);
    ck ^= HASH * ~ck_size;
    ck ^= HASH * ~etype;
    ck ^= HASH * (uint32_t)((~t48) >> 32);
    ck ^= HASH * (uint32_t)(~t48);
    ck ^= HASH * ~block_size;
    if (sz_mask > 0) {
        ck ^= HASH * (uint32_t)(((uint64_t)~input_size) >> 32);
        ck ^= HASH * (uint32_t)(~input_size);
    }
    return ((ck >> 23) ^ (ck >> 3)) & 0xFFFFFFu;
}

uint32_t build_stream_header(std::vector<uint8_t>& out, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size) {
    out.assign(64, 0);
    BitWriter bw(out.data());
    const uint32_t ck_size = checksum_bits == 32 ? 1 : (checksum_bits == 64 ? 2 : 0);
    bw.write(BS_TYPE, 32);
    bw.write(BS_VERSION, 4);
    bw.write(ck_size, 2);
    bw.write(etype, 5);
    bw.write((uint32_t)(t48 >> 32), 16);
    bw.write((uint32_t)t48, 32);
    bw.write(block_size >> 4, 28);
    uint32_t sz_mask;
    if (input_size <= 0 || input_size >= ((int64_t)1 << 48)) sz_mask = 0;
    else if (input_size >= ((int64_t)1 << 32)) sz_mask = 3;
    else if (input_size >= ((int64_t)1 << 16)) sz_mask = 2;
    else sz_mask = 1;
    bw.write(sz_mask, 2);
    if (sz_mask == 3) bw.write((uint32_t)((uint64_t)input_size >> 32), 16);
    if (sz_mask >= 2) bw.write((uint32_t)input_size, 32);
    if (sz_mask == 1) bw.write((uint32_t)input_size, 16);
    bw.write(0, 15);
    bw.write(header_checksum(ck_size, etype, t48, block_size, sz_mask, sz_mask ? input_size : 0), 24);
    return bw.pos;
}

struct StreamHeader {
    uint64_t t48;
    uint32_t etype, block_size, checksum_bits;
    int64_t input_size;
    uint64_t bits;  // header length
};

int parse_stream_header(kz_ctx* ctx, const uint8_t* bytes, size_t n, StreamHeader& h) {
    uint32_t tmp[20];
    memset(tmp, 0, sizeof(tmp));
    memcpy(tmp, bytes, std::min<size_t>(n, 64));
    BitReader br(tmp, 0, 8ull * std::min<size_t>(n, 64));
    if (br.read(32) != BS_TYPE) return ctx->fail(KZ_ERR_INVALID_FILE, "Invalid stream type");
    uint32_t ver = br.read(4);
    if (ver != BS_VERSION) return ctx->fail(KZ_ERR_STREAM_VERSION, "Only bitstream version 6 is supported");
    uint32_t ck_size = br.read(2);
    if (ck_size == 3) return ctx->fail(KZ_ERR_INVALID_CODEC, "Invalid bitstream, incorrect checksum size");
    h.checksum_bits = 32 * ck_size;
    h.etype = br.read(5);
    uint64_t hi = br.read(16);
    h.t48 = (hi << 32) | br.read(32);
    h.block_size = br.read(28) << 4;
    if (h.block_size < 1024 || h.block_size > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "Invalid bitstream, incorrect block size");
    uint32_t sz_mask = br.read(2);
    h.input_size = 0;
    if (sz_mask == 3) h.input_size = (int64_t)br.read(16) << 32;
    if (sz_mask >= 2) h.input_size |= br.read(32);
    if (sz_mask == 1) h.input_size = br.read(16);
    br.read(15);
    uint32_t ck1 = br.read(24);
    if (br.overrun) return ctx->fail(KZ_ERR_INVALID_FILE, "Invalid bitstream: truncated header");
    if (ck1 != header_checksum(ck_size, h.etype, h.t48, h.block_size, sz_mask, h.input_size))
        return ctx->fail(KZ_ERR_CRC_CHECK, "Invalid bitstream: checksum mismatch");
    h.bits = br.pos;
    return 0;
}

int count_transforms(uint64_t t48, uint64_t* ids) {
    int n = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (t48 >> (42 - 6 * i)) & 63;
        if (t != 0) ids[n++] = t;
    }
    if (n == 0) ids[n++] = 0;  // Factory.New :73-76 keeps one NONE
    return n;
}

// The data type encodingTask.encode derives from the first four bytes of a block before the transforms run
// (io/CompressedStream.go:806-816 over internal/Magic.go:73-112 GetMagicType and :116-206 IsDataCompressed / IsDataMultimedia /
// IsDataExecutable): 7 = DT_BIN for compressed containers, 2 = DT_MULTIMEDIA, 3 = DT_EXE, 0 = DT_UNDEFINED otherwise.
int32_t magic_data_type(const uint8_t* p, size_t n) {
    if (n < 4) return 0;
    const uint32_t key = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    const uint32_t k24 = key >> 8, k16 = key >> 16;
    // compressed: JPEG (APP0 marker only: the other FFD8FFEx keys match no class), GIF, PNG, 7z, zstd, brotli, CAB, ZIP, FLAC, xz, KANZ, RAR,
    // bzip2 / ID3 (24 bits), gzip (16 bits)
    switch (key) {
        case 0xFFD8FFE0u: case 0x47494638u: case 0x89504E47u: case 0x377ABCAFu: case 0x28B52FFDu: case 0x81CFB2CEu: case 0x4D534346u:
        case 0x504B0304u: case 0x664C6143u: case 0xFD377A58u: case 0x4B414E5Au: case 0x52617221u:
            return 7;
        default: break;
    }
    if ((key & ~0x0Fu) == 0xFFD8FFE0u) return 0;  // GetMagicType returns the key itself: no class matches
    if (k24 == 0x425A68u || k24 == 0x494433u) return 7;
    if (key == 0x25504446u) return 0;  // PDF: recognised, in no class
    if (key == 0x52494646u) return 2;  // RIFF
    if (key == 0x7F454C46u || key == 0xFEEDFACEu || key == 0xCEFAEDFEu || key == 0xFEEDFACFu || key == 0xCFFAEDFEu) return 3;  // ELF, Mach-O
    if (k16 == 0x1F8Bu) return 7;  // gzip
    if (k16 == 0x424Du) return 2;  // BMP
    if (k16 == 0x4D5Au) return 3;  // MZ
    if (k16 == 0x5034u || k16 == 0x5035u || k16 == 0x5036u) {  // binary PBM / PGM / PPM: "P4".."P6" followed by white space
        const uint32_t sub = (key >> 8) & 0xFF;
        if (sub == 0x07 || sub == 0x0A || sub == 0x0D || sub == 0x20) return 2;
    }
    return 0;
}

struct TransformPlan {
    int nt = 0;
    uint64_t ids[8];
    bool none() const { return nt == 1 && ids[0] == KZ_T_NONE; }
    bool bwt() const { return nt == 1 && ids[0] == KZ_T_BWT; }
    bool lz() const { return nt == 1 && (ids[0] == KZ_T_LZ || ids[0] == KZ_T_LZX); }
    bool lzx() const { return nt == 1 && ids[0] == KZ_T_LZX; }
    static bool seq_id(uint64_t t) {
        return t == KZ_T_BWT || t == KZ_T_BWTS || t == KZ_T_RANK || t == KZ_T_MTFT || t == KZ_T_ZRLT || t == KZ_T_ROLZ || t == KZ_T_PACK || t == KZ_T_DNA || t == KZ_T_LZ ||
               t == KZ_T_LZX || t == KZ_T_MM || t == KZ_T_TEXT || t == KZ_T_UTF || t == KZ_T_EXE || t == KZ_T_EXE;
    }
    // any sequence of BWT / BWTS / RANK / MTFT / ZRLT / ROLZ stages (e.g. "BWT+RANK+ZRLT", the transform chain of kanzi -l 5)
    // stages that read ctx["dataType"] in a way the block's magic number can change (LZ / LZX only look for DNA / small alphabets)
    bool uses_magic() const {
        for (int i = 0; i < nt; i++)
            if (ids[i] == KZ_T_ROLZ || ids[i] == KZ_T_PACK || ids[i] == KZ_T_DNA || ids[i] == KZ_T_MM || ids[i] == KZ_T_TEXT || ids[i] == KZ_T_UTF || ids[i] == KZ_T_EXE) return true;
        return false;
    }
    bool generic() const {
        if (bwt() || lz()) return false;  // single BWT / LZ / LZX keep their dedicated paths
        for (int i = 0; i < nt; i++)
            if (!seq_id(ids[i])) return false;
        return true;
    }
};

// The stages added last (PACK / DNA / MM, LZ / LZX inside sequences, the magic-number data type, TEXT) can be switched off with
// KZ_EXPERIMENTAL=0: the library then answers "not available" for them, as it did before they existed.
bool experimental_stages() {
    const char* e = getenv("KZ_EXPERIMENTAL");
    return !(e && e[0] == '0');
}

// etype: the entropy codec of the stream. The reference picks the TEXT encoding from it (transform/Factory.go:99-119): encoding 2 for NONE /
// ANS0 / HUFFMAN / RANGE, encoding 1 otherwise; only encoding 2 (textCodec2) exists here, so TEXT with any other codec is refused
// rather than written in a form the reference could not read.
int plan_transforms(kz_ctx* ctx, uint64_t t48, uint32_t etype, TransformPlan& p, int err_code) {
    p.nt = count_transforms(t48, p.ids);
    for (int i = 0; i < p.nt; i++)
        if (p.ids[i] == KZ_T_TEXT && !(etype == KZ_E_NONE || etype == KZ_E_ANS0 || etype == KZ_E_HUFFMAN || etype == KZ_E_RANGE))
            return ctx->fail(err_code, "TEXT with this entropy codec selects the reference's text encoding 1, which the GPU path does not implement (NONE / ANS0 / HUFFMAN / RANGE select encoding 2)");
    if (!experimental_stages()) {
        for (int i = 0; i < p.nt; i++) {
            const uint64_t t = p.ids[i];
            const bool newer = t == KZ_T_PACK || t == KZ_T_DNA || t == KZ_T_MM || t == KZ_T_TEXT || t == KZ_T_UTF || t == KZ_T_EXE || ((t == KZ_T_LZ || t == KZ_T_LZX) && p.nt > 1);
            if (newer) return ctx->fail(err_code, "transform stage switched off by KZ_EXPERIMENTAL=0: PACK / DNA / MM / TEXT, LZ / LZX inside a sequence");
        }
    }
    for (int i = 0; i < p.nt; i++)
        if (p.ids[i] == KZ_T_TEXT && !text_available())
            return ctx->fail(err_code, "TEXT: the library was built without the static dictionary (kanzi-go_b200/gen_text_dict.py)");
    if (!(p.none() || p.bwt() || p.lz() || p.generic()))
        return ctx->fail(err_code, "transform sequence not available on the GPU path (sequences of BWT / BWTS / LZ / LZX / ROLZ / RANK / MTFT / ZRLT / PACK / DNA are)");
    return 0;
}


// ------------------------------------------------------------------------------------------------------------
// Generic transform sequences (ByteTransformSequence, transform/Sequence.go:64-186) over the stages that exist as
// kernels: BWT, RANK, MTFT, ZRLT. Every stage processes all blocks of the batch; a stage that declines a block
// ("skip", Sequence.go:100-105) leaves its skip bit set and the block moves on unchanged.
// ------------------------------------------------------------------------------------------------------------
// One forward stage. Block b: d_in + b*istride (len[b] bytes) -> d_out + b*ostride. applied[b] = 1 and len[b] updated on success.
int apply_forward(kz_ctx* ctx, uint64_t t, const uint8_t* d_in, uint64_t istride, uint8_t* d_out, uint64_t ostride, std::vector<uint32_t>& len,
                  const std::vector<uint8_t>& active, std::vector<uint8_t>& applied, std::vector<int32_t>* dts = nullptr) {
    const uint32_t nblocks = (uint32_t)len.size();
    applied.assign(nblocks, 0);
    if (t == KZ_T_BWT) {
        CK(ctx->d_lens.ensure((size_t)nblocks * 4 + 64));
        uint32_t* d_post = ctx->d_lens.as<uint32_t>();
        static const bool one_by_one = [] {  // KZ_BWT_FWD=single: one suffix sort per block (round 1)
            const char* e = getenv("KZ_BWT_FWD");
            return e && e[0] == 's' && e[1] == 'i';
        }();
        if (one_by_one) {
            uint32_t max_len = 0;
            for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
            CK(ctx->d_ws.ensure(bwt_forward_workspace(max_len)));
            for (uint32_t b = 0; b < nblocks; b++) {
                if (!active[b] || len[b] < 2) continue;
                LaunchScope ls(ctx, "bwt_forward");
                cudaError_t e = bwt_forward_device(d_in + b * istride, len[b], d_out + b * ostride, d_post + b, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream,
                                                   &ctx->launches);
                if (e == cudaSuccess) applied[b] = 1;
                else if (e != cudaErrorInvalidValue) return ctx->cuda_fail(e, "bwt_forward");
            }
        } else {
            // all blocks of a sub-batch are suffix-sorted together (kz_bwt.cu: bwt_forward_batch)
            for (uint32_t b0 = 0; b0 < nblocks;) {
                std::vector<uint64_t> so, dof;
                std::vector<uint32_t> nn, num;
                uint64_t total = 0;
                uint32_t b = b0;
                for (; b < nblocks; b++) {
                    if (!active[b] || len[b] < 2) continue;
                    if (len[b] > bwt_forward_batch_max_total()) {  // a block the batch's key layout cannot hold (1 GiB): its own suffix sort
                        if (!nn.empty()) break;                    // flush what has been collected first
                        CK(ctx->d_ws.ensure(bwt_forward_workspace(len[b])));
                        LaunchScope ls(ctx, "bwt_forward");
                        cudaError_t e = bwt_forward_device(d_in + b * istride, len[b], d_out + b * ostride, d_post + b, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                           ctx->stream, &ctx->launches);
                        if (e == cudaSuccess) applied[b] = 1;
                        else if (e != cudaErrorInvalidValue) return ctx->cuda_fail(e, "bwt_forward");
                        continue;
                    }
                    if (!nn.empty() && (total + len[b] > bwt_forward_batch_max_total() || nn.size() >= bwt_forward_batch_max_blocks())) break;
                    so.push_back(b * istride), dof.push_back(b * ostride), nn.push_back(len[b]), num.push_back(b);
                    total += len[b];
                }
                b0 = b;
                if (nn.empty()) continue;
                CK(ctx->d_ws.ensure(bwt_forward_batch_workspace(total, (uint32_t)nn.size())));
                LaunchScope ls(ctx, "bwt_forward");
                cudaError_t e = bwt_forward_batch(d_in, d_out, so.data(), dof.data(), nn.data(), num.data(), (uint32_t)nn.size(), d_post, ctx->d_ws.as<uint8_t>(),
                                                  ctx->d_ws.cap, ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_forward");
                for (uint32_t k : num) applied[k] = 1;
            }
        }
        std::vector<uint32_t> post(nblocks);
        CK(cudaMemcpyAsync(post.data(), d_post, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++)
            if (applied[b]) len[b] = post[b];
    } else if (t == KZ_T_BWTS) {  // BWTS.Forward: n bytes -> n bytes, never declines
        uint32_t max_len = 0;
        for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
        CK(ctx->d_ws.ensure(bwts_forward_workspace(max_len)));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            LaunchScope ls(ctx, "bwts_forward");
            cudaError_t e = bwts_forward_device(d_in + b * istride, len[b], d_out + b * ostride, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream,
                                                &ctx->launches);
            if (e != cudaSuccess) return ctx->cuda_fail(e, "bwts_forward");
            applied[b] = 1;
        }
    } else if (t == KZ_T_ROLZ) {  // ROLZCodec.Forward: declines small blocks and blocks it cannot shrink
        uint32_t max_len = 0;
        for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
        const size_t sstride = rolz_scratch_bytes(max_len);
        CK(ctx->d_ws.ensure(sstride * nblocks + 256));
        std::vector<RolzBlock> rb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            rb[b].src_off = b * istride;
            rb[b].dst_off = b * ostride;
            rb[b].scratch_off = b * sstride;
            rb[b].len = active[b] ? len[b] : 0;
            rb[b].cap = (uint32_t)std::min<uint64_t>(ostride, 0xFFFFFFF0u);
            rb[b].data_type = dts ? (*dts)[b] : 0;
            rb[b].pad = 0;
        }
        Packer pk;
        const size_t o_rb = pk.add(rb.data(), rb.size() * sizeof(RolzBlock));
        const size_t o_res = pk.reserve((size_t)nblocks * sizeof(RolzResult));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            LaunchScope ls(ctx, "rolz_forward");
            rolz_forward_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, (const RolzBlock*)(T + o_rb), (int)nblocks, ctx->d_ws.as<uint8_t>(), d_out,
                                                                (RolzResult*)(T + o_res));
        }
        CK(cudaGetLastError());
        std::vector<RolzResult> hr(nblocks);
        CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nblocks * sizeof(RolzResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (active[b] && len[b] && hr[b].status == 0) {
                applied[b] = 1;
                len[b] = hr[b].out_len;
            }
            if (dts && active[b] && len[b]) (*dts)[b] = hr[b].data_type;  // ctx["dataType"] after the call
        }
    } else if (t == KZ_T_LZ || t == KZ_T_LZX) {  // LZCodec as a stage of a sequence (e.g. "DNA+LZ" = kanzi -l 2)
        const bool extra = t == KZ_T_LZX;
        uint32_t max_len = 0;
        for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
        if ((uint64_t)max_len + max_len / 64 + 128 > ostride) return ctx->fail(KZ_ERR_UNKNOWN, "internal: stage buffer too small for LZ");
        const uint64_t sstride = (lz_scratch_bytes(max_len) + 255) & ~size_t(255);
        const size_t hbytes = ((size_t)nblocks << (extra ? 19 : 16)) * 4;
        std::vector<LzBlock> lb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            lb[b].src_off = b * istride;
            lb[b].dst_off = b * ostride;
            lb[b].scratch_off = (uint64_t)b * sstride;
            lb[b].len = active[b] ? len[b] : 0;
            lb[b].cap = 0;
            lb[b].data_type = dts ? (*dts)[b] : 0;
            lb[b].pad = 0;
        }
        std::vector<LzResult> hres;
        {
            int rc = lz_forward_batch(ctx, d_in, d_out, lb, extra, hres);
            if (rc) return rc;
        }
        for (uint32_t b = 0; b < nblocks; b++)
            if (active[b] && len[b] && hres[b].status == 0) {
                applied[b] = 1;
                len[b] = hres[b].out_len;
            }
    } else if (t == KZ_T_TEXT) {  // TextCodec (encoding 2): declines unless the block looks like text
        if (!text_available()) return ctx->fail(KZ_ERR_CREATE_CODEC, "TEXT: the library was built without the static dictionary");
        const uint32_t SUB = 64;  // ~9 MB of dictionary per block
        const uint64_t sbs = ctx->stream_bs ? ctx->stream_bs : *std::max_element(len.begin(), len.end());
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            uint32_t max_len = 0;
            for (uint32_t b = b0; b < b0 + nb; b++) max_len = std::max(max_len, active[b] ? len[b] : 0u);
            if (max_len == 0) continue;
            CK(ctx->d_ws.ensure(text_workspace(nb, sbs, max_len)));
            std::vector<TextBlock> tb(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                tb[k].src_off = b * istride;
                tb[k].dst_off = b * ostride;
                tb[k].len = active[b] ? len[b] : 0;
                tb[k].cap = len[b];  // MaxEncodedLen = srcLen (:1720-1724)
                tb[k].data_type = dts ? (*dts)[b] : 0;
                tb[k].pad = 0;
            }
            Packer pk;
            const size_t o_tb = pk.add(tb.data(), tb.size() * sizeof(TextBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(TextResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "text_forward");
                cudaError_t e = text_forward_batch(d_in, d_out, (const TextBlock*)(T + o_tb), tb.data(), nb, max_len, sbs, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                   (TextResult*)(T + o_res), ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "text_forward");
            }
            std::vector<TextResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(TextResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (dts) (*dts)[b] = hr[k].data_type;
                if (hr[k].status == 0) {
                    applied[b] = 1;
                    len[b] = hr[k].out_len;
                }
            }
        }
    } else if (t == KZ_T_UTF) {  // UTFCodec.Forward: declines unless the block is UTF-8 with enough multi-byte sequences
        const uint32_t SUB = 32;  // 16 MB of code point counters per block
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            bool any = false;
            for (uint32_t b = b0; b < b0 + nb; b++) any = any || (active[b] && len[b]);
            if (!any) continue;
            CK(ctx->d_ws.ensure(utf_workspace(nb)));
            std::vector<UtfBlock> ub(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                ub[k].src_off = b * istride;
                ub[k].dst_off = b * ostride;
                ub[k].len = active[b] ? len[b] : 0;
                ub[k].cap = (uint32_t)std::min<uint64_t>((uint64_t)len[b] + 8192, 0xFFFFFFF0u);  // MaxEncodedLen (:386-388)
                ub[k].data_type = dts ? (*dts)[b] : 0;
                ub[k].pad = 0;
            }
            Packer pk;
            const size_t o_ub = pk.add(ub.data(), ub.size() * sizeof(UtfBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(UtfResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "utf_forward");
                cudaError_t e = utf_forward_batch(d_in, d_out, (const UtfBlock*)(T + o_ub), nb, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, (UtfResult*)(T + o_res),
                                                  ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "utf_forward");
            }
            std::vector<UtfResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(UtfResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (dts) (*dts)[b] = hr[k].data_type;
                if (hr[k].status == 0) {
                    applied[b] = 1;
                    len[b] = hr[k].out_len;
                }
            }
        }
    } else if (t == KZ_T_EXE) {  // EXECodec.Forward: declines unless the block is x86 / ARM64 code
        std::vector<ExeBlock> eb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            const uint64_t mel = len[b] <= 256 ? (uint64_t)len[b] + 32 : (uint64_t)len[b] + len[b] / 8;  // MaxEncodedLen (:701-708)
            if (active[b] && mel > ostride) return ctx->fail(KZ_ERR_UNKNOWN, "internal: stage buffer too small for EXE");
            eb[b].src_off = b * istride;
            eb[b].dst_off = b * ostride;
            eb[b].len = active[b] ? len[b] : 0;
            eb[b].cap = (uint32_t)std::min<uint64_t>(mel, 0xFFFFFFF0u);
            eb[b].data_type = dts ? (*dts)[b] : 0;
            eb[b].pad = 0;
        }
        Packer pk;
        const size_t o_eb = pk.add(eb.data(), eb.size() * sizeof(ExeBlock));
        const size_t o_res = pk.reserve((size_t)nblocks * sizeof(ExeResult));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            LaunchScope ls(ctx, "exe_forward");
            exe_forward_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, (const ExeBlock*)(T + o_eb), (int)nblocks, d_out, (ExeResult*)(T + o_res));
            ctx->launches++;
        }
        CK(cudaGetLastError());
        std::vector<ExeResult> hr(nblocks);
        CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nblocks * sizeof(ExeResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (hr[b].status < 0) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "EXE codec: malformed executable header (the reference panics on this block)");
            if (dts) (*dts)[b] = hr[b].data_type;
            if (hr[b].status == 0) {
                applied[b] = 1;
                len[b] = hr[b].out_len;
            }
        }
    } else if (t == KZ_T_MM) {  // FSDCodec.Forward: declines unless a fixed-step delta lowers the sampled entropy
        const uint32_t SUB = 1024;
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            uint32_t max_len = 0;
            for (uint32_t b = b0; b < b0 + nb; b++) max_len = std::max(max_len, active[b] ? len[b] : 0u);
            if (max_len == 0) continue;
            if ((uint64_t)fsd_max_encoded_len(max_len) > ostride) return ctx->fail(KZ_ERR_UNKNOWN, "internal: stage buffer too small for MM");
            CK(ctx->d_ws.ensure(fsd_workspace(nb, max_len)));
            std::vector<FsdBlock> fb(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                fb[k].src_off = b * istride;
                fb[k].dst_off = b * ostride;
                fb[k].len = active[b] ? len[b] : 0;
                fb[k].cap = fsd_max_encoded_len(len[b]);  // the sequence sizes dst with MaxEncodedLen (Sequence.go:81-93)
                fb[k].data_type = dts ? (*dts)[b] : 0;
                fb[k].pad = 0;
            }
            Packer pk;
            const size_t o_fb = pk.add(fb.data(), fb.size() * sizeof(FsdBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(FsdResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "fsd_forward");
                cudaError_t e = fsd_forward_batch(d_in, d_out, (const FsdBlock*)(T + o_fb), nb, max_len, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                  (FsdResult*)(T + o_res), ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "fsd_forward");
            }
            std::vector<FsdResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(FsdResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (dts) (*dts)[b] = hr[k].data_type;
                if (hr[k].status == 0) {
                    applied[b] = 1;
                    len[b] = hr[k].out_len;
                }
            }
        }
    } else if (t == KZ_T_PACK || t == KZ_T_DNA) {  // AliasCodec.Forward: bit packing of small alphabets / digram aliases; often declines
        const uint32_t SUB = 256;  // blocks per launch sequence (bounds the workspace: ~390 KB per block)
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            uint32_t max_len = 0;
            for (uint32_t b = b0; b < b0 + nb; b++) max_len = std::max(max_len, active[b] ? len[b] : 0u);
            if (max_len == 0) continue;
            CK(ctx->d_ws.ensure(alias_workspace(nb, max_len)));
            std::vector<AliasBlock> ab(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                ab[k].src_off = b * istride;
                ab[k].dst_off = b * ostride;
                ab[k].len = active[b] ? len[b] : 0;
                ab[k].cap = (uint32_t)std::min<uint64_t>((uint64_t)len[b] + 1024, 0xFFFFFFF0u);  // the sequence sizes dst with MaxEncodedLen (Sequence.go:81-93)
                ab[k].data_type = dts ? (*dts)[b] : 0;
                ab[k].only_dna = t == KZ_T_DNA ? 1u : 0u;
            }
            Packer pk;
            const size_t o_ab = pk.add(ab.data(), ab.size() * sizeof(AliasBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(AliasResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "alias_forward");
                cudaError_t e = alias_forward_batch(d_in, d_out, (const AliasBlock*)(T + o_ab), nb, max_len, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                    (AliasResult*)(T + o_res), ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "alias_forward");
            }
            std::vector<AliasResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(AliasResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (dts) (*dts)[b] = hr[k].data_type;
                if (hr[k].status == 0) {
                    applied[b] = 1;
                    len[b] = hr[k].out_len;
                }
            }
        }
    } else if (t == KZ_T_RANK || t == KZ_T_MTFT) {  // SBRT.Forward never declines (buffers are sized for MaxEncodedLen)
        std::vector<SbrtBlock> sb(nblocks);
        std::vector<uint32_t> seg_block;
        uint32_t max_segs = 0;
        for (uint32_t b = 0; b < nblocks; b++) {
            sb[b].src_off = b * istride;
            sb[b].dst_off = b * ostride;
            sb[b].len = len[b];
            sb[b].seg_base = (uint32_t)seg_block.size();
            sb[b].active = active[b] && len[b] > 0;
            sb[b].pad = 0;
            const uint32_t nseg = sb[b].active ? (len[b] + SBRT_SEG_BYTES - 1) / SBRT_SEG_BYTES : 0;
            max_segs = std::max(max_segs, nseg);
            for (uint32_t s = 0; s < nseg; s++) seg_block.push_back(b);
            applied[b] = sb[b].active ? 1 : 0;
        }
        const uint32_t nsegs = (uint32_t)seg_block.size();
        if (nsegs) {
            Packer pk;
            const size_t o_sb = pk.add(sb.data(), sb.size() * sizeof(SbrtBlock));
            const size_t o_map = pk.add(seg_block.data(), seg_block.size() * 4);
            int rc = upload(ctx, pk);
            if (rc) return rc;
            CK(ctx->d_seg.ensure((size_t)nsegs * 512 * 4 + 64));
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            const SbrtBlock* d_sb = (const SbrtBlock*)(T + o_sb);
            const int mode = t == KZ_T_MTFT ? 1 : 2;
            {
                LaunchScope ls(ctx, "sbrt_last2");
                sbrt_last2_kernel<<<dim3(max_segs, nblocks), 256, 0, ctx->stream>>>(d_in, d_sb, ctx->d_seg.as<int32_t>());
            }
            {
                LaunchScope ls(ctx, "sbrt_carry");
                sbrt_carry_kernel<<<nblocks, 256, 0, ctx->stream>>>(d_sb, ctx->d_seg.as<int32_t>());
            }
            {
                LaunchScope ls(ctx, "sbrt_rank");
                sbrt_rank_kernel<<<(nsegs + 3) / 4, 128, 0, ctx->stream>>>(d_in, d_sb, (int)nblocks, mode, (const uint32_t*)(T + o_map), nsegs,
                                                                          ctx->d_seg.as<int32_t>(), d_out);
            }
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(ctx->stream));  // the tables live in d_tables, which the next stage reuses
        }
    } else if (t == KZ_T_ZRLT) {
        std::vector<ZrltBlock> zb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            zb[b].src_off = b * istride;
            zb[b].dst_off = b * ostride;
            zb[b].len = len[b];
            zb[b].cap = len[b];
            zb[b].active = active[b] && len[b] > 0;
            zb[b].pad = 0;
        }
        Packer pk;
        const size_t o_zb = pk.add(zb.data(), zb.size() * sizeof(ZrltBlock));
        const size_t o_len = pk.reserve((size_t)nblocks * 4);
        const size_t o_st = pk.reserve((size_t)nblocks * 4);
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            LaunchScope ls(ctx, "zrlt_forward");
            zrlt_forward_kernel<<<nblocks, 512, 0, ctx->stream>>>(d_in, (const ZrltBlock*)(T + o_zb), (int)nblocks, d_out, (uint32_t*)(T + o_len),
                                                                  (int32_t*)(T + o_st));
        }
        CK(cudaGetLastError());
        std::vector<uint32_t> hl(nblocks);
        std::vector<int32_t> hs(nblocks);
        CK(cudaMemcpyAsync(hl.data(), T + o_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hs.data(), T + o_st, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (zb[b].active && hs[b] == 0) {
                applied[b] = 1;
                len[b] = hl[b];
            }
        }
    } else {
        return ctx->fail(KZ_ERR_CREATE_CODEC, "transform not available on the GPU path");
    }
    return 0;
}

// ByteTransformSequence.Forward for every block of a batch. jobs[] must be initialised (copy flags, lengths).
int forward_generic(kz_ctx* ctx, const TransformPlan& plan, const uint8_t* d_src, uint64_t stride, const std::vector<uint32_t>& blen, const uint8_t** d_data,
                    std::vector<EncJob>& jobs, std::vector<int32_t>& dts) {
    const uint32_t nblocks = (uint32_t)blen.size();
    uint32_t max_len = 0;
    for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, blen[b]);
    // room for the largest stage output: +33 bytes per BWT / SBRT stage, n/64 + 128 for LZ, n/16 for MM, headers of PACK
    const uint64_t gstride = ((uint64_t)max_len + max_len / 8 + 33ull * plan.nt + 8192 + 1024 + 256 + 15) & ~15ull;
    CK(ctx->d_tmp.ensure(gstride * nblocks + 64));
    CK(ctx->d_tmp2.ensure(gstride * nblocks + 64));
    std::vector<uint32_t> len(blen);
    std::vector<uint8_t> active(nblocks), applied;
    std::vector<uint8_t> flags(nblocks, 0xFF);
    for (uint32_t b = 0; b < nblocks; b++) active[b] = !jobs[b].copy && blen[b] > 0;
    const uint8_t* cur = d_src;
    uint64_t cur_stride = stride;
    for (int i = 0; i < plan.nt; i++) {
        uint8_t* outb = (i & 1) ? ctx->d_tmp2.as<uint8_t>() : ctx->d_tmp.as<uint8_t>();
        uint64_t bytes_in = 0, bytes_out = 0;
        for (uint32_t b = 0; b < nblocks; b++) bytes_in += active[b] ? len[b] : 0;
        int rc = apply_forward(ctx, plan.ids[i], cur, cur_stride, outb, gstride, len, active, applied, &dts);
        if (rc) return rc;
        for (uint32_t b = 0; b < nblocks; b++) bytes_out += active[b] ? len[b] : 0;
        note_stage_bytes(ctx, "fwd", plan.ids[i], bytes_in, bytes_out);
        for (uint32_t b = 0; b < nblocks; b++) {
            if (applied[b]) flags[b] &= (uint8_t)~(1u << (7 - i));
            else if (len[b]) CK(cudaMemcpyAsync(outb + b * gstride, cur + b * cur_stride, len[b], cudaMemcpyDeviceToDevice, ctx->stream));
        }
        cur = outb;
        cur_stride = gstride;
    }
    for (uint32_t b = 0; b < nblocks; b++) {
        EncJob& j = jobs[b];
        j.data_off = b * gstride;
        j.post_len = len[b];
        if (!j.copy) j.skip_flags = flags[b];
    }
    *d_data = cur;
    return 0;
}

// One inverse stage, mirror of apply_forward. cap = capacity of a destination block.
int apply_inverse(kz_ctx* ctx, uint64_t t, const uint8_t* d_in, uint8_t* d_out, uint64_t stride, uint32_t cap, std::vector<uint32_t>& len,
                  const std::vector<uint8_t>& active) {
    const uint32_t nblocks = (uint32_t)len.size();
    if (t == KZ_T_BWT) {
        std::vector<uint8_t> heads((size_t)nblocks * 64, 0);
        CK(cudaMemcpy2DAsync(heads.data(), 64, d_in, stride, 33, nblocks, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        static const bool one_by_one = [] {  // KZ_BWT_INV=single: one set of launches per block (round 1)
            const char* e = getenv("KZ_BWT_INV");
            return e && e[0] == 's';
        }();
        if (one_by_one) {
            uint32_t max_len = 0;
            for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
            CK(ctx->d_ws.ensure(bwt_inverse_workspace(max_len)));
            for (uint32_t b = 0; b < nblocks; b++) {
                if (!active[b]) continue;
                uint32_t produced = 0;
                LaunchScope ls(ctx, "bwt_inverse");
                cudaError_t e = bwt_inverse_device(d_in + b * stride, &heads[(size_t)b * 64], len[b], d_out + b * stride, cap, &produced, ctx->d_ws.as<uint8_t>(),
                                                   ctx->d_ws.cap, ctx->stream, &ctx->launches);
                if (e == cudaErrorInvalidValue) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header");
                if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_inverse");
                len[b] = produced;
            }
        } else {
            for (uint32_t b0 = 0; b0 < nblocks;) {
                std::vector<uint64_t> so, dof;
                std::vector<uint32_t> ll, num;
                std::vector<uint8_t> hh;
                uint64_t total = 0;
                uint32_t b = b0;
                for (; b < nblocks; b++) {
                    if (!active[b]) continue;
                    if (!ll.empty() && (total + len[b] > bwt_inverse_batch_max_total() || ll.size() >= bwt_inverse_batch_max_blocks())) break;
                    so.push_back(b * stride), dof.push_back(b * stride), ll.push_back(len[b]), num.push_back(b);
                    hh.insert(hh.end(), heads.begin() + (size_t)b * 64, heads.begin() + (size_t)(b + 1) * 64);
                    total += len[b];
                }
                b0 = b;
                if (ll.empty()) continue;
                CK(ctx->d_ws.ensure(bwt_inverse_batch_workspace(total, (uint32_t)ll.size())));
                std::vector<uint32_t> produced(ll.size(), 0);
                LaunchScope ls(ctx, "bwt_inverse");
                cudaError_t e = bwt_inverse_batch(d_in, d_out, so.data(), dof.data(), ll.data(), hh.data(), (uint32_t)ll.size(), cap, produced.data(),
                                                  ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream, &ctx->launches);
                if (e == cudaErrorInvalidValue) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header or data");
                if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_inverse");
                for (size_t k = 0; k < num.size(); k++) len[num[k]] = produced[k];
            }
        }
    } else if (t == KZ_T_BWTS) {
        uint32_t max_len = 0;
        for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, len[b]);
        CK(ctx->d_ws.ensure(bwts_inverse_workspace(max_len)));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (len[b] > cap) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWTS inverse transform failed: output buffer too small");
            LaunchScope ls(ctx, "bwts_inverse");
            cudaError_t e = bwts_inverse_device(d_in + b * stride, len[b], d_out + b * stride, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream, &ctx->launches);
            if (e != cudaSuccess) return ctx->cuda_fail(e, "bwts_inverse");
        }
    } else if (t == KZ_T_ROLZ) {
        const size_t sstride = rolz_scratch_bytes(cap);
        CK(ctx->d_ws.ensure(sstride * nblocks + 256));
        std::vector<RolzBlock> rb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            rb[b].src_off = b * stride;
            rb[b].dst_off = b * stride;
            rb[b].scratch_off = b * sstride;
            rb[b].len = active[b] ? len[b] : 0;
            rb[b].cap = cap;
            rb[b].data_type = 0;
            rb[b].pad = 0;
        }
        Packer pk;
        const size_t o_rb = pk.add(rb.data(), rb.size() * sizeof(RolzBlock));
        const size_t o_res = pk.reserve((size_t)nblocks * sizeof(RolzResult));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            LaunchScope ls(ctx, "rolz_inverse");
            rolz_inverse_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, (const RolzBlock*)(T + o_rb), (int)nblocks, ctx->d_ws.as<uint8_t>(), d_out,
                                                                (RolzResult*)(T + o_res));
        }
        CK(cudaGetLastError());
        std::vector<RolzResult> hr(nblocks);
        CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nblocks * sizeof(RolzResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (hr[b].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "ROLZ codec inverse transform failed: invalid data");
            len[b] = hr[b].out_len;
        }
    } else if (t == KZ_T_LZ || t == KZ_T_LZX) {
        std::vector<LzBlock> lb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            lb[b].src_off = b * stride;
            lb[b].dst_off = b * stride;
            lb[b].scratch_off = 0;
            lb[b].len = active[b] ? len[b] : 0;
            lb[b].cap = cap;
            lb[b].data_type = 0;
            lb[b].pad = 0;
        }
        std::vector<uint32_t> hl(nblocks);
        std::vector<int32_t> hs(nblocks);
        static const bool serial_inverse = [] {  // KZ_LZ_INV=serial: the round-1 kernel (one warp per block copies match by match)
            const char* e = getenv("KZ_LZ_INV");
            return e && e[0] == 's';
        }();
        if (!serial_inverse) {
            CtxLzWorkspace W(ctx);
            CtxLzHook H(ctx);
            cudaError_t e = lz_inverse_parallel(d_in, d_out, lb, W, ctx->stream, hl, hs, &H, &ctx->launches);
            if (e != cudaSuccess) return ctx->cuda_fail(e, "lz_inverse");
        } else {
            Packer pk;
            const size_t o_lb = pk.add(lb.data(), lb.size() * sizeof(LzBlock));
            const size_t o_len = pk.reserve((size_t)nblocks * 4);
            const size_t o_st = pk.reserve((size_t)nblocks * 4);
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "lz_inverse");
                lz_inverse_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, (LzBlock*)(T + o_lb), (int)nblocks, d_out, (uint32_t*)(T + o_len), (int32_t*)(T + o_st));
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(hl.data(), T + o_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(hs.data(), T + o_st, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (hs[b]) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "LZCodec inverse transform failed");
            len[b] = hl[b];
        }
    } else if (t == KZ_T_TEXT) {
        if (!text_available()) return ctx->fail(KZ_ERR_CREATE_CODEC, "TEXT: the library was built without the static dictionary");
        const uint32_t SUB = 64;
        const uint64_t sbs = ctx->stream_bs ? ctx->stream_bs : cap;
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            bool any = false;
            for (uint32_t b = b0; b < b0 + nb; b++) any = any || (active[b] && len[b]);
            if (!any) continue;
            uint32_t longest = 0;
            for (uint32_t b = b0; b < b0 + nb; b++) longest = std::max(longest, active[b] ? len[b] : 0u);
            CK(ctx->d_ws.ensure(text_workspace_inverse(nb, sbs, longest)));
            std::vector<TextBlock> tb(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                tb[k].src_off = b * stride;
                tb[k].dst_off = b * stride;
                tb[k].len = active[b] ? len[b] : 0;
                tb[k].cap = cap;
                tb[k].data_type = 0;
                tb[k].pad = 0;
            }
            Packer pk;
            const size_t o_tb = pk.add(tb.data(), tb.size() * sizeof(TextBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(TextResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "text_inverse");
                cudaError_t e = text_inverse_batch(d_in, d_out, (const TextBlock*)(T + o_tb), tb.data(), nb, sbs, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                   (TextResult*)(T + o_res), ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "text_inverse");
            }
            std::vector<TextResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(TextResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (hr[k].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "Text transform failed: invalid data");
                len[b] = hr[k].out_len;
            }
        }
    } else if (t == KZ_T_UTF) {
        const uint32_t SUB = 32;
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            bool any = false;
            for (uint32_t b = b0; b < b0 + nb; b++) any = any || (active[b] && len[b]);
            if (!any) continue;
            CK(ctx->d_ws.ensure(utf_workspace(nb)));
            std::vector<UtfBlock> ub(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                ub[k].src_off = b * stride;
                ub[k].dst_off = b * stride;
                ub[k].len = active[b] ? len[b] : 0;
                ub[k].cap = cap;
                ub[k].data_type = 0;
                ub[k].pad = 0;
            }
            Packer pk;
            const size_t o_ub = pk.add(ub.data(), ub.size() * sizeof(UtfBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(UtfResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "utf_inverse");
                cudaError_t e = utf_inverse_batch(d_in, d_out, (const UtfBlock*)(T + o_ub), nb, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, (UtfResult*)(T + o_res),
                                                  ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "utf_inverse");
            }
            std::vector<UtfResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(UtfResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (hr[k].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "UTF inverse transform failed: invalid data");
                len[b] = hr[k].out_len;
            }
        }
    } else if (t == KZ_T_EXE) {
        std::vector<ExeBlock> eb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            eb[b].src_off = b * stride;
            eb[b].dst_off = b * stride;
            eb[b].len = active[b] ? len[b] : 0;
            eb[b].cap = cap;
            eb[b].data_type = 0;
            eb[b].pad = 0;
        }
        Packer pk;
        const size_t o_eb = pk.add(eb.data(), eb.size() * sizeof(ExeBlock));
        const size_t o_res = pk.reserve((size_t)nblocks * sizeof(ExeResult));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            LaunchScope ls(ctx, "exe_inverse");
            exe_inverse_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_in, (const ExeBlock*)(T + o_eb), (int)nblocks, d_out, (ExeResult*)(T + o_res));
            ctx->launches++;
        }
        CK(cudaGetLastError());
        std::vector<ExeResult> hr(nblocks);
        CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nblocks * sizeof(ExeResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (hr[b].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "ExeCodec inverse transform failed: invalid data");
            len[b] = hr[b].out_len;
        }
    } else if (t == KZ_T_MM) {
        std::vector<FsdBlock> fb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            fb[b].src_off = b * stride;
            fb[b].dst_off = b * stride;
            fb[b].len = active[b] ? len[b] : 0;
            fb[b].cap = cap;
            fb[b].data_type = 0;
            fb[b].pad = 0;
        }
        Packer pk;
        const size_t o_fb = pk.add(fb.data(), fb.size() * sizeof(FsdBlock));
        const size_t o_res = pk.reserve((size_t)nblocks * sizeof(FsdResult));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        {
            static const bool warp_per_block = [] {  // KZ_MM_INV=warp: one warp per block walks the rounds (first data-parallel version)
                const char* e = getenv("KZ_MM_INV");
                return e && e[0] == 'w';
            }();
            LaunchScope ls(ctx, "fsd_inverse");
            cudaError_t e;
            if (warp_per_block) {
                e = fsd_inverse_batch(d_in, d_out, (const FsdBlock*)(T + o_fb), nblocks, (FsdResult*)(T + o_res), ctx->stream, &ctx->launches);
            } else {
                uint64_t total_len = 0;
                for (uint32_t b = 0; b < nblocks; b++) total_len += fb[b].len;
                CK(ctx->d_ws.ensure(fsd_inverse_workspace(nblocks, total_len)));
                e = fsd_inverse_tiles(d_in, d_out, fb.data(), nblocks, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, (FsdResult*)(T + o_res), ctx->stream, &ctx->launches);
            }
            if (e != cudaSuccess) return ctx->cuda_fail(e, "fsd_inverse");
        }
        std::vector<FsdResult> hr(nblocks);
        CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nblocks * sizeof(FsdResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!active[b] || len[b] == 0) continue;
            if (hr[b].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "FSD inverse transform failed: invalid data");
            len[b] = hr[b].out_len;
        }
    } else if (t == KZ_T_PACK || t == KZ_T_DNA) {
        const uint32_t SUB = 256;
        for (uint32_t b0 = 0; b0 < nblocks; b0 += SUB) {
            const uint32_t nb = std::min(SUB, nblocks - b0);
            bool any = false;
            for (uint32_t b = b0; b < b0 + nb; b++) any = any || (active[b] && len[b]);
            if (!any) continue;
            CK(ctx->d_ws.ensure(alias_workspace(nb, cap)));
            std::vector<AliasBlock> ab(nb);
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                ab[k].src_off = b * stride;
                ab[k].dst_off = b * stride;
                ab[k].len = active[b] ? len[b] : 0;
                ab[k].cap = cap;
                ab[k].data_type = 0;
                ab[k].only_dna = 0;
            }
            Packer pk;
            const size_t o_ab = pk.add(ab.data(), ab.size() * sizeof(AliasBlock));
            const size_t o_res = pk.reserve((size_t)nb * sizeof(AliasResult));
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "alias_inverse");
                cudaError_t e = alias_inverse_batch(d_in, d_out, (const AliasBlock*)(T + o_ab), nb, cap, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                                    (AliasResult*)(T + o_res), ctx->stream, &ctx->launches);
                if (e != cudaSuccess) return ctx->cuda_fail(e, "alias_inverse");
            }
            std::vector<AliasResult> hr(nb);
            CK(cudaMemcpyAsync(hr.data(), T + o_res, (size_t)nb * sizeof(AliasResult), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t k = 0; k < nb; k++) {
                const uint32_t b = b0 + k;
                if (!active[b] || len[b] == 0) continue;
                if (hr[k].status) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "Alias codec inverse transform failed: invalid data");
                len[b] = hr[k].out_len;
            }
        }
    } else if (t == KZ_T_RANK || t == KZ_T_MTFT) {
        std::vector<SbrtBlock> sb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            sb[b].src_off = b * stride;
            sb[b].dst_off = b * stride;
            sb[b].len = len[b];
            sb[b].seg_base = 0;
            sb[b].active = active[b] && len[b] > 0;
            sb[b].pad = 0;
            if (sb[b].active && len[b] > cap) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "SBRT inverse transform failed: output buffer too small");
        }
        Packer pk;
        const size_t o_sb = pk.add(sb.data(), sb.size() * sizeof(SbrtBlock));
        int rc = upload(ctx, pk);
        if (rc) return rc;
        {
            LaunchScope ls(ctx, "sbrt_inverse");
            uint32_t longest = 0;
            for (uint32_t b = 0; b < nblocks; b++) longest = std::max(longest, len[b]);
            sbrt_inverse_launch_len(d_in, (const SbrtBlock*)(ctx->d_tables.as<uint8_t>() + o_sb), (int)nblocks, t == KZ_T_MTFT ? 1 : 2, d_out, longest, ctx->stream);
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    } else if (t == KZ_T_ZRLT) {
        std::vector<ZrltBlock> zb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            zb[b].src_off = b * stride;
            zb[b].dst_off = b * stride;
            zb[b].len = len[b];
            zb[b].cap = cap;
            zb[b].active = active[b] && len[b] > 0;
            zb[b].pad = 0;
        }
        Packer pk;
        const size_t o_zb = pk.add(zb.data(), zb.size() * sizeof(ZrltBlock));
        const size_t o_len = pk.reserve((size_t)nblocks * 4);
        const size_t o_st = pk.reserve((size_t)nblocks * 4);
        int rc = upload(ctx, pk);
        if (rc) return rc;
        uint8_t* T = ctx->d_tables.as<uint8_t>();
        CK(cudaMemsetAsync(d_out, 0, stride * nblocks, ctx->stream));  // runs of zeros are never written, only skipped
        {
            LaunchScope ls(ctx, "zrlt_inverse");
            zrlt_inverse_kernel<<<nblocks, 512, 0, ctx->stream>>>(d_in, (const ZrltBlock*)(T + o_zb), (int)nblocks, d_out, (uint32_t*)(T + o_len),
                                                                  (int32_t*)(T + o_st));
        }
        CK(cudaGetLastError());
        std::vector<uint32_t> hl(nblocks);
        std::vector<int32_t> hs(nblocks);
        CK(cudaMemcpyAsync(hl.data(), T + o_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hs.data(), T + o_st, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < nblocks; b++) {
            if (!zb[b].active) continue;
            if (hs[b]) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "ZRLT inverse transform failed: output buffer is too small");
            len[b] = hl[b];
        }
    } else {
        return ctx->fail(KZ_ERR_INVALID_CODEC, "transform not available on the GPU path");
    }
    return 0;
}

// ByteTransformSequence.Inverse for every block of a batch. Block b's pre_len bytes are at d_pre + b*pstride (a scratch
// buffer of this context with room for blk_cap bytes per block); decoded blocks go to d_dst back to back (compact) or at
// b * slot.
int inverse_generic(kz_ctx* ctx, const TransformPlan& plan, std::vector<DecJob>& jobs, uint8_t* d_pre, uint64_t pstride, uint32_t blk_cap, uint8_t* d_dst,
                    size_t cap, std::vector<uint32_t>& out_len, uint64_t* total, uint64_t slot, bool compact) {
    const uint32_t nblocks = (uint32_t)jobs.size();
    std::vector<uint32_t> len(nblocks);
    for (uint32_t b = 0; b < nblocks; b++) len[b] = jobs[b].pre_len;
    CK(ctx->d_tmp2.ensure(pstride * nblocks + 64));
    uint8_t* cur = d_pre;
    uint8_t* other = ctx->d_tmp2.as<uint8_t>();
    if (other == cur) return ctx->fail(KZ_ERR_UNKNOWN, "internal: scratch buffers alias");
    std::vector<uint8_t> active(nblocks);
    for (int i = plan.nt - 1; i >= 0; i--) {
        bool any = false;
        for (uint32_t b = 0; b < nblocks; b++) {
            active[b] = !jobs[b].copy && !(jobs[b].skip_flags & (1u << (7 - i))) && len[b] > 0;
            any = any || active[b];
        }
        if (!any) continue;
        uint64_t bytes_in = 0, bytes_out = 0;
        for (uint32_t b = 0; b < nblocks; b++) bytes_in += active[b] ? len[b] : 0;
        int rc = apply_inverse(ctx, plan.ids[i], cur, other, pstride, blk_cap, len, active);
        if (rc) return rc;
        for (uint32_t b = 0; b < nblocks; b++) bytes_out += active[b] ? len[b] : 0;
        note_stage_bytes(ctx, "inv", plan.ids[i], bytes_in, bytes_out);
        for (uint32_t b = 0; b < nblocks; b++)
            if (!active[b] && len[b]) CK(cudaMemcpyAsync(other + b * pstride, cur + b * pstride, len[b], cudaMemcpyDeviceToDevice, ctx->stream));
        std::swap(cur, other);
    }
    uint64_t off = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint64_t at = compact ? off : (uint64_t)b * slot;
        if (at + len[b] > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
        if (len[b]) CK(cudaMemcpyAsync(d_dst + at, cur + b * pstride, len[b], cudaMemcpyDeviceToDevice, ctx->stream));
        out_len[b] = len[b];
        off += len[b];
    }
    *total = off;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Forward transform stage (ByteTransformSequence.Forward, transform/Sequence.go:64-125) of every block of a batch.
// Block b is d_src[b*stride : +len_b]. On return *d_data is the base of the post-transform data and jobs[] describes
// every block for the entropy stage. NONE leaves the data in place; BWT writes header+BWT into ctx->d_tmp.
int forward_stage(kz_ctx* ctx, const TransformPlan& plan, const uint8_t* d_src, size_t n, uint32_t block_size, uint64_t stride, const uint32_t* lens,
                  uint32_t nblocks, const uint8_t** d_data, std::vector<EncJob>& jobs, const std::vector<int32_t>* init_dt = nullptr, bool allow_copy = true) {
    jobs.resize(nblocks);
    std::vector<uint32_t> blen(nblocks);
    uint32_t max_len = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        blen[b] = lens ? lens[b] : (uint32_t)std::min<uint64_t>(block_size, n - (uint64_t)b * block_size);
        max_len = std::max(max_len, blen[b]);
    }
    *d_data = d_src;
    for (uint32_t b = 0; b < nblocks; b++) {
        EncJob& j = jobs[b];
        j.data_off = (uint64_t)b * stride;
        j.post_len = blen[b];
        j.ntransforms = (uint32_t)plan.nt;
        j.skip_flags = 0x7F;  // NullTransform succeeds: bit 7 cleared (Sequence.go:65,107)
        j.copy = allow_copy && blen[b] <= SMALL_BLOCK;  // io/CompressedStream.go:773-776: transform and entropy forced to NONE
        j.checksum = 0;
    }
    if (plan.none() || nblocks == 0) return 0;
    // ctx["dataType"] of every block: the caller's value (from the block's magic number, encodingTask.encode :808-816), updated by the stages
    std::vector<int32_t> dts(nblocks, 0);
    if (init_dt) dts = *init_dt;
    if (plan.generic()) return forward_generic(ctx, plan, d_src, stride, blen, d_data, jobs, dts);
    if (plan.lz()) {
        // ---- LZ / LZX (LZXCodec.Forward): one serial parse per block, all blocks of the batch in parallel
        const bool extra = plan.lzx();
        const uint64_t tstride = ((uint64_t)max_len + max_len / 64 + 64 + 64 + 15) & ~15ull;
        const uint64_t sstride = (lz_scratch_bytes(max_len) + 255) & ~size_t(255);
        const size_t hbytes = ((size_t)nblocks << (extra ? 19 : 16)) * 4;
        CK(ctx->d_tmp.ensure(tstride * nblocks + 64));
        std::vector<LzBlock> lb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            lb[b].src_off = (uint64_t)b * stride;
            lb[b].dst_off = (uint64_t)b * tstride;
            lb[b].scratch_off = (uint64_t)b * sstride;
            lb[b].len = jobs[b].copy ? 0 : blen[b];  // COPY blocks: transform forced to NONE
            lb[b].cap = 0;
            lb[b].data_type = dts[b];
            lb[b].pad = 0;
        }
        std::vector<LzResult> hres;
        {
            int rc = lz_forward_batch(ctx, d_src, ctx->d_tmp.as<uint8_t>(), lb, extra, hres);
            if (rc) return rc;
        }
        for (uint32_t b = 0; b < nblocks; b++) {
            EncJob& j = jobs[b];
            j.data_off = (uint64_t)b * tstride;
            if (!j.copy && hres[b].status == 0) {
                j.post_len = hres[b].out_len;
                j.skip_flags = 0x7F;
            } else {
                if (!j.copy) j.skip_flags = 0xFF;
                if (blen[b]) CK(cudaMemcpyAsync(ctx->d_tmp.as<uint8_t>() + j.data_off, d_src + (uint64_t)b * stride, blen[b], cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        *d_data = ctx->d_tmp.as<uint8_t>();
        return 0;
    }
    // ---- BWT (BWTBlockCodec.Forward): the blocks of the batch are suffix-sorted together (kz_bwt.cu: bwt_forward_batch)
    const uint64_t tstride = ((uint64_t)max_len + 33 + 64 + 15) & ~15ull;
    CK(ctx->d_tmp.ensure(tstride * nblocks + 64));
    CK(ctx->d_lens.ensure((size_t)nblocks * 4 + 64));
    uint8_t* tmp = ctx->d_tmp.as<uint8_t>();
    uint32_t* d_post = ctx->d_lens.as<uint32_t>();
    std::vector<uint8_t> ok(nblocks, 0);
    for (uint32_t b0 = 0; b0 < nblocks;) {
        std::vector<uint64_t> so, dof;
        std::vector<uint32_t> nn, num;
        uint64_t total = 0;
        uint32_t b = b0;
        for (; b < nblocks; b++) {
            const uint8_t* src = d_src + (uint64_t)b * stride;
            uint8_t* dst = tmp + (uint64_t)b * tstride;
            if (jobs[b].copy || blen[b] < 2) {
                if (blen[b]) CK(cudaMemcpyAsync(dst, src, blen[b], cudaMemcpyDeviceToDevice, ctx->stream));  // transform skipped: data unchanged
                continue;
            }
            if (blen[b] > bwt_forward_batch_max_total()) {  // 1 GiB block: its own suffix sort
                if (!nn.empty()) break;
                CK(ctx->d_ws.ensure(bwt_forward_workspace(blen[b])));
                LaunchScope ls(ctx, "bwt_forward");
                cudaError_t e = bwt_forward_device(src, blen[b], dst, d_post + b, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream, &ctx->launches);
                if (e == cudaSuccess) ok[b] = 1;
                else if (e != cudaErrorInvalidValue) return ctx->cuda_fail(e, "bwt_forward");
                else CK(cudaMemcpyAsync(dst, src, blen[b], cudaMemcpyDeviceToDevice, ctx->stream));
                continue;
            }
            if (!nn.empty() && (total + blen[b] > bwt_forward_batch_max_total() || nn.size() >= bwt_forward_batch_max_blocks())) break;
            so.push_back((uint64_t)b * stride), dof.push_back((uint64_t)b * tstride), nn.push_back(blen[b]), num.push_back(b);
            total += blen[b];
        }
        b0 = b;
        if (nn.empty()) continue;
        CK(ctx->d_ws.ensure(bwt_forward_batch_workspace(total, (uint32_t)nn.size())));
        LaunchScope ls(ctx, "bwt_forward");
        cudaError_t e = bwt_forward_batch(d_src, tmp, so.data(), dof.data(), nn.data(), num.data(), (uint32_t)nn.size(), d_post, ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap,
                                          ctx->stream, &ctx->launches);
        if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_forward");
        for (uint32_t k : num) ok[k] = 1;
    }
    std::vector<uint32_t> post(nblocks);
    CK(cudaMemcpyAsync(post.data(), d_post, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t b = 0; b < nblocks; b++) {
        EncJob& j = jobs[b];
        j.data_off = (uint64_t)b * tstride;
        if (ok[b]) {
            j.post_len = post[b];
            j.skip_flags = 0x7F;
        } else if (!j.copy) {
            j.skip_flags = 0xFF;  // the only transform failed -> all skip bits stay set (Sequence.go:100-105)
        }
    }
    *d_data = tmp;
    return 0;
}

// Inverse transform stage (ByteTransformSequence.Inverse, transform/Sequence.go:131-186). The entropy stage left block
// b's pre_len bytes at d_pre + b*pstride; the decoded blocks are written back to back at d_dst. *total = bytes produced.
// lz_slot / compact only matter for LZ: block b is decoded at d_dst + b * lz_slot; compact = the slots must tile the
// output exactly (whole stream), otherwise the caller collects the slots itself (block batch).
int inverse_stage(kz_ctx* ctx, const TransformPlan& plan, std::vector<DecJob>& jobs, const uint8_t* d_pre, uint64_t pstride, uint8_t* d_dst, size_t cap,
                  std::vector<uint32_t>& out_len, uint64_t* total, uint64_t lz_slot = 0, bool compact = true) {
    const uint32_t nblocks = (uint32_t)jobs.size();
    out_len.assign(nblocks, 0);
    *total = 0;
    if (nblocks == 0) return 0;
    if (plan.generic())
        return inverse_generic(ctx, plan, jobs, const_cast<uint8_t*>(d_pre), pstride, (uint32_t)std::min<uint64_t>(pstride - 64, 0xFFFFFFF0u), d_dst, cap, out_len,
                               total, lz_slot, compact);
    if (plan.lz()) {
        // every block decodes into its own slot (valid streams: all blocks but the last are block_size long)
        uint32_t max_pre = 0;
        for (auto& j : jobs) max_pre = std::max(max_pre, j.pre_len);
        const uint64_t slot = lz_slot;
        std::vector<LzBlock> lb(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            const bool skip = jobs[b].copy || (jobs[b].skip_flags & 0x80);
            lb[b].src_off = (uint64_t)b * pstride;
            lb[b].dst_off = (uint64_t)b * slot;
            lb[b].scratch_off = 0;
            lb[b].len = skip ? 0 : jobs[b].pre_len;
            const uint64_t room = (uint64_t)b * slot < cap ? cap - (uint64_t)b * slot : 0;
            lb[b].cap = (uint32_t)std::min<uint64_t>(room, slot);
            lb[b].data_type = 0;
            lb[b].pad = 0;
        }
        std::vector<uint32_t> hl(nblocks);
        std::vector<int32_t> hs(nblocks);
        static const bool serial_inverse = [] {  // KZ_LZ_INV=serial: the round-1 kernel (one warp per block copies match by match)
            const char* e = getenv("KZ_LZ_INV");
            return e && e[0] == 's';
        }();
        if (!serial_inverse) {  // the stream's only stage is LZ / LZX (kanzi -l 1): the same data-parallel inverse as inside a sequence
            CtxLzWorkspace W(ctx);
            CtxLzHook H(ctx);
            cudaError_t e = lz_inverse_parallel(d_pre, d_dst, lb, W, ctx->stream, hl, hs, &H, &ctx->launches);
            if (e != cudaSuccess) return ctx->cuda_fail(e, "lz_inverse");
        } else {
            Packer pk;
            const size_t o_lb = pk.add(lb.data(), lb.size() * sizeof(LzBlock));
            const size_t o_len = pk.reserve((size_t)nblocks * 4);
            const size_t o_st = pk.reserve((size_t)nblocks * 4);
            int rc = upload(ctx, pk);
            if (rc) return rc;
            uint8_t* T = ctx->d_tables.as<uint8_t>();
            {
                LaunchScope ls(ctx, "lz_inverse");
                lz_inverse_kernel<<<nblocks, 32, 0, ctx->stream>>>(d_pre, (LzBlock*)(T + o_lb), (int)nblocks, d_dst, (uint32_t*)(T + o_len), (int32_t*)(T + o_st));
            }
            CK(cudaMemcpyAsync(hl.data(), T + o_len, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(hs.data(), T + o_st, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        uint64_t off = 0;
        for (uint32_t b = 0; b < nblocks; b++) {
            const bool skip = jobs[b].copy || (jobs[b].skip_flags & 0x80);
            if (skip) {
                if ((uint64_t)b * slot + jobs[b].pre_len > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
                CK(cudaMemcpyAsync(d_dst + (uint64_t)b * slot, d_pre + (uint64_t)b * pstride, jobs[b].pre_len, cudaMemcpyDeviceToDevice, ctx->stream));
                out_len[b] = jobs[b].pre_len;
            } else {
                if (hs[b]) {
                    jobs[b].status = hs[b];
                    return ctx->fail(KZ_ERR_PROCESS_BLOCK, "LZCodec inverse transform failed");
                }
                out_len[b] = hl[b];
            }
            if (b + 1 < nblocks && out_len[b] != slot && compact) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "unexpected decoded block length");
            off += out_len[b];
        }
        *total = compact ? off : 0;
        return 0;
    }
    // block headers of the BWT codec (<= 33 bytes) are needed on the host
    std::vector<uint8_t> heads((size_t)nblocks * 64, 0);
    CK(cudaMemcpy2DAsync(heads.data(), 64, d_pre, pstride, 33, nblocks, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // output offsets follow from the block headers (decoded length = stream length - header), so all blocks can be inverted together
    uint64_t off = 0;
    std::vector<uint64_t> so, dof;
    std::vector<uint32_t> ll, num;
    std::vector<uint8_t> hh;
    uint64_t batch_total = 0;
    auto flush = [&]() -> int {
        if (ll.empty()) return 0;
        CK(ctx->d_ws.ensure(bwt_inverse_batch_workspace(batch_total, (uint32_t)ll.size())));
        std::vector<uint32_t> produced(ll.size(), 0);
        LaunchScope ls(ctx, "bwt_inverse");
        cudaError_t e = bwt_inverse_batch(d_pre, d_dst, so.data(), dof.data(), ll.data(), hh.data(), (uint32_t)ll.size(), 0xFFFFFFFFu, produced.data(),
                                          ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream, &ctx->launches);
        if (e == cudaErrorInvalidValue) {
            for (uint32_t k : num) jobs[k].status = -KZ_ERR_PROCESS_BLOCK;
            return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header or data");
        }
        if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_inverse");
        so.clear(), dof.clear(), ll.clear(), num.clear(), hh.clear();
        batch_total = 0;
        return 0;
    };
    for (uint32_t b = 0; b < nblocks; b++) {
        DecJob& j = jobs[b];
        const uint8_t* src = d_pre + (uint64_t)b * pstride;
        const bool skip = j.copy || (j.skip_flags & 0x80) || !plan.bwt();
        if (skip) {
            if (off + j.pre_len > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
            CK(cudaMemcpyAsync(d_dst + off, src, j.pre_len, cudaMemcpyDeviceToDevice, ctx->stream));
            out_len[b] = j.pre_len;
        } else {
            const uint8_t* h = &heads[(size_t)b * 64];
            const uint32_t header = (1u << ((h[0] >> 2) & 7)) * ((uint32_t)(h[0] & 3) + 1) + 1;
            if (j.pre_len < 2 || j.pre_len < header) {
                j.status = -KZ_ERR_PROCESS_BLOCK;
                return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header");
            }
            const uint32_t produced = j.pre_len - header;
            if (off + produced > cap) {
                j.status = -KZ_ERR_PROCESS_BLOCK;
                return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header");
            }
            if (!ll.empty() && (batch_total + j.pre_len > bwt_inverse_batch_max_total() || ll.size() >= bwt_inverse_batch_max_blocks())) {
                int rc = flush();
                if (rc) return rc;
            }
            so.push_back((uint64_t)b * pstride), dof.push_back(off), ll.push_back(j.pre_len), num.push_back(b);
            hh.insert(hh.end(), h, h + 64);
            batch_total += j.pre_len;
            out_len[b] = produced;
        }
        off += out_len[b];
    }
    {
        int rc = flush();
        if (rc) return rc;
    }
    *total = off;
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int kz_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int kz_init(int device, kz_ctx** out) {
    if (!out) return -KZ_ERR_INVALID_PARAM;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return -KZ_ERR_CREATE_COMPRESSOR;
    if (cudaSetDevice(device) != cudaSuccess) return -KZ_ERR_CREATE_COMPRESSOR;
    kz_ctx* ctx = new kz_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return -KZ_ERR_CREATE_COMPRESSOR;
    }
    cudaFuncSetAttribute(ans0_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM);
    cudaFuncSetAttribute(ans0_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC_SMEM);
    cudaFuncSetAttribute(ans0_decode4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC4_SMEM);
    cudaFuncSetAttribute(ans0_decode5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC4_SMEM);
    cudaFuncSetAttribute(ans0_decode6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC6_SMEM);
    if (const char* v = getenv("KZ_ANS0_DEC")) ctx->ans0_dec_version = atoi(v);  // 3 = previous kernel generation (A/B measurements)
    cudaFuncSetAttribute(huf_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HUF_DEC_SMEM);
    *out = ctx;
    return 0;
}

void kz_destroy(kz_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drain_profile(ctx);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    DevBuf* bufs[] = {&ctx->d_in, &ctx->d_out, &ctx->d_tmp, &ctx->d_tmp2, &ctx->d_seg, &ctx->d_chunks, &ctx->d_tables, &ctx->d_enc_tab, &ctx->d_hdr, &ctx->d_pay, &ctx->d_small, &ctx->d_dl, &ctx->d_rcp, &ctx->d_ws, &ctx->d_lens, &ctx->d_hist};
    for (auto b : bufs) b->release();
    ctx->h_stage.release();
    ctx->h_dl.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* kz_last_error(kz_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void* kz_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void kz_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

void* kz_cuda_stream(kz_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

uint64_t kz_launch_count(kz_ctx* ctx, int reset) {
    uint64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

void kz_profile(kz_ctx* ctx, int enable) {
    drain_profile(ctx);
    ctx->profile = enable != 0;
}
uint32_t kz_kernel_time(kz_ctx* ctx, const char* name, double* total_ms) {
    drain_profile(ctx);
    auto it = ctx->prof.find(name);
    if (it == ctx->prof.end()) {
        if (total_ms) *total_ms = 0;
        return 0;
    }
    if (total_ms) *total_ms = it->second.ms;
    return it->second.launches;
}
void kz_profile_reset(kz_ctx* ctx) {
    drain_profile(ctx);
    ctx->prof.clear();
}
size_t kz_profile_names(kz_ctx* ctx, char* buf, size_t cap) {
    drain_profile(ctx);
    std::string all;
    for (auto& kv : ctx->prof) all += kv.first + "\n";
    if (buf && cap) {
        const size_t k = std::min(cap - 1, all.size());
        memcpy(buf, all.data(), k);
        buf[k] = 0;
    }
    return all.size() + 1;
}
uint32_t kz_stage_bytes(kz_ctx* ctx, const char* name, uint64_t* bytes_in, uint64_t* bytes_out) {
    auto it = ctx->prof.find(name);
    if (bytes_in) *bytes_in = it == ctx->prof.end() ? 0 : it->second.bytes_in;
    if (bytes_out) *bytes_out = it == ctx->prof.end() ? 0 : it->second.bytes_out;
    return it == ctx->prof.end() ? 0 : it->second.launches;
}

// worst cases: ANS0/Huffman ~440/550 header bytes per 16 KiB chunk; ANS1 256 context tables (<= 104 KiB) per 4 MiB chunk
size_t kz_max_block_output(size_t n) { return n + (n >> 3) + 1024 + 64 * (n / 16384 + 1) + 106496 * (n / (4u << 20) + 1); }
size_t kz_max_stream_output(size_t n, uint32_t block_size) {
    size_t nblocks = block_size ? (n + block_size - 1) / block_size : 1;
    return n + (n >> 3) + 64 * (n / 16384 + 1) + 32 * nblocks + 1024 + 106496 * (n / (4u << 20) + nblocks + 1);
}

void kz_set_stream_block_size(kz_ctx* ctx, uint64_t block_size) {
    if (ctx) ctx->user_stream_bs = block_size;
}

size_t kz_transform_max_encoded_len(uint64_t type, size_t n) {
    switch (type) {
        case KZ_T_NONE: return n;
        case KZ_T_BWT: return n + 33;  // transform/BWTBlockCodec.go:228
        case KZ_T_RANK:
        case KZ_T_MTFT: return n + 33;  // transform/SBRT.go:229-231
        case KZ_T_ZRLT: return n;       // transform/ZRLT.go:228-230
        case KZ_T_BWTS: return n;       // transform/BWTS.go:291-293
        case KZ_T_ROLZ: return n <= 512 ? n + 64 : n;  // transform/ROLZCodec.go:916-918
        case KZ_T_PACK: case KZ_T_DNA: return n + 1024;  // transform/AliasCodec.go:437-439
        case KZ_T_MM: return n + std::max<size_t>(n >> 4, 64);  // transform/FSDCodec.go:128-130
        case KZ_T_TEXT: return n;  // transform/TextCodec.go:1720-1724
        case KZ_T_UTF: return n + 8192;  // transform/UTFCodec.go:386-388
        case KZ_T_EXE: return n <= 256 ? n + 32 : n + n / 8;  // transform/EXECodec.go:701-708
        case KZ_T_LZ:
        case KZ_T_LZX: return n <= 1024 ? n + 16 : n + n / 64;  // transform/LZCodec.go:935-941
        default: return 0;
    }
}

// ---- whole stream, device resident ---------------------------------------------------------------------------
static int compress_device_impl(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size, const void* d_src,
                                size_t n, void* d_dst, size_t cap, size_t* out_n, bool fragment, uint64_t* out_bits);
int kz_compress_stream_device(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size,
                              const void* d_src, size_t n, void* d_dst, size_t cap, size_t* out_n) {
    return compress_device_impl(ctx, t48, etype, block_size, checksum_bits, input_size, d_src, n, d_dst, cap, out_n, false, nullptr);
}
// The block records (5 b | lw b | payload each) of n bytes at d_src, without stream header and without end marker: what one rank of a sharded
// compression contributes (blocks are independent, io/CompressedStream.go:896-898). *out_bits = exact bit length.
int kz_compress_fragment_device(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, const void* d_src, size_t n, void* d_dst,
                                size_t cap, uint64_t* out_bits) {
    size_t bytes = 0;
    if (!out_bits) return -KZ_ERR_INVALID_PARAM;
    return compress_device_impl(ctx, t48, etype, block_size, checksum_bits, 0, d_src, n, d_dst, cap, &bytes, true, out_bits);
}
static int compress_device_impl(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size, const void* d_src,
                                size_t n, void* d_dst, size_t cap, size_t* out_n, bool fragment, uint64_t* out_bits) {
    if (!ctx || !out_n) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    ctx->stream_bs = block_size;
    if (block_size < 1024 || block_size > (1u << 30) || (block_size & 15)) return ctx->fail(KZ_ERR_INVALID_PARAM, "Invalid block size");
    if (checksum_bits != 0 && checksum_bits != 32 && checksum_bits != 64) return ctx->fail(KZ_ERR_INVALID_PARAM, "checksum must be 0, 32 or 64 bits");
    if (((uintptr_t)d_src & 15) || ((uintptr_t)d_dst & 15)) return ctx->fail(KZ_ERR_INVALID_PARAM, "device buffers must be 16-byte aligned");
    const uint32_t nblocks = (uint32_t)((n + block_size - 1) / block_size);
    std::vector<EncJob> jobs;
    TransformPlan plan;
    int rc = plan_transforms(ctx, t48, etype, plan, KZ_ERR_CREATE_CODEC);
    if (rc) return rc;
    const uint8_t* d_data = nullptr;
    std::vector<int32_t> init_dt;
    if (plan.uses_magic() && nblocks && experimental_stages()) {  // the first four bytes of every block decide its initial data type
        std::vector<uint8_t> heads((size_t)nblocks * 4, 0);
        const size_t last = n - (size_t)(nblocks - 1) * block_size;
        if (nblocks > 1) CK(cudaMemcpy2DAsync(heads.data(), 4, d_src, block_size, 4, nblocks - 1, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&heads[(size_t)(nblocks - 1) * 4], (const uint8_t*)d_src + (size_t)(nblocks - 1) * block_size, std::min<size_t>(4, last),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        init_dt.resize(nblocks);
        for (uint32_t b = 0; b < nblocks; b++)
            init_dt[b] = magic_data_type(&heads[(size_t)b * 4], b + 1 < nblocks ? block_size : last);
    }
    rc = forward_stage(ctx, plan, (const uint8_t*)d_src, n, block_size, block_size, nullptr, nblocks, &d_data, jobs, init_dt.empty() ? nullptr : &init_dt);
    if (rc) return rc;
    if (checksum_bits && nblocks) {  // encodingTask.encode :753-763: hash of the original block
        std::vector<uint64_t> hoff(nblocks), hv;
        std::vector<uint32_t> hlen(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            hoff[b] = (uint64_t)b * block_size;
            hlen[b] = (uint32_t)std::min<uint64_t>(block_size, n - (uint64_t)b * block_size);
        }
        rc = hash_blocks(ctx, (const uint8_t*)d_src, hoff, hlen, checksum_bits, hv);
        if (rc) return rc;
        for (uint32_t b = 0; b < nblocks; b++) jobs[b].checksum = hv[b];
    }
    std::vector<uint8_t> hdr;
    uint32_t hdr_bits = build_stream_header(hdr, t48, etype, block_size, checksum_bits, input_size);
    if (fragment) hdr_bits = 0;
    EncResult res;
    ctx->fragment = fragment;
    rc = encode_batch(ctx, LAYOUT_STREAM, etype, checksum_bits, d_data, jobs, hdr, hdr_bits, (uint8_t*)d_dst, cap, {}, res);
    ctx->fragment = false;
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    *out_n = (size_t)((res.total_bits + 7) / 8);
    if (out_bits) *out_bits = res.total_bits;
    return 0;
}

// Stream header bytes (Writer.writeHeader, io/CompressedStream.go:429-519) for the rank that commits a sharded stream. out: >= 32 bytes.
int kz_stream_header(uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size, uint8_t* out, uint32_t* out_bits) {
    if (!out || !out_bits) return -KZ_ERR_INVALID_PARAM;
    std::vector<uint8_t> hdr;
    *out_bits = build_stream_header(hdr, t48, etype, block_size, checksum_bits, input_size);
    memset(out, 0, 32);
    memcpy(out, hdr.data(), std::min<size_t>(hdr.size(), 32));
    return 0;
}

// Bit-granular concatenation of nseg device bit strings (MSB first) into d_dst: the ordered commit of Writer.processBlock (:951-976) for
// fragments gathered from other ranks. Segments are cut into 64 KiB pieces, one CTA each. *total_bits = length of the result.
int kz_concat_bits_device(kz_ctx* ctx, uint32_t nseg, const void* const* d_seg, const uint64_t* seg_bits, void* d_dst, size_t cap, uint64_t* total_bits) {
    if (!ctx || !total_bits || (nseg && (!d_seg || !seg_bits)) || !d_dst) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    std::vector<uint64_t> src, bits;
    uint64_t total = 0;
    for (uint32_t i = 0; i < nseg; i++) {
        if (((uintptr_t)d_seg[i] & 3) != 0) return ctx->fail(KZ_ERR_INVALID_PARAM, "segments must be 4-byte aligned");
        const uint64_t nbytes = (seg_bits[i] + 7) / 8;
        for (uint64_t o = 0; o < nbytes; o += RAW_PIECE) {
            src.push_back((uint64_t)(uintptr_t)((const uint8_t*)d_seg[i] + o));
            bits.push_back(std::min<uint64_t>(8 * RAW_PIECE, seg_bits[i] - 8 * o));
        }
        total += seg_bits[i];
    }
    *total_bits = total;
    if ((total + 7) / 8 + 8 > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
    const int nsegs = (int)src.size();
    if (nsegs == 0) return 0;
    Packer pk;
    const size_t o_src = pk.add(src.data(), src.size() * 8);
    const size_t o_bits = pk.add(bits.data(), bits.size() * 8);
    const size_t o_dst = pk.reserve((size_t)(nsegs + 1) * 8);
    const size_t o_total = pk.reserve(8);
    int rc = upload(ctx, pk);
    if (rc) return rc;
    uint8_t* T = ctx->d_tables.as<uint8_t>();
    {
        LaunchScope ls(ctx, "scan");
        scan_bits_kernel<<<1, 1024, 0, ctx->stream>>>((const uint64_t*)(T + o_bits), (uint64_t*)(T + o_dst), nsegs + 1, 0, (uint64_t*)(T + o_total));
    }
    {
        LaunchScope ls(ctx, "concat_zero");
        concat_zero_kernel<<<(nsegs + 255) / 256, 256, 0, ctx->stream>>>((const uint64_t*)(T + o_bits), (const uint64_t*)(T + o_dst), nsegs, (uint32_t*)d_dst);
    }
    {
        LaunchScope ls(ctx, "concat");
        concat_copy_kernel<<<nsegs, 128, 0, ctx->stream>>>((const uint64_t*)(T + o_src), (const uint64_t*)(T + o_bits), (const uint64_t*)(T + o_dst), nsegs, (uint32_t*)d_dst);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int decompress_device_impl(kz_ctx* ctx, const StreamHeader* given, uint64_t start_bit, uint32_t stop_after, const void* d_src, size_t n, void* d_dst,
                                  size_t cap, size_t* out_n);
int kz_decompress_stream_device(kz_ctx* ctx, const void* d_src, size_t n, void* d_dst, size_t cap, size_t* out_n) {
    return decompress_device_impl(ctx, nullptr, 0, 0, d_src, n, d_dst, cap, out_n);
}
// Decodes `nblocks` block records that start at bit `start_bit` of the fragment at d_src (frag_bytes bytes, >= 64 readable bytes behind them):
// what one rank of a sharded decompression does with its part of the stream. The stream parameters come from the header rank 0 parsed.
int kz_decompress_fragment_device(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, const void* d_src, size_t frag_bytes,
                                  uint64_t start_bit, uint32_t nblocks, void* d_dst, size_t cap, size_t* out_n) {
    if (!ctx || !out_n) return -KZ_ERR_INVALID_PARAM;
    *out_n = 0;
    if (nblocks == 0) return 0;
    StreamHeader sh;
    sh.t48 = t48;
    sh.etype = etype;
    sh.block_size = block_size;
    sh.checksum_bits = checksum_bits;
    sh.input_size = 0;
    sh.bits = 0;
    return decompress_device_impl(ctx, &sh, start_bit, nblocks, d_src, frag_bytes, d_dst, cap, out_n);
}
// Parses the header of the stream at d_src and walks its block records: rec_bit[b] = first bit of record b (its 5-bit prefix), rec_bit[nblocks] =
// first bit behind the last record (the end marker). rec_bit must hold max_blocks + 1 entries.
int kz_stream_index_device(kz_ctx* ctx, const void* d_src, size_t n, uint32_t max_blocks, uint64_t* t48, uint32_t* etype, uint32_t* block_size,
                           uint32_t* checksum_bits, int64_t* input_size, uint64_t* rec_bit, uint32_t* nblocks) {
    if (!ctx || !d_src || !rec_bit || !nblocks) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    uint8_t head[64];
    memset(head, 0, sizeof(head));
    CK(cudaMemcpyAsync(head, d_src, std::min<size_t>(n, 64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    StreamHeader sh;
    int rc = parse_stream_header(ctx, head, n, sh);
    if (rc) return rc;
    if (t48) *t48 = sh.t48;
    if (etype) *etype = sh.etype;
    if (block_size) *block_size = sh.block_size;
    if (checksum_bits) *checksum_bits = sh.checksum_bits;
    if (input_size) *input_size = sh.input_size;
    CK(ctx->d_dl.ensure((size_t)max_blocks * 16 + 64));
    uint64_t* d_boff = ctx->d_dl.as<uint64_t>();
    uint64_t* d_bbits = d_boff + max_blocks;
    uint32_t* d_cnt = (uint32_t*)(d_bbits + max_blocks);
    {
        LaunchScope ls(ctx, "stream_walk");
        stream_walk_kernel<<<1, 32, 0, ctx->stream>>>((const uint32_t*)d_src, sh.bits, 8ull * n, max_blocks, d_boff, d_bbits, d_cnt);
    }
    uint32_t cs[2] = {0, 0};
    CK(cudaMemcpyAsync(cs, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((int32_t)cs[1] != 0) return ctx->fail(-(int32_t)cs[1], "Invalid bitstream: corrupted block length");
    const uint32_t nb = cs[0];
    std::vector<uint64_t> boff(nb), bbits(nb);
    if (nb) {
        CK(cudaMemcpyAsync(boff.data(), d_boff, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(bbits.data(), d_bbits, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    rec_bit[0] = sh.bits;
    for (uint32_t b = 0; b < nb; b++) rec_bit[b + 1] = boff[b] + bbits[b];
    *nblocks = nb;
    return 0;
}
static int decompress_device_impl(kz_ctx* ctx, const StreamHeader* given, uint64_t start_bit, uint32_t stop_after, const void* d_src, size_t n, void* d_dst,
                                  size_t cap, size_t* out_n) {
    if (!ctx || !out_n) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    if (((uintptr_t)d_src & 15) || ((uintptr_t)d_dst & 15)) return ctx->fail(KZ_ERR_INVALID_PARAM, "device buffers must be 16-byte aligned");
    StreamHeader sh;
    int rc = 0;
    if (given) {
        sh = *given;
        sh.bits = start_bit;
    } else {
        uint8_t head[64];
        memset(head, 0, sizeof(head));
        CK(cudaMemcpyAsync(head, d_src, std::min<size_t>(n, 64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        rc = parse_stream_header(ctx, head, n, sh);
        if (rc) return rc;
    }
    TransformPlan plan;
    rc = plan_transforms(ctx, sh.t48, sh.etype, plan, KZ_ERR_INVALID_CODEC);
    if (rc) return rc;
    const uint32_t* d_words = (const uint32_t*)d_src;
    const uint64_t words_len = (n + 3) / 4 + 16;  // callers guarantee >= 64 readable bytes past n
    // block prefixes
    uint32_t max_blocks = (uint32_t)std::min<uint64_t>((uint64_t)cap / 1024 + 16, 1u << 22);
    CK(ctx->d_dl.ensure((size_t)max_blocks * 16 + 64));
    uint64_t* d_boff = ctx->d_dl.as<uint64_t>();
    uint64_t* d_bbits = d_boff + max_blocks;
    uint32_t* d_cnt = (uint32_t*)(d_bbits + max_blocks);
    {
        LaunchScope ls(ctx, "stream_walk");
        stream_walk_kernel<<<1, 32, 0, ctx->stream>>>(d_words, sh.bits, 8ull * n, max_blocks, d_boff, d_bbits, d_cnt, stop_after);
    }
    // count + the first SPEC block descriptors in one round trip (streams with more blocks pay a second copy)
    const uint32_t SPEC = std::min<uint32_t>(max_blocks, 1024);
    CK(ctx->h_dl.ensure((size_t)SPEC * 16 + 64));
    uint32_t* cs = (uint32_t*)ctx->h_dl.p;
    uint64_t* h_boff = (uint64_t*)((uint8_t*)ctx->h_dl.p + 64);
    uint64_t* h_bbits = h_boff + SPEC;
    CK(cudaMemcpyAsync(cs, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_boff, d_boff, (size_t)SPEC * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_bbits, d_bbits, (size_t)SPEC * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((int32_t)cs[1] != 0) return ctx->fail(-(int32_t)cs[1], "Invalid bitstream: corrupted block length");
    const uint32_t nblocks = cs[0];
    std::vector<uint64_t> boff(nblocks), bbits(nblocks);
    if (nblocks && nblocks <= SPEC) {
        memcpy(boff.data(), h_boff, (size_t)nblocks * 8);
        memcpy(bbits.data(), h_bbits, (size_t)nblocks * 8);
    } else if (nblocks) {
        CK(cudaMemcpyAsync(boff.data(), d_boff, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(bbits.data(), d_bbits, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<DecJob> jobs(nblocks);
    for (uint32_t b = 0; b < nblocks; b++) {
        jobs[b] = DecJob();
        jobs[b].bit_off = boff[b];
        jobs[b].bits = bbits[b];
    }
    ctx->stream_bs = sh.block_size;
    rc = decode_parse_headers(ctx, d_words, jobs, sh.checksum_bits, sh.block_size);
    if (rc) return rc;
    uint64_t out_off = 0;
    if (plan.none()) {
        for (uint32_t b = 0; b < nblocks; b++) {
            if (jobs[b].status) return ctx->fail(-jobs[b].status, "Invalid compressed block header");
            jobs[b].out_off = out_off;  // NONE transform: decoded length == pre_len; every block but the last is block_size (multiple of 16)
            out_off += jobs[b].pre_len;
            if (out_off > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
            if (b + 1 < nblocks && (jobs[b].pre_len & 15)) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "unexpected block length");
        }
        rc = decode_batch(ctx, sh.etype, d_words, words_len, jobs, (uint8_t*)d_dst);
        if (rc) return rc;
        for (uint32_t b = 0; b < nblocks; b++)
            if (jobs[b].status) return ctx->fail(-jobs[b].status, "Invalid bitstream: block decoding failed");
        if (sh.checksum_bits) {
            std::vector<uint64_t> hoff(nblocks);
            std::vector<uint32_t> hlen(nblocks);
            for (uint32_t b = 0; b < nblocks; b++) {
                hoff[b] = jobs[b].out_off;
                hlen[b] = jobs[b].pre_len;
            }
            rc = verify_checksums(ctx, (const uint8_t*)d_dst, hoff, hlen, sh.checksum_bits, jobs);
            if (rc) return rc;
        }
        *out_n = (size_t)out_off;
        return 0;
    }
    // transforms: entropy-decode every block into scratch, then invert the transforms into d_dst
    uint32_t max_pre = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        if (jobs[b].status) return ctx->fail(-jobs[b].status, "Invalid compressed block header");
        max_pre = std::max(max_pre, jobs[b].pre_len);
    }
    uint64_t blk_room = max_pre;
    if (plan.generic())  // Reader.processBlock :1649-1653: the task buffers hold blockSize + max(512, blockSize >> 4) bytes
        blk_room = std::max<uint64_t>(blk_room, (uint64_t)sh.block_size + std::max<uint64_t>(512, sh.block_size >> 4));
    const uint64_t pstride = (blk_room + 64 + 15) & ~15ull;
    CK(ctx->d_tmp.ensure(pstride * nblocks + 64));
    for (uint32_t b = 0; b < nblocks; b++) jobs[b].out_off = (uint64_t)b * pstride;
    rc = decode_batch(ctx, sh.etype, d_words, words_len, jobs, ctx->d_tmp.as<uint8_t>());
    if (rc) return rc;
    for (uint32_t b = 0; b < nblocks; b++)
        if (jobs[b].status) return ctx->fail(-jobs[b].status, "Invalid bitstream: block decoding failed");
    std::vector<uint32_t> olen;
    rc = inverse_stage(ctx, plan, jobs, ctx->d_tmp.as<uint8_t>(), pstride, (uint8_t*)d_dst, cap, olen, &out_off, sh.block_size, true);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    if (sh.checksum_bits) {  // decoded blocks lie back to back (every block but the last is block_size long)
        std::vector<uint64_t> hoff(nblocks);
        uint64_t o = 0;
        for (uint32_t b = 0; b < nblocks; b++) {
            hoff[b] = o;
            o += olen[b];
        }
        rc = verify_checksums(ctx, (const uint8_t*)d_dst, hoff, olen, sh.checksum_bits, jobs);
        if (rc) return rc;
    }
    *out_n = (size_t)out_off;
    return 0;
}

// ---- whole stream, host buffers ------------------------------------------------------------------------------
int kz_compress_stream(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t block_size, uint32_t checksum_bits, int64_t input_size,
                       const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if (!ctx || !out_n || (!src && n) || !dst) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    const size_t dcap = kz_max_stream_output(n, block_size);
    CK(ctx->d_in.ensure(n + 64));
    CK(ctx->d_out.ensure(dcap + 64));
    if (n) CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    size_t produced = 0;
    int rc = kz_compress_stream_device(ctx, t48, etype, block_size, checksum_bits, input_size, ctx->d_in.p, n, ctx->d_out.p, dcap, &produced);
    if (rc) return rc;
    if (produced > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
    CK(cudaMemcpyAsync(dst, ctx->d_out.p, produced, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out_n = produced;
    return 0;
}

int kz_decompress_stream(kz_ctx* ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if (!ctx || !out_n || !src || (!dst && cap)) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_in.ensure(n + 128));
    CK(ctx->d_out.ensure(cap + 64));
    CK(cudaMemsetAsync((uint8_t*)ctx->d_in.p + (n & ~size_t(3)), 0, 128, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    size_t produced = 0;
    int rc = kz_decompress_stream_device(ctx, ctx->d_in.p, n, ctx->d_out.p, cap, &produced);
    if (rc) return rc;
    if (produced) CK(cudaMemcpyAsync(dst, ctx->d_out.p, produced, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out_n = produced;
    return 0;
}

// ---- block batch, host buffers (Writer.processBlock / Reader.processBlock replacement) -----------------------
int kz_encode_blocks(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t checksum_bits, const uint8_t* slab, uint64_t block_stride,
                     const uint32_t* block_len, uint32_t nblocks, uint8_t* out, uint64_t out_stride, uint64_t* out_bits, int32_t* status) {
    if (!ctx || !slab || !block_len || !out || !out_bits) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    ctx->stream_bs = ctx->user_stream_bs;
    if (checksum_bits != 0 && checksum_bits != 32 && checksum_bits != 64) return ctx->fail(KZ_ERR_INVALID_PARAM, "checksum must be 0, 32 or 64 bits");
    if (nblocks == 0) return 0;
    // device layout: block b at b * dstride (16-byte aligned)
    uint32_t max_len = 0;
    for (uint32_t b = 0; b < nblocks; b++) max_len = std::max(max_len, block_len[b]);
    const uint64_t dstride = ((uint64_t)max_len + 15) & ~15ull;
    const uint64_t ostride = (kz_max_block_output(max_len) + 15) & ~size_t(15);
    if (out_stride < ostride) return ctx->fail(KZ_ERR_INVALID_PARAM, "out_stride must be >= kz_max_block_output(max block length)");
    CK(ctx->d_in.ensure(dstride * nblocks + 64));
    CK(ctx->d_out.ensure(ostride * nblocks + 64));
    for (uint32_t b = 0; b < nblocks; b++)
        if (block_len[b])
            CK(cudaMemcpyAsync((uint8_t*)ctx->d_in.p + b * dstride, slab + b * block_stride, block_len[b], cudaMemcpyHostToDevice, ctx->stream));
    std::vector<EncJob> jobs;
    TransformPlan plan;
    int rc = plan_transforms(ctx, t48, etype, plan, KZ_ERR_CREATE_CODEC);
    if (rc) return rc;
    const uint8_t* d_data = nullptr;
    std::vector<int32_t> init_dt(nblocks, 0);
    if (experimental_stages())
        for (uint32_t b = 0; b < nblocks; b++) init_dt[b] = magic_data_type(slab + b * block_stride, block_len[b]);
    rc = forward_stage(ctx, plan, ctx->d_in.as<uint8_t>(), 0, 0, dstride, block_len, nblocks, &d_data, jobs, &init_dt);
    if (rc) return rc;
    if (checksum_bits) {  // encodingTask.encode :753-763: hash of the original block
        std::vector<uint64_t> hoff(nblocks), hv;
        std::vector<uint32_t> hlen(block_len, block_len + nblocks);
        for (uint32_t b = 0; b < nblocks; b++) hoff[b] = (uint64_t)b * dstride;
        rc = hash_blocks(ctx, ctx->d_in.as<uint8_t>(), hoff, hlen, checksum_bits, hv);
        if (rc) return rc;
        for (uint32_t b = 0; b < nblocks; b++) jobs[b].checksum = hv[b];
    }
    std::vector<uint64_t> base(nblocks);
    for (uint32_t b = 0; b < nblocks; b++) base[b] = 8 * (uint64_t)b * ostride;
    EncResult res;
    rc = encode_batch(ctx, LAYOUT_BLOCKS, etype, checksum_bits, d_data, jobs, {}, 0, ctx->d_out.as<uint8_t>(), ostride * nblocks, base, res);
    if (rc) return rc;
    for (uint32_t b = 0; b < nblocks; b++) {
        out_bits[b] = res.blk_bits[b];
        if (status) status[b] = 0;
        size_t nbytes = (size_t)((res.blk_bits[b] + 7) / 8);
        CK(cudaMemcpyAsync(out + b * out_stride, (uint8_t*)ctx->d_out.p + b * ostride, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int kz_decode_blocks(kz_ctx* ctx, uint64_t t48, uint32_t etype, uint32_t checksum_bits, const uint8_t* in, const uint64_t* in_off,
                     const uint64_t* in_bits, uint32_t nblocks, uint32_t block_size, uint8_t* out, uint64_t out_stride, uint32_t* out_len,
                     int32_t* status) {
    if (!ctx || !in || !in_off || !in_bits || !out || !out_len) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    ctx->stream_bs = block_size;
    if (checksum_bits != 0 && checksum_bits != 32 && checksum_bits != 64) return ctx->fail(KZ_ERR_INVALID_PARAM, "checksum must be 0, 32 or 64 bits");
    TransformPlan plan;
    {
        int prc = plan_transforms(ctx, t48, etype, plan, KZ_ERR_INVALID_CODEC);
        if (prc) return prc;
    }
    if (nblocks == 0) return 0;
    // pack the block bit strings at 16-byte aligned device offsets
    std::vector<DecJob> jobs(nblocks);
    uint64_t doff = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        jobs[b] = DecJob();
        jobs[b].bit_off = 8 * doff;
        jobs[b].bits = in_bits[b];
        doff += (((in_bits[b] + 7) / 8) + 8 + 15) & ~15ull;
    }
    const uint64_t dsz = doff;
    CK(ctx->d_in.ensure(dsz + 128));
    CK(cudaMemsetAsync(ctx->d_in.p, 0, dsz + 128, ctx->stream));
    for (uint32_t b = 0; b < nblocks; b++)
        CK(cudaMemcpyAsync((uint8_t*)ctx->d_in.p + jobs[b].bit_off / 8, in + in_off[b], (size_t)((in_bits[b] + 7) / 8), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t* d_words = ctx->d_in.as<uint32_t>();
    int rc = decode_parse_headers(ctx, d_words, jobs, checksum_bits, block_size);
    if (rc) return rc;
    const uint64_t ostride = (((uint64_t)block_size + block_size / 2 + 2048) + 15) & ~15ull;
    CK(ctx->d_out.ensure(ostride * nblocks + 64));
    for (uint32_t b = 0; b < nblocks; b++) {
        jobs[b].out_off = b * ostride;
        if (jobs[b].status == 0 && jobs[b].pre_len > out_stride) jobs[b].status = -KZ_ERR_WRITE_FILE;
    }
    rc = decode_batch(ctx, etype, d_words, dsz / 4 + 16, jobs, ctx->d_out.as<uint8_t>());
    if (rc) return rc;
    int worst = 0;
    for (uint32_t b = 0; b < nblocks; b++)
        if (jobs[b].status && !worst) worst = jobs[b].status;
    if (checksum_bits && plan.none() && !worst) {
        std::vector<uint64_t> hoff(nblocks);
        std::vector<uint32_t> hlen(nblocks);
        for (uint32_t b = 0; b < nblocks; b++) {
            hoff[b] = b * ostride;
            hlen[b] = jobs[b].pre_len;
        }
        int vrc = verify_checksums(ctx, ctx->d_out.as<uint8_t>(), hoff, hlen, checksum_bits, jobs);
        if (vrc) return vrc;
    }
    if (plan.none() || worst) {
        for (uint32_t b = 0; b < nblocks; b++) {
            if (status) status[b] = jobs[b].status;
            if (jobs[b].status) {
                out_len[b] = 0;
                continue;
            }
            out_len[b] = jobs[b].pre_len;
            CK(cudaMemcpyAsync(out + b * out_stride, (uint8_t*)ctx->d_out.p + b * ostride, jobs[b].pre_len, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        if (worst) return ctx->fail(-worst, "block decoding failed");
        return 0;
    }
    // inverse transforms into a packed scratch area, then out
    CK(ctx->d_tmp.ensure(ostride * nblocks + 64));
    std::vector<uint32_t> olen;
    uint64_t total = 0;
    rc = inverse_stage(ctx, plan, jobs, ctx->d_out.as<uint8_t>(), ostride, ctx->d_tmp.as<uint8_t>(), ostride * nblocks, olen, &total, ostride, false);
    if (rc) return rc;
    if (checksum_bits) {
        std::vector<uint64_t> hoff(nblocks);
        uint64_t o = 0;
        for (uint32_t b = 0; b < nblocks; b++) {
            hoff[b] = (plan.lz() || plan.generic()) ? (uint64_t)b * ostride : o;
            o += olen[b];
        }
        int vrc = verify_checksums(ctx, ctx->d_tmp.as<uint8_t>(), hoff, olen, checksum_bits, jobs);
        if (vrc) return vrc;
    }
    uint64_t off = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        if (status) status[b] = 0;
        out_len[b] = olen[b];
        if (olen[b] > out_stride) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
        const uint64_t from = (plan.lz() || plan.generic()) ? (uint64_t)b * ostride : off;
        CK(cudaMemcpyAsync(out + b * out_stride, ctx->d_tmp.as<uint8_t>() + from, olen[b], cudaMemcpyDeviceToHost, ctx->stream));
        off += olen[b];
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- single block entropy codec (EntropyEncoder.Write / EntropyDecoder.Read) ----------------------------------
int kz_entropy_encode(kz_ctx* ctx, uint32_t type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, uint64_t* out_bits) {
    if (!ctx || (!src && n) || !dst || !out_bits) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    if (n > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
    *out_bits = 0;
    if (n == 0) return 0;
    const size_t dcap = kz_max_block_output(n);
    CK(ctx->d_in.ensure(n + 64));
    CK(ctx->d_out.ensure(dcap + 64));
    CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<EncJob> jobs(1);
    jobs[0].data_off = 0;
    jobs[0].post_len = (uint32_t)n;
    jobs[0].ntransforms = 1;
    jobs[0].skip_flags = 0;
    jobs[0].copy = false;
    jobs[0].checksum = 0;
    EncResult res;
    int rc = encode_batch(ctx, LAYOUT_RAW, type, 0, ctx->d_in.as<uint8_t>(), jobs, {}, 0, ctx->d_out.as<uint8_t>(), dcap, {0}, res);
    if (rc) return rc;
    const size_t nbytes = (size_t)((res.blk_bits[0] + 7) / 8);
    if (nbytes > cap) return ctx->fail(KZ_ERR_WRITE_FILE, "output buffer too small");
    CK(cudaMemcpyAsync(dst, ctx->d_out.p, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out_bits = res.blk_bits[0];
    return 0;
}

int kz_entropy_decode(kz_ctx* ctx, uint32_t type, const uint8_t* src, size_t src_bytes, uint8_t* dst, size_t n, uint64_t* consumed_bits) {
    if (!ctx || !src || (!dst && n)) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    if (consumed_bits) *consumed_bits = 0;
    if (n == 0) return 0;
    if (n > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
    CK(ctx->d_in.ensure(src_bytes + 128));
    CK(ctx->d_out.ensure(n + 64));
    CK(cudaMemsetAsync((uint8_t*)ctx->d_in.p + (src_bytes & ~size_t(3)), 0, 128, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in.p, src, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<DecJob> jobs(1);
    jobs[0] = DecJob();
    jobs[0].bit_off = 0;
    jobs[0].bits = 8ull * src_bytes;
    jobs[0].out_off = 0;
    jobs[0].pre_len = (uint32_t)n;
    jobs[0].data_bit = 0;
    jobs[0].copy = false;
    jobs[0].status = 0;
    int rc = decode_batch(ctx, type, ctx->d_in.as<uint32_t>(), src_bytes / 4 + 16, jobs, ctx->d_out.as<uint8_t>());
    if (rc) return rc;
    if (jobs[0].status) return ctx->fail(-jobs[0].status, "Invalid bitstream: entropy decoding failed");
    CK(cudaMemcpyAsync(dst, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (consumed_bits) *consumed_bits = jobs[0].end_pos;
    return 0;
}

int kz_transform_forward(kz_ctx* ctx, uint64_t type, int* data_type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if ((type == KZ_T_PACK || type == KZ_T_DNA || type == KZ_T_MM || type == KZ_T_TEXT || type == KZ_T_UTF || type == KZ_T_EXE) && ctx && !experimental_stages())
        return ctx->fail(KZ_ERR_CREATE_CODEC, "transform switched off by KZ_EXPERIMENTAL=0: PACK / DNA / MM / TEXT");
    if (!ctx || (!src && n) || !dst || !out_n) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    ctx->stream_bs = ctx->user_stream_bs;
    *out_n = 0;
    if (type == KZ_T_LZ || type == KZ_T_LZX) {  // LZXCodec.Forward (reads ctx["dataType"], never writes it)
        if (n == 0) return 0;
        if (n > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
        if (cap < kz_transform_max_encoded_len(type, n)) return ctx->fail(KZ_ERR_INVALID_PARAM, "Output buffer is too small");
        TransformPlan plan;
        plan.nt = 1;
        plan.ids[0] = type;
        CK(ctx->d_in.ensure(n + 64));
        CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
        const uint32_t len = (uint32_t)n;
        const uint8_t* d_data = nullptr;
        std::vector<EncJob> jobs;
        const std::vector<int32_t> dt1(1, data_type ? *data_type : 0);
        int rc = forward_stage(ctx, plan, ctx->d_in.as<uint8_t>(), n, 0, 0, &len, 1, &d_data, jobs, &dt1, false);
        if (rc) return rc;
        if (jobs[0].skip_flags & 0x80) return 1;
        CK(cudaMemcpyAsync(dst, d_data, jobs[0].post_len, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *out_n = jobs[0].post_len;
        return 0;
    }
    if (type == KZ_T_RANK || type == KZ_T_MTFT || type == KZ_T_ZRLT || type == KZ_T_BWTS || type == KZ_T_ROLZ || type == KZ_T_PACK || type == KZ_T_DNA || type == KZ_T_MM || type == KZ_T_TEXT || type == KZ_T_UTF || type == KZ_T_EXE) {  // SBRT / ZRLT / BWTS / ROLZ / Alias / FSD / Text / UTF / EXE Forward
        if (n == 0) return 0;
        if (n > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
        if (cap < kz_transform_max_encoded_len(type, n)) return 1;  // Forward returns an error: the sequence skips the transform
        const uint64_t stride = ((uint64_t)n + n / 8 + 8192 + 1024 + 64 + 64 + 15) & ~15ull;
        CK(ctx->d_in.ensure(stride));
        CK(ctx->d_out.ensure(stride));
        CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<uint32_t> len(1, (uint32_t)n);
        std::vector<uint8_t> active(1, 1), applied;
        std::vector<int32_t> dt1(1, data_type ? *data_type : 0);
        int rc = apply_forward(ctx, type, ctx->d_in.as<uint8_t>(), stride, ctx->d_out.as<uint8_t>(), stride, len, active, applied, &dt1);
        if (rc) return rc;
        if (data_type) *data_type = dt1[0];
        if (!applied[0]) return 1;
        CK(cudaMemcpyAsync(dst, ctx->d_out.p, len[0], cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *out_n = len[0];
        return 0;
    }
    if (type != KZ_T_BWT) return ctx->fail(KZ_ERR_CREATE_CODEC, "transform not available on the GPU path (BWT, LZ, LZX, RANK, MTFT, ZRLT are)");
    if (n == 0) return 0;
    if (n > (1u << 30)) return ctx->fail(KZ_ERR_BLOCK_SIZE, "The max BWT block size is 1 GiB");
    if (cap < n + 33) return ctx->fail(KZ_ERR_INVALID_PARAM, "Output buffer is too small");
    if (n < 2) return 1;  // BWTBlockCodec.Forward: invalid index size -> the sequence skips the transform
    CK(ctx->d_in.ensure(n + 64));
    CK(ctx->d_out.ensure(n + 128));
    CK(ctx->d_ws.ensure(bwt_forward_workspace((uint32_t)n)));
    CK(ctx->d_lens.ensure(64));
    CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e;
    {
        LaunchScope ls(ctx, "bwt_forward");
        e = bwt_forward_device(ctx->d_in.as<uint8_t>(), (uint32_t)n, ctx->d_out.as<uint8_t>(), ctx->d_lens.as<uint32_t>(), ctx->d_ws.as<uint8_t>(),
                               ctx->d_ws.cap, ctx->stream, &ctx->launches);
    }
    if (e == cudaErrorInvalidValue) return 1;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_forward");
    uint32_t produced = 0;
    CK(cudaMemcpyAsync(&produced, ctx->d_lens.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(dst, ctx->d_out.p, produced, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out_n = produced;
    return 0;
}

int kz_transform_inverse(kz_ctx* ctx, uint64_t type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n) {
    if ((type == KZ_T_PACK || type == KZ_T_DNA || type == KZ_T_MM || type == KZ_T_TEXT || type == KZ_T_UTF || type == KZ_T_EXE) && ctx && !experimental_stages())
        return ctx->fail(KZ_ERR_CREATE_CODEC, "transform switched off by KZ_EXPERIMENTAL=0: PACK / DNA / MM / TEXT");
    if (!ctx || (!src && n) || !dst || !out_n) return -KZ_ERR_INVALID_PARAM;
    CK(cudaSetDevice(ctx->device));
    ctx->stream_bs = ctx->user_stream_bs;
    *out_n = 0;
    if (type == KZ_T_LZ || type == KZ_T_LZX) {  // LZXCodec.inverseV6; cap = len(dst) of the reference call
        if (n == 0) return 0;
        if (n > (1u << 30) || cap > (1u << 30) + 64) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
        TransformPlan plan;
        plan.nt = 1;
        plan.ids[0] = type;
        CK(ctx->d_in.ensure(n + 64));
        CK(ctx->d_out.ensure(cap + 64));
        CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<DecJob> jobs(1);
        jobs[0] = DecJob();
        jobs[0].pre_len = (uint32_t)n;
        std::vector<uint32_t> olen;
        uint64_t total = 0;
        int rc = inverse_stage(ctx, plan, jobs, ctx->d_in.as<uint8_t>(), n, ctx->d_out.as<uint8_t>(), cap, olen, &total, cap, false);
        if (rc) return rc;
        if (olen[0]) CK(cudaMemcpyAsync(dst, ctx->d_out.p, olen[0], cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *out_n = olen[0];
        return 0;
    }
    if (type == KZ_T_RANK || type == KZ_T_MTFT || type == KZ_T_ZRLT || type == KZ_T_BWTS || type == KZ_T_ROLZ || type == KZ_T_PACK || type == KZ_T_DNA || type == KZ_T_MM || type == KZ_T_TEXT || type == KZ_T_UTF || type == KZ_T_EXE) {  // SBRT / ZRLT / BWTS / ROLZ / Alias / FSD / Text / UTF / EXE Inverse; cap = len(dst)
        if (n == 0 || cap == 0) return 0;
        if (n > (1u << 30) || cap > (1u << 30) + 4096) return ctx->fail(KZ_ERR_BLOCK_SIZE, "block too large");
        const uint64_t stride = ((uint64_t)std::max(n, cap) + 64 + 15) & ~15ull;
        CK(ctx->d_in.ensure(stride));
        CK(ctx->d_out.ensure(stride));
        CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<uint32_t> len(1, (uint32_t)n);
        std::vector<uint8_t> active(1, 1);
        int rc = apply_inverse(ctx, type, ctx->d_in.as<uint8_t>(), ctx->d_out.as<uint8_t>(), stride, (uint32_t)cap, len, active);
        if (rc) return rc;
        if (len[0]) CK(cudaMemcpyAsync(dst, ctx->d_out.p, len[0], cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *out_n = len[0];
        return 0;
    }
    if (type != KZ_T_BWT) return ctx->fail(KZ_ERR_CREATE_CODEC, "transform not available on the GPU path (BWT, LZ, LZX, RANK, MTFT, ZRLT are)");
    if (n == 0) return 0;
    if (n == 1) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid size");
    if (n > (1u << 30) + 33) return ctx->fail(KZ_ERR_BLOCK_SIZE, "The max BWT block size is 1 GiB");
    CK(ctx->d_in.ensure(n + 64));
    CK(ctx->d_out.ensure(n + 64));
    CK(ctx->d_ws.ensure(bwt_inverse_workspace((uint32_t)n)));
    CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t head[64];
    memset(head, 0, sizeof(head));
    memcpy(head, src, std::min<size_t>(n, 33));
    uint32_t produced = 0;
    cudaError_t e;
    {
        LaunchScope ls(ctx, "bwt_inverse");
        e = bwt_inverse_device(ctx->d_in.as<uint8_t>(), head, (uint32_t)n, ctx->d_out.as<uint8_t>(), (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFu), &produced,
                               ctx->d_ws.as<uint8_t>(), ctx->d_ws.cap, ctx->stream, &ctx->launches);
    }
    if (e == cudaErrorInvalidValue) return ctx->fail(KZ_ERR_PROCESS_BLOCK, "BWT inverse transform failed: invalid header");
    if (e != cudaSuccess) return ctx->cuda_fail(e, "bwt_inverse");
    if (produced) CK(cudaMemcpyAsync(dst, ctx->d_out.p, produced, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out_n = produced;
    return 0;
}

}  // extern "C"
