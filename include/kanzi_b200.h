/*
 * kanzi_b200.h — C ABI of the B200-native block compression engine (libkanzi_b200.so).
 *
 * Drop-in boundary for the per-block encode/decode path of flanglet/kanzi-go (bitstream format version 6).
 * Every entry point names the reference interface it replaces (paths relative to the reference's v2/ directory).
 * The Go host binds these through cgo; tests and bench.py bind them through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are HOST memory unless the name ends in _device;
 *   - return value 0 = success; negative = -(kanzi error code, Definitions.go:25-46); positive 1 from
 *     kz_transform_forward = "transform asks to be skipped" (the reference returns a non-nil error, Sequence.go:100-105);
 *   - no entry point aborts or throws; kz_last_error() gives the message of the last failure on that context;
 *   - there is NO CPU fallback: if no CUDA device is usable kz_init fails with -ERR_CREATE_COMPRESSOR.
 *   - entropy type ids (entropy/EntropyCodecFactory.go:26-35): NONE 0, HUFFMAN 1, RANGE 4, ANS0 5, ANS1 8.
 *   - threads: a context is used by ONE host thread at a time; different contexts are independent (own CUDA stream, own scratch) and may be
 *     driven by different threads (goroutines locked to OS threads or not: every call selects its device itself) on the same GPU at the same
 *     time - their kernels run side by side (tests/test_gpu_contexts.py; bench.py times the -l 3 and -l 5 round trips of a step that way);
 *   - subsets of the v6 format the decoders accept (anything else answers an error, never wrong bytes; the Go caller keeps its own path then):
 *     ANS0 / ANS1 log range <= 12 (what every kanzi encoder writes; the format allows 8..15, entropy/ANSRangeCodec.go:606-609), ROLZ
 *     logPosChecks = 4 (what `-t ROLZ` writes), TEXT encoding 2 only (entropy NONE / ANS0 / HUFFMAN / RANGE, transform/Factory.go:99-119),
 *     bitstream version 6;
 *   - transform ids (transform/Factory.go:31-50): NONE 0, BWT 1, BWTS 2, LZ 3, ZRLT 6, MTFT 7, RANK 8, ROLZ 11, LZX 16; a "transform48"
 *     word packs up to 8 six-bit ids, first transform in the top 6 bits (Factory.go:26-54).
 */
#ifndef KANZI_B200_H
#define KANZI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kz_ctx kz_ctx;

/* kanzi error codes (Definitions.go:25-46) */
enum {
    KZ_ERR_MISSING_PARAM = 1, KZ_ERR_BLOCK_SIZE = 2, KZ_ERR_INVALID_CODEC = 3, KZ_ERR_CREATE_COMPRESSOR = 4,
    KZ_ERR_CREATE_DECOMPRESSOR = 5, KZ_ERR_WRITE_FILE = 12, KZ_ERR_PROCESS_BLOCK = 13, KZ_ERR_CREATE_CODEC = 14,
    KZ_ERR_INVALID_FILE = 15, KZ_ERR_STREAM_VERSION = 16, KZ_ERR_INVALID_PARAM = 18, KZ_ERR_CRC_CHECK = 19,
    KZ_ERR_UNKNOWN = 127
};

enum { KZ_E_NONE = 0, KZ_E_HUFFMAN = 1, KZ_E_RANGE = 4, KZ_E_ANS0 = 5, KZ_E_ANS1 = 8 };
enum { KZ_T_NONE = 0, KZ_T_BWT = 1, KZ_T_BWTS = 2, KZ_T_LZ = 3, KZ_T_ZRLT = 6, KZ_T_MTFT = 7, KZ_T_RANK = 8, KZ_T_EXE = 9, KZ_T_TEXT = 10, KZ_T_ROLZ = 11, KZ_T_MM = 15, KZ_T_LZX = 16, KZ_T_UTF = 17, KZ_T_PACK = 18, KZ_T_DNA = 19 };

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int kz_device_count(void);
/* One context per GPU (one process per GPU in multi-GPU runs). Owns streams and reusable HBM scratch. */
int kz_init(int device, kz_ctx** out);
void kz_destroy(kz_ctx* ctx);
const char* kz_last_error(kz_ctx* ctx);
/* Page-locked host buffers for the Writer/Reader block buffers (io/CompressedStream.go:384-394 allocates them
 * with make([]byte)); pinned memory lets the block path overlap PCIe copies with kernels. */
void* kz_alloc_pinned(size_t bytes);
void kz_free_pinned(void* p);
/* the context's cudaStream_t (all kernels of the context are launched on it), for event timing by the caller */
void* kz_cuda_stream(kz_ctx* ctx);
/* number of kernels launched by this context since the last call with reset != 0 (bench.py's gpu_launches) */
uint64_t kz_launch_count(kz_ctx* ctx, int reset);

/* ---- kanzi.EntropyEncoder / kanzi.EntropyDecoder (Definitions.go:154-179), one block ----------------------- */
/* Replaces EntropyEncoder.Write(block) as called by encodingTask.encode (io/CompressedStream.go:906): encodes
 * src[0:n] with the codec `type` into a fresh bit string. dst receives ceil(out_bits/8) bytes, zero padded; the Go
 * shim then does obs.WriteArray(dst, out_bits). Implementations: entropy/ANSRangeCodec.go:274 (ANS0/ANS1),
 * entropy/HuffmanCodec.go:390, entropy/RangeCodec.go:223, entropy/NullEntropyCodec.go:43. */
int kz_entropy_encode(kz_ctx* ctx, uint32_t type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, uint64_t* out_bits);
/* Replaces EntropyDecoder.Read(block) (io/CompressedStream.go:1952): decodes exactly n bytes from the bit string
 * src (src_bytes bytes, the block-local buffer the decodingTask already holds, :1875); *consumed_bits = bits read. */
int kz_entropy_decode(kz_ctx* ctx, uint32_t type, const uint8_t* src, size_t src_bytes, uint8_t* dst, size_t n, uint64_t* consumed_bits);

/* ---- kanzi.ByteTransform (Definitions.go:78-91), one block ------------------------------------------------ */
/* Replaces ByteTransform.Forward(src,dst) for one transform id (transform/BWTBlockCodec.go:78, LZCodec.go:249 ...).
 * *data_type is ctx["dataType"] (in/out, internal/Global.go DataType). Returns 1 when the transform declines. */
int kz_transform_forward(kz_ctx* ctx, uint64_t type, int* data_type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
/* Replaces ByteTransform.Inverse(src,dst) (BWTBlockCodec.go:141, LZCodec.go:621 ...). */
int kz_transform_inverse(kz_ctx* ctx, uint64_t type, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
/* Replaces ByteTransform.MaxEncodedLen (BWTBlockCodec.go:228, LZCodec.go:935). */
size_t kz_transform_max_encoded_len(uint64_t type, size_t n);

/* ---- block dispatch: Writer.processBlock / Reader.processBlock fan-out (io/CompressedStream.go:621-710, -------
 *      :1614-1744) handing ALL blocks of a batch to the GPU in one call ------------------------------------- */
/* Replaces the goroutine-per-block encodingTask.encode (io/CompressedStream.go:729-933, everything before the
 * ordered commit). Block b is slab[b*block_stride : +block_len[b]]. On return out[b*out_stride ...] holds block b's
 * block-local bit string (mode byte, length, optional checksum, entropy coded data), zero padded, and out_bits[b]
 * its exact bit length ("written", :914); status[b] = 0 or -(error code). The caller commits them in order with
 * the 5-bit/lw-bit prefixes (:951-976). out_stride must be >= kz_max_block_output(max block_len). */
int kz_encode_blocks(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t checksum_bits, const uint8_t* slab,
                     uint64_t block_stride, const uint32_t* block_len, uint32_t nblocks, uint8_t* out, uint64_t out_stride,
                     uint64_t* out_bits, int32_t* status);
/* Replaces the concurrent part of decodingTask.decode (io/CompressedStream.go:1873-2011). Block b's bit string is
 * in[in_off[b] : in_off[b] + ceil(in_bits[b]/8)] (byte aligned, as extracted by the serial section :1816-1852).
 * Decoded bytes go to out[b*out_stride ...], out_len[b] = decoded length. block_size = stream block size (:1896). */
int kz_decode_blocks(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t checksum_bits, const uint8_t* in,
                     const uint64_t* in_off, const uint64_t* in_bits, uint32_t nblocks, uint32_t block_size, uint8_t* out,
                     uint64_t out_stride, uint32_t* out_len, int32_t* status);
size_t kz_max_block_output(size_t block_len);

/* ---- whole stream: io.Writer.Write+Close / io.Reader.Read (io/CompressedStream.go:216-256, :524-619, :1047-1100,
 *      :1556-1612) for an in-memory input, stream header included (:429-519 / :1316-1522). jobs has no meaning on
 *      the GPU (all blocks of the input form one batch per GPU pass). ------------------------------------------ */
int kz_compress_stream(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t block_size, uint32_t checksum_bits,
                       int64_t input_size_hint, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
int kz_decompress_stream(kz_ctx* ctx, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_n);
size_t kz_max_stream_output(size_t n, uint32_t block_size);

/* ---- device-resident variants (inputs and outputs already in HBM; used by multi-GPU pipelines and bench.py's
 *      kernel-only `value`). Pointers are CUDA device pointers on ctx's device. Lengths / offsets are host arrays. */
/* Compresses n bytes at d_src into the complete stream at d_dst (header, blocks, end marker). *out_n = stream bytes. */
int kz_compress_stream_device(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t block_size, uint32_t checksum_bits,
                              int64_t input_size_hint, const void* d_src, size_t n, void* d_dst, size_t cap, size_t* out_n);
/* Decompresses the complete stream of n bytes at d_src (must be followed by >= 64 readable bytes) into d_dst. */
int kz_decompress_stream_device(kz_ctx* ctx, const void* d_src, size_t n, void* d_dst, size_t cap, size_t* out_n);

/* ---- sharded streams (io.Writer / io.Reader with the blocks of ONE stream spread over several GPUs; the NCCL scatter / gather of the
 *      byte ranges is the host side's business, kanzi-go_b200/parallel.py; blocks are independent: io/CompressedStream.go:896-898) ---------- */
/* Block records (5 b | lw b | payload each, :951-976) of the n bytes at d_src: no stream header, no end marker. *out_bits = exact length. */
int kz_compress_fragment_device(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t block_size, uint32_t checksum_bits, const void* d_src, size_t n,
                                void* d_dst, size_t cap, uint64_t* out_bits);
/* Stream header (Writer.writeHeader :429-519) as bytes (out: >= 32 bytes) + its bit length. Host only. */
int kz_stream_header(uint64_t transform48, uint32_t entropy, uint32_t block_size, uint32_t checksum_bits, int64_t input_size_hint, uint8_t* out, uint32_t* out_bits);
/* Ordered, bit-granular commit: concatenates nseg device bit strings (4-byte aligned bases) into d_dst. */
int kz_concat_bits_device(kz_ctx* ctx, uint32_t nseg, const void* const* d_seg, const uint64_t* seg_bits, void* d_dst, size_t cap, uint64_t* total_bits);
/* Header fields and record boundaries of the complete stream at d_src: rec_bit[0 .. *nblocks] (max_blocks + 1 entries). */
int kz_stream_index_device(kz_ctx* ctx, const void* d_src, size_t n, uint32_t max_blocks, uint64_t* transform48, uint32_t* entropy, uint32_t* block_size,
                           uint32_t* checksum_bits, int64_t* input_size, uint64_t* rec_bit, uint32_t* nblocks);
/* Decodes nblocks records starting at bit start_bit of the fragment at d_src (16-byte aligned, >= 64 readable bytes behind frag_bytes). */
int kz_decompress_fragment_device(kz_ctx* ctx, uint64_t transform48, uint32_t entropy, uint32_t block_size, uint32_t checksum_bits, const void* d_src,
                                  size_t frag_bytes, uint64_t start_bit, uint32_t nblocks, void* d_dst, size_t cap, size_t* out_n);

/* ---- instrumentation ------------------------------------------------------------------------------------- */
/* Milliseconds (CUDA events on the context's stream) spent in named kernels since the last reset; names:
 * "ans0_decode", "ans0_encode", "ans0_stats", "ans_walk", "concat", "huf_encode", "huf_decode", ...  Returns
 * the number of launches accumulated, 0 if the name is unknown. Only collected when kz_profile(ctx,1) was set. */
void kz_profile(kz_ctx* ctx, int enable);
uint32_t kz_kernel_time(kz_ctx* ctx, const char* name, double* total_ms);
void kz_profile_reset(kz_ctx* ctx);
/* Algorithmic bytes of a transform stage accumulated since the last reset: name = "stage:fwd:<transform id>" or
 * "stage:inv:<transform id>" (ids of transform/Factory.go:31-50); *bytes_in / *bytes_out = stage input / output bytes summed over
 * the blocks the stage ran on. Returns the number of batches accumulated (0 = unknown name). Only with kz_profile(ctx,1).
 * Also "kernel names" (enumeration): kz_profile_names writes up to cap bytes of newline separated names, returns the full size. */
uint32_t kz_stage_bytes(kz_ctx* ctx, const char* name, uint64_t* bytes_in, uint64_t* bytes_out);
size_t kz_profile_names(kz_ctx* ctx, char* buf, size_t cap);

/* ctx["blockSize"] of the stream the following kz_encode_blocks / kz_transform_forward / kz_transform_inverse calls belong to
 * (io/CompressedStream.go:220,1406). Only TEXT reads it (it sizes its hash map from it, transform/TextCodec.go:1143-1156);
 * 0 (default) = the longest block of the call. kz_decode_blocks and the stream entry points know the value themselves. */
void kz_set_stream_block_size(kz_ctx* ctx, uint64_t block_size);

#ifdef __cplusplus
}
#endif
#endif /* KANZI_B200_H */
