"""kanzi-go_b200 — B200-native drop-in for the per-block encode/decode path of flanglet/kanzi-go.

Python host-side mirror of the reference's operator interfaces for this path (the Go toolchain is absent from the
build image, so the parity tests and bench drive the C ABI of include/kanzi_b200.h through ctypes; INTEGRATION.md
shows the cgo binding a kanzi-go maintainer would add). Names follow the reference:

  EntropyEncoder / EntropyDecoder   v2/Definitions.go:154-179  (Write(block) / Read(block))
  ByteTransform                     v2/Definitions.go:78-91    (Forward / Inverse / MaxEncodedLen)
  Writer / Reader                   v2/io/CompressedStream.go  (NewWriter :216, NewReader :1047)

There is no CPU fallback: importing works without a GPU (so that the library's symbols can be checked), but creating a
Context without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkanzi_b200.so")

E_NONE, E_HUFFMAN, E_RANGE, E_ANS0, E_ANS1 = 0, 1, 4, 5, 8
T_NONE, T_BWT, T_BWTS, T_LZ, T_ROLZ, T_LZX = 0, 1, 2, 3, 11, 16
T_ZRLT, T_MTFT, T_RANK = 6, 7, 8
T_PACK, T_DNA, T_MM, T_TEXT, T_UTF, T_EXE = 18, 19, 15, 10, 17, 9
ENTROPY_IDS = {"NONE": 0, "HUFFMAN": 1, "FPAQ": 2, "RANGE": 4, "ANS0": 5, "CM": 6, "TPAQ": 7, "ANS1": 8, "TPAQX": 9}
TRANSFORM_IDS = {"NONE": 0, "BWT": 1, "BWTS": 2, "LZ": 3, "RLT": 5, "ZRLT": 6, "MTFT": 7, "RANK": 8, "EXE": 9, "TEXT": 10, "ROLZ": 11,
                 "ROLZX": 12, "SRT": 13, "LZP": 14, "MM": 15, "LZX": 16, "UTF": 17, "PACK": 18, "DNA": 19}

ABI_SYMBOLS = [
    "kz_device_count", "kz_init", "kz_destroy", "kz_last_error", "kz_alloc_pinned", "kz_free_pinned", "kz_cuda_stream", "kz_launch_count",
    "kz_entropy_encode", "kz_entropy_decode", "kz_transform_forward", "kz_transform_inverse", "kz_transform_max_encoded_len",
    "kz_encode_blocks", "kz_decode_blocks", "kz_max_block_output", "kz_compress_stream", "kz_decompress_stream", "kz_max_stream_output",
    "kz_compress_stream_device", "kz_decompress_stream_device", "kz_profile", "kz_kernel_time", "kz_profile_reset", "kz_set_stream_block_size",
    "kz_stage_bytes", "kz_profile_names",
    "kz_compress_fragment_device", "kz_stream_header", "kz_concat_bits_device", "kz_stream_index_device", "kz_decompress_fragment_device",
]


class KanziError(RuntimeError):
    """Mirrors io.IOError{msg, code} (v2/io/CompressedStream.go:56-75)."""

    def __init__(self, code, msg):
        self.code = code
        super().__init__("%s (code %d)" % (msg, code))


def transform_type(names):
    """transform.GetType (v2/transform/Factory.go:288-328): "BWT+RANK+ZRLT" -> 48-bit id word."""
    res, shift, count = 0, 42, 0
    for tok in names.split("+"):
        t = TRANSFORM_IDS.get(tok.upper())
        if t is None:
            raise KanziError(18, "Unknown transform type: '%s'" % tok)
        if t != 0:
            count += 1
            if count > 8:
                raise KanziError(18, "Only 8 transforms allowed: '%s'" % names)
            res |= t << shift
            shift -= 6
    return res


def entropy_type(name):
    """entropy.GetType (v2/entropy/EntropyCodecFactory.go:170-206)."""
    t = ENTROPY_IDS.get(name.upper())
    if t is None:
        raise KanziError(18, "Unsupported entropy codec type: '%s'" % name)
    return t


_lib = None


def load_library(build_if_missing=True):
    """Loads libkanzi_b200.so (building it with nvcc when absent). Fails loudly: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise KanziError(4, "libkanzi_b200.so is missing; run python -c 'import __graft_entry__ as g; g.build()'")
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
    L.kz_device_count.restype = i32
    L.kz_init.argtypes = [i32, C.POINTER(vp)]
    L.kz_destroy.argtypes = [vp]
    L.kz_destroy.restype = None
    L.kz_last_error.argtypes = [vp]
    L.kz_last_error.restype = C.c_char_p
    L.kz_alloc_pinned.argtypes = [sz]
    L.kz_alloc_pinned.restype = vp
    L.kz_free_pinned.argtypes = [vp]
    L.kz_free_pinned.restype = None
    L.kz_cuda_stream.argtypes = [vp]
    L.kz_cuda_stream.restype = vp
    L.kz_launch_count.argtypes = [vp, i32]
    L.kz_launch_count.restype = u64
    L.kz_entropy_encode.argtypes = [vp, u32, vp, sz, vp, sz, C.POINTER(u64)]
    L.kz_entropy_decode.argtypes = [vp, u32, vp, sz, vp, sz, C.POINTER(u64)]
    L.kz_transform_forward.argtypes = [vp, u64, C.POINTER(i32), vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_transform_inverse.argtypes = [vp, u64, vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_transform_max_encoded_len.argtypes = [u64, sz]
    L.kz_transform_max_encoded_len.restype = sz
    L.kz_encode_blocks.argtypes = [vp, u64, u32, u32, vp, u64, vp, u32, vp, u64, vp, vp]
    L.kz_decode_blocks.argtypes = [vp, u64, u32, u32, vp, vp, vp, u32, u32, vp, u64, vp, vp]
    L.kz_max_block_output.argtypes = [sz]
    L.kz_max_block_output.restype = sz
    L.kz_compress_stream.argtypes = [vp, u64, u32, u32, u32, C.c_int64, vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_decompress_stream.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_max_stream_output.argtypes = [sz, u32]
    L.kz_max_stream_output.restype = sz
    L.kz_compress_stream_device.argtypes = [vp, u64, u32, u32, u32, C.c_int64, vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_decompress_stream_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.kz_profile.argtypes = [vp, i32]
    L.kz_profile.restype = None
    L.kz_kernel_time.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
    L.kz_kernel_time.restype = u32
    L.kz_profile_reset.argtypes = [vp]
    L.kz_profile_reset.restype = None
    L.kz_set_stream_block_size.argtypes = [vp, C.c_uint64]
    L.kz_compress_fragment_device.argtypes = [vp, u64, u32, u32, u32, vp, sz, vp, sz, C.POINTER(u64)]
    L.kz_stream_header.argtypes = [u64, u32, u32, u32, C.c_int64, vp, C.POINTER(u32)]
    L.kz_concat_bits_device.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64), vp, sz, C.POINTER(u64)]
    L.kz_stream_index_device.argtypes = [vp, vp, sz, u32, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_int64), C.POINTER(u64),
                                         C.POINTER(u32)]
    L.kz_decompress_fragment_device.argtypes = [vp, u64, u32, u32, u32, vp, sz, u64, u32, vp, sz, C.POINTER(sz)]
    L.kz_stage_bytes.argtypes = [vp, C.c_char_p, C.POINTER(u64), C.POINTER(u64)]
    L.kz_stage_bytes.restype = u32
    L.kz_profile_names.argtypes = [vp, C.c_char_p, sz]
    L.kz_profile_names.restype = sz
    L.kz_set_stream_block_size.restype = None
    _lib = L
    return L


def _u8(a):
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.uint8)
    return np.frombuffer(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data if a.size else None


class Context:
    """One engine context per GPU (kz_init). All codec objects of a process share it."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.kz_init(device, C.byref(h))
        if rc != 0:
            raise KanziError(-rc, "cannot create the B200 engine on CUDA device %d (no GPU / no driver?) - there is no CPU fallback" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.kz_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise KanziError(-rc, self.lib.kz_last_error(self.h).decode(errors="replace"))
        return rc

    # ---- raw ABI helpers -------------------------------------------------------------------------------------
    def entropy_encode(self, etype, block):
        a = _u8(block)
        cap = self.lib.kz_max_block_output(a.size)
        out = np.empty(cap, np.uint8)
        bits = C.c_uint64(0)
        self._check(self.lib.kz_entropy_encode(self.h, etype, _p(a), a.size, out.ctypes.data, cap, C.byref(bits)))
        return out[: (bits.value + 7) // 8].copy(), bits.value

    def entropy_decode(self, etype, stream, n):
        s = _u8(stream)
        out = np.empty(max(n, 1), np.uint8)
        used = C.c_uint64(0)
        self._check(self.lib.kz_entropy_decode(self.h, etype, _p(s), s.size, out.ctypes.data, n, C.byref(used)))
        return out[:n], used.value

    def set_stream_block_size(self, block_size):
        """ctx["blockSize"] for the following block / single-transform calls (only TEXT reads it); 0 = the longest block of the call"""
        self.lib.kz_set_stream_block_size(self.h, block_size)

    def transform_forward(self, ttype, block, data_type=0):
        """ByteTransform.Forward for one transform id -> (bytes or None when the transform declines, data_type after)."""
        a = _u8(block)
        cap = max(int(self.lib.kz_transform_max_encoded_len(ttype, a.size)), a.size) + 64
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        dt = C.c_int(data_type)
        rc = self._check(self.lib.kz_transform_forward(self.h, ttype, C.byref(dt), _p(a), a.size, out.ctypes.data, cap, C.byref(n)))
        if rc == 1:
            return None, dt.value
        return out[: n.value].copy(), dt.value

    def transform_inverse(self, ttype, block, cap):
        a = _u8(block)
        out = np.empty(max(cap, 1), np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.kz_transform_inverse(self.h, ttype, _p(a), a.size, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def encode_blocks(self, transform48, etype, blocks, checksum_bits=0):
        """blocks: list of byte arrays -> list of (bytes, nbits). One kz_encode_blocks call (Writer.processBlock batch)."""
        nb = len(blocks)
        arrs = [_u8(b) for b in blocks]
        lens = np.array([a.size for a in arrs], np.uint32)
        stride = int((max(lens.max(initial=0), 1) + 15) // 16 * 16)
        slab = np.zeros(stride * max(nb, 1), np.uint8)
        for i, a in enumerate(arrs):
            slab[i * stride: i * stride + a.size] = a
        ostride = int((self.lib.kz_max_block_output(int(lens.max(initial=0))) + 15) // 16 * 16)
        out = np.zeros(ostride * max(nb, 1), np.uint8)
        bits = np.zeros(max(nb, 1), np.uint64)
        status = np.zeros(max(nb, 1), np.int32)
        self._check(self.lib.kz_encode_blocks(self.h, transform48, etype, checksum_bits, slab.ctypes.data, stride, lens.ctypes.data, nb,
                                              out.ctypes.data, ostride, bits.ctypes.data, status.ctypes.data))
        return [(out[i * ostride: i * ostride + (int(bits[i]) + 7) // 8].copy(), int(bits[i])) for i in range(nb)]

    def decode_blocks(self, transform48, etype, payloads, block_size, checksum_bits=0):
        """payloads: list of (bytes, nbits) -> list of decoded byte arrays (Reader.processBlock batch)."""
        nb = len(payloads)
        offs, bits, chunks, o = [], [], [], 0
        for p, nbits in payloads:
            a = _u8(p)
            offs.append(o)
            bits.append(nbits)
            chunks.append(a)
            o += a.size
        blob = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
        offs = np.array(offs, np.uint64)
        bits = np.array(bits, np.uint64)
        ostride = int(block_size + block_size // 2 + 2048)
        out = np.zeros(ostride * max(nb, 1), np.uint8)
        lens = np.zeros(max(nb, 1), np.uint32)
        status = np.zeros(max(nb, 1), np.int32)
        self._check(self.lib.kz_decode_blocks(self.h, transform48, etype, checksum_bits, _p(blob), offs.ctypes.data, bits.ctypes.data, nb, block_size,
                                              out.ctypes.data, ostride, lens.ctypes.data, status.ctypes.data))
        return [out[i * ostride: i * ostride + int(lens[i])].copy() for i in range(nb)]

    def compress(self, data, transform="NONE", entropy="NONE", block_size=4 << 20, checksum_bits=0, input_size=0):
        a = _u8(data)
        t48 = transform_type(transform) if isinstance(transform, str) else transform
        et = entropy_type(entropy) if isinstance(entropy, str) else entropy
        cap = self.lib.kz_max_stream_output(a.size, block_size)
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.kz_compress_stream(self.h, t48, et, block_size, checksum_bits, input_size, _p(a), a.size, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def decompress(self, stream, out_cap):
        s = _u8(stream)
        out = np.empty(max(out_cap, 1), np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.kz_decompress_stream(self.h, _p(s), s.size, out.ctypes.data, out_cap, C.byref(n)))
        return out[: n.value]

    # ---- device-resident entry points (pointers are CUDA device addresses, e.g. torch tensors' data_ptr()) -------
    def compress_device(self, d_src, n, d_dst, cap, transform48, etype, block_size, checksum_bits=0, input_size=0):
        out_n = C.c_size_t(0)
        self._check(self.lib.kz_compress_stream_device(self.h, transform48, etype, block_size, checksum_bits, input_size, d_src, n, d_dst, cap, C.byref(out_n)))
        return out_n.value

    def decompress_device(self, d_src, n, d_dst, cap):
        out_n = C.c_size_t(0)
        self._check(self.lib.kz_decompress_stream_device(self.h, d_src, n, d_dst, cap, C.byref(out_n)))
        return out_n.value

    def compress_host(self, src_ptr, n, dst_ptr, cap, transform48, etype, block_size, checksum_bits=0, input_size=0):
        out_n = C.c_size_t(0)
        self._check(self.lib.kz_compress_stream(self.h, transform48, etype, block_size, checksum_bits, input_size, src_ptr, n, dst_ptr, cap, C.byref(out_n)))
        return out_n.value

    def decompress_host(self, src_ptr, n, dst_ptr, cap):
        out_n = C.c_size_t(0)
        self._check(self.lib.kz_decompress_stream(self.h, src_ptr, n, dst_ptr, cap, C.byref(out_n)))
        return out_n.value

    # ---- sharded streams (one stream, blocks spread over ranks; kanzi-go_b200/parallel.py drives these) ---------------------
    def compress_fragment_device(self, d_src, n, d_dst, cap, transform48, etype, block_size, checksum_bits=0):
        bits = C.c_uint64(0)
        self._check(self.lib.kz_compress_fragment_device(self.h, transform48, etype, block_size, checksum_bits, d_src, n, d_dst, cap, C.byref(bits)))
        return bits.value

    def stream_header(self, transform48, etype, block_size, checksum_bits=0, input_size=0):
        out = np.zeros(32, np.uint8)
        bits = C.c_uint32(0)
        rc = self.lib.kz_stream_header(transform48, etype, block_size, checksum_bits, input_size, out.ctypes.data, C.byref(bits))
        if rc:
            raise KanziError(-rc, "kz_stream_header")
        return out, bits.value

    def concat_bits_device(self, seg_ptrs, seg_bits, d_dst, cap):
        n = len(seg_ptrs)
        ptrs = (C.c_void_p * max(n, 1))(*seg_ptrs)
        bits = (C.c_uint64 * max(n, 1))(*seg_bits)
        total = C.c_uint64(0)
        self._check(self.lib.kz_concat_bits_device(self.h, n, ptrs, bits, d_dst, cap, C.byref(total)))
        return total.value

    def stream_index_device(self, d_src, n, max_blocks):
        t48, et, bs, ck, isz, nb = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_int64(0), C.c_uint32(0)
        rec = (C.c_uint64 * (max_blocks + 1))()
        self._check(self.lib.kz_stream_index_device(self.h, d_src, n, max_blocks, C.byref(t48), C.byref(et), C.byref(bs), C.byref(ck), C.byref(isz), rec,
                                                    C.byref(nb)))
        return {"transform48": t48.value, "entropy": et.value, "block_size": bs.value, "checksum_bits": ck.value, "input_size": isz.value,
                "rec_bit": [int(rec[i]) for i in range(nb.value + 1)], "nblocks": nb.value}

    def decompress_fragment_device(self, d_src, frag_bytes, start_bit, nblocks, d_dst, cap, transform48, etype, block_size, checksum_bits=0):
        out_n = C.c_size_t(0)
        self._check(self.lib.kz_decompress_fragment_device(self.h, transform48, etype, block_size, checksum_bits, d_src, frag_bytes, start_bit, nblocks, d_dst,
                                                           cap, C.byref(out_n)))
        return out_n.value

    def cuda_stream(self):
        return self.lib.kz_cuda_stream(self.h)

    # ---- instrumentation ---------------------------------------------------------------------------------------
    def profile(self, on=True):
        self.lib.kz_profile(self.h, 1 if on else 0)

    def profile_reset(self):
        self.lib.kz_profile_reset(self.h)

    def kernel_time(self, name):
        ms = C.c_double(0)
        n = self.lib.kz_kernel_time(self.h, name.encode(), C.byref(ms))
        return n, ms.value

    def stage_bytes(self, name):
        """-> (batches, bytes_in, bytes_out) of "stage:fwd:<id>" / "stage:inv:<id>" since the last profile_reset"""
        a, b = C.c_uint64(0), C.c_uint64(0)
        k = self.lib.kz_stage_bytes(self.h, name.encode(), C.byref(a), C.byref(b))
        return k, a.value, b.value

    def profile_names(self):
        n = self.lib.kz_profile_names(self.h, None, 0)
        buf = C.create_string_buffer(n + 16)
        self.lib.kz_profile_names(self.h, buf, n + 16)
        return [x for x in buf.value.decode().split("\n") if x]

    def launch_count(self, reset=False):
        return self.lib.kz_launch_count(self.h, 1 if reset else 0)


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


class OutputBitStream:
    """Minimal stand-in for kanzi.OutputBitStream (v2/Definitions.go:119-140) over a growing byte buffer: just what
    the codec mirrors need (WriteArray of the codec's bit string, Written). MSB first, like DefaultOutputBitStream."""

    def __init__(self):
        self._bits = []  # list of (np.uint8 array, nbits)
        self._written = 0

    def write_array(self, data, nbits):
        self._bits.append((np.array(_u8(data)[: (nbits + 7) // 8], copy=True), nbits))
        self._written += nbits

    def written(self):
        return self._written

    def to_bytes(self):
        """Zero padded bytes of everything written so far (bit-exact concatenation)."""
        total = np.zeros((self._written + 7) // 8 + 8, np.uint8)
        pos = 0
        for arr, nbits in self._bits:
            if nbits == 0:
                continue
            bits = np.unpackbits(arr)[:nbits]
            byte0, off = pos // 8, pos % 8
            padded = np.concatenate([np.zeros(off, np.uint8), bits])
            padded = np.concatenate([padded, np.zeros((-len(padded)) % 8, np.uint8)])
            chunk = np.packbits(padded)
            total[byte0: byte0 + len(chunk)] |= chunk
            pos += nbits
        return total[: (self._written + 7) // 8]


class EntropyEncoder:
    """kanzi.EntropyEncoder for the GPU codecs (entropy.NewEntropyEncoder(obs, ctx, type), EntropyCodecFactory.go:91-134)."""

    def __init__(self, bitstream, entropy, ctx=None):
        self.type = entropy_type(entropy) if isinstance(entropy, str) else entropy
        self.bs = bitstream
        self.ctx = ctx or default_context()

    def write(self, block):
        data, nbits = self.ctx.entropy_encode(self.type, block)
        self.bs.write_array(data, nbits)
        return len(block)

    def bit_stream(self):
        return self.bs

    def dispose(self):
        pass


class EntropyDecoder:
    """kanzi.EntropyDecoder (entropy.NewEntropyDecoder(ibs, ctx, type), EntropyCodecFactory.go:45-88). The input
    bit stream is the block-local byte buffer (io/CompressedStream.go:1875) plus a bit cursor."""

    def __init__(self, data, entropy, ctx=None, bit_offset=0):
        if bit_offset % 8:
            raise KanziError(18, "the block-local stream is byte aligned when the entropy decoder is created")
        self.type = entropy_type(entropy) if isinstance(entropy, str) else entropy
        self.data = _u8(data)[bit_offset // 8:]
        self.ctx = ctx or default_context()
        self.consumed = 0

    def read(self, n):
        out, used = self.ctx.entropy_decode(self.type, self.data, n)
        self.consumed = used
        return out

    def dispose(self):
        pass


class ByteTransform:
    """kanzi.ByteTransform (v2/Definitions.go:78-91) for the GPU transforms: transform.New(ctx, type) of a single id."""

    def __init__(self, name, ctx=None):
        self.type = TRANSFORM_IDS[name.upper()]
        self.ctx = ctx or default_context()
        self.data_type = 0

    def forward(self, src):
        """-> transformed bytes, or None when the transform declines (the reference returns an error = "skip me")."""
        out, self.data_type = self.ctx.transform_forward(self.type, src, self.data_type)
        return out

    def inverse(self, src, max_len):
        return self.ctx.transform_inverse(self.type, src, max_len)

    def max_encoded_len(self, n):
        return int(self.ctx.lib.kz_transform_max_encoded_len(self.type, n))


class Writer:
    """io.Writer (CompressedOutputStream): NewWriter(os, transform, entropy, blockSize, jobs, checksum, fileSize, headerless)
    (io/CompressedStream.go:216). write() buffers, close() compresses everything that was written in one GPU batch
    (the reference cuts batches of `jobs` blocks; batches do not change the bytes) and returns nothing; the stream
    bytes are in .getvalue()."""

    def __init__(self, transform="NONE", entropy="NONE", block_size=4 << 20, jobs=1, checksum=0, file_size=0, ctx=None):
        if not (1 <= jobs <= 64):
            raise KanziError(18, "The number of jobs must be in [1..64], got %d" % jobs)
        if block_size > (1 << 30):
            raise KanziError(18, "The block size must be at most 1024 MB")
        if block_size < 1024:
            raise KanziError(18, "The block size must be at least 1024")
        if block_size & 15:
            raise KanziError(18, "The block size must be a multiple of 16")
        self.t48 = transform_type(transform)
        self.etype = entropy_type(entropy)
        self.block_size = block_size
        self.checksum = checksum
        self.file_size = file_size
        self.ctx = ctx or default_context()
        self._chunks = []
        self._out = None

    def write(self, block):
        if self._out is not None:
            raise KanziError(12, "Stream closed")
        self._chunks.append(np.array(_u8(block), copy=True))
        return len(block)

    def close(self):
        if self._out is None:
            data = np.concatenate(self._chunks) if self._chunks else np.zeros(0, np.uint8)
            self._out = self.ctx.compress(data, self.t48, self.etype, self.block_size, self.checksum, self.file_size)

    def getvalue(self):
        self.close()
        return self._out


class Reader:
    """io.Reader (CompressedInputStream): NewReader(is, jobs) (io/CompressedStream.go:1047)."""

    def __init__(self, stream, jobs=1, ctx=None):
        self.stream = _u8(stream)
        self.ctx = ctx or default_context()

    def read_all(self, max_size):
        return self.ctx.decompress(self.stream, max_size)
