// Package gpu binds libkanzi_b200.so (include/kanzi_b200.h) into kanzi-go through cgo. It lives at v2/gpu/ in a kanzi-go tree.
// Not compiled in this repository's build image (no Go toolchain there): source for the maintainer, see INTEGRATION.md.
package gpu

/*
#cgo LDFLAGS: -L${SRCDIR}/../../lib -lkanzi_b200
#include "kanzi_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"sync"
	"unsafe"

	kanzi "github.com/flanglet/kanzi-go/v2"
	"github.com/flanglet/kanzi-go/v2/internal"
)

// Context wraps kz_ctx (one CUDA stream + device scratch). The C side serves one caller at a time per context, so every call below holds mu:
// the per-block goroutines of Writer / Reader may share one Context (their calls queue up; the batch entry points of batch.go are the
// efficient way to give the GPU all blocks at once). Streams that are compressed or decompressed concurrently should each get their own
// Context: different contexts run side by side on the GPU (include/kanzi_b200.h, "threads").
type Context struct {
	h  *C.kz_ctx
	mu sync.Mutex
}

func NewContext(device int) (*Context, error) {
	var h *C.kz_ctx
	if rc := C.kz_init(C.int(device), &h); rc != 0 {
		return nil, fmt.Errorf("kanzi gpu: cannot open CUDA device %d (code %d)", device, -int(rc))
	}
	return &Context{h: h}, nil
}
func (c *Context) Close() {
	c.mu.Lock()
	defer c.mu.Unlock()
	C.kz_destroy(c.h)
	c.h = nil
}

// err is called with mu held (kz_last_error belongs to the failed call)
func (c *Context) err(rc C.int) error {
	return errors.New(C.GoString(C.kz_last_error(c.h)))
}

// EntropyEncoder implements kanzi.EntropyEncoder (v2/Definitions.go:154-165) for ANS0 / ANS1 / HUFFMAN / RANGE / NONE.
type EntropyEncoder struct {
	ctx *Context
	bs  kanzi.OutputBitStream
	typ uint32 // entropy.ANS0_TYPE, entropy.HUFFMAN_TYPE ...
	buf []byte
}

func (e *EntropyEncoder) Write(block []byte) (int, error) {
	if len(block) == 0 {
		return 0, nil
	}
	need := int(C.kz_max_block_output(C.size_t(len(block))))
	if len(e.buf) < need {
		e.buf = make([]byte, need)
	}
	var bits C.uint64_t
	e.ctx.mu.Lock()
	defer e.ctx.mu.Unlock()
	rc := C.kz_entropy_encode(e.ctx.h, C.uint32_t(e.typ), (*C.uint8_t)(unsafe.Pointer(&block[0])), C.size_t(len(block)),
		(*C.uint8_t)(unsafe.Pointer(&e.buf[0])), C.size_t(len(e.buf)), &bits)
	if rc != 0 {
		return 0, e.ctx.err(rc)
	}
	e.bs.WriteArray(e.buf, uint(bits)) // the codec's bit string, bit exact (DefaultOutputBitStream.go:101)
	return len(block), nil
}
func (e *EntropyEncoder) BitStream() kanzi.OutputBitStream { return e.bs }
func (e *EntropyEncoder) Dispose()                         {}

// EntropyDecoder implements kanzi.EntropyDecoder (:168-179). decodingTask.decode already holds the block-local bytes
// (io/CompressedStream.go:1875). entropy.NewEntropyDecoder(ibs, ctx, type) (EntropyCodecFactory.go:45) has no parameter for them, so
// decodingTask.decode puts them into the context map before it calls the factory (ctx["gpuBlock"] = data, ctx["gpuBlockOffset"] = byte offset
// of the entropy data behind the block header, which it has just read from ibs) and the factory's gpu branch calls NewEntropyDecoder below.
type EntropyDecoder struct {
	ctx  *Context
	bs   kanzi.InputBitStream
	typ  uint32
	data []byte // block-local buffer, entropy data starts at byte `off`
	off  int
}

// NewEntropyDecoder is what the gpu branch of entropy.NewEntropyDecoder returns.
func NewEntropyDecoder(c *Context, ibs kanzi.InputBitStream, ctx map[string]any, typ uint32) (*EntropyDecoder, error) {
	data, ok := ctx["gpuBlock"].([]byte)
	if !ok {
		return nil, errors.New("kanzi gpu: ctx[\"gpuBlock\"] is not set (decodingTask.decode sets it before creating the decoder)")
	}
	off, _ := ctx["gpuBlockOffset"].(int)
	return &EntropyDecoder{ctx: c, bs: ibs, typ: typ, data: data, off: off}, nil
}

// NewEntropyEncoder is what the gpu branch of entropy.NewEntropyEncoder returns.
func NewEntropyEncoder(c *Context, obs kanzi.OutputBitStream, typ uint32) *EntropyEncoder {
	return &EntropyEncoder{ctx: c, bs: obs, typ: typ}
}

func (d *EntropyDecoder) Read(block []byte) (int, error) {
	if len(block) == 0 {
		return 0, nil
	}
	var used C.uint64_t
	d.ctx.mu.Lock()
	defer d.ctx.mu.Unlock()
	rc := C.kz_entropy_decode(d.ctx.h, C.uint32_t(d.typ), (*C.uint8_t)(unsafe.Pointer(&d.data[d.off])), C.size_t(len(d.data)-d.off),
		(*C.uint8_t)(unsafe.Pointer(&block[0])), C.size_t(len(block)), &used)
	if rc != 0 {
		return 0, d.ctx.err(rc)
	}
	// keep the caller's bitstream in step: skip the bits the GPU consumed
	for n := uint64(used); n > 0; {
		k := uint(64)
		if n < 64 {
			k = uint(n)
		}
		d.bs.ReadBits(k)
		n -= uint64(k)
	}
	return len(block), nil
}

// Transform implements kanzi.ByteTransform (v2/Definitions.go:78-91) for one transform id of transform/Factory.go:31-50:
// BWT_TYPE 1, BWTS_TYPE 2, LZ_TYPE 3, ZRLT_TYPE 6, MTFT_TYPE 7, RANK_TYPE 8, DICT_TYPE (TEXT) 10, ROLZ_TYPE 11, MM_TYPE 15, LZX_TYPE 16,
// UTF_TYPE 17, PACK_TYPE 18, DNA_TYPE 19. The shim hands ctx["dataType"] in and out through &dt, like the Go transforms do.
type Transform struct {
	ctx *Context
	typ uint64
	dt  *internal.DataType // ctx["dataType"], read by LZ/LZX (LZCodec.go:298-311)
}

func (t *Transform) Forward(src, dst []byte) (uint, uint, error) {
	var n C.size_t
	dt := C.int(*t.dt)
	t.ctx.mu.Lock()
	defer t.ctx.mu.Unlock()
	rc := C.kz_transform_forward(t.ctx.h, C.uint64_t(t.typ), &dt, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)),
		(*C.uint8_t)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)), &n)
	if rc == 1 {
		return 0, 0, errors.New("forward transform skip") // non-nil error = "skip me" (Sequence.go:100-105)
	}
	if rc != 0 {
		return 0, 0, t.ctx.err(rc)
	}
	*t.dt = internal.DataType(dt)
	return uint(len(src)), uint(n), nil
}
func (t *Transform) Inverse(src, dst []byte) (uint, uint, error) {
	var n C.size_t
	t.ctx.mu.Lock()
	defer t.ctx.mu.Unlock()
	rc := C.kz_transform_inverse(t.ctx.h, C.uint64_t(t.typ), (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)),
		(*C.uint8_t)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)), &n)
	if rc != 0 {
		return 0, 0, t.ctx.err(rc) // fatal for the block (Sequence.go:168-171)
	}
	return uint(len(src)), uint(n), nil
}
func (t *Transform) MaxEncodedLen(n int) int { return int(C.kz_transform_max_encoded_len(C.uint64_t(t.typ), C.size_t(n))) }

func (d *EntropyDecoder) BitStream() kanzi.InputBitStream { return d.bs }
func (d *EntropyDecoder) Dispose()                        {}

// NewTransform is what the gpu branch of transform.newToken returns; dt points at the sequence's ctx["dataType"] cell.
func NewTransform(c *Context, typ uint64, dt *internal.DataType) *Transform {
	return &Transform{ctx: c, typ: typ, dt: dt}
}
