// Batch entry points: one cgo call per batch of blocks instead of one goroutine per block
// (io/CompressedStream.go:621-710 Writer.processBlock, :1614-1744 Reader.processBlock). Lives at v2/gpu/batch.go.
package gpu

/*
#include "kanzi_b200.h"
*/
import "C"

import (
	"time"
	"unsafe"

	kanzi "github.com/flanglet/kanzi-go/v2"
	"github.com/flanglet/kanzi-go/v2/internal"
)

// EncodedBlock is what encodingTask.encode leaves behind for the ordered commit (:951-976).
type EncodedBlock struct {
	Bits []byte // block-local bit string, zero padded
	N    uint64 // "written" (:914)
}

// EncodeBlocks replaces the goroutine fan-out of Writer.processBlock (:658-701). slab holds nbTasks blocks of lens[i] bytes at a
// distance of stride bytes (the Writer's buffers are carved from one kz_alloc_pinned slab). The four listener events of every block
// (:766-771 before transform, :850-855 after transform, :889-894 before entropy, :916-932 after entropy) are emitted here with the
// batch's timestamps: the post-transform length and the checksum are read back from the head of each block's bit string
// (mode byte, optional skip-flag byte, length on 1..4 bytes, checksum; :869-886), so the ABI needs no extra outputs.
func (c *Context) EncodeBlocks(transformType uint64, entropyType uint32, ckBits uint, slab []byte, stride uint64, lens []uint32, firstBlockID int,
	listeners []kanzi.Listener) ([]EncodedBlock, error) {
	n := len(lens)
	maxLen := uint32(0)
	for _, l := range lens {
		if l > maxLen {
			maxLen = l
		}
	}
	outStride := uint64(C.kz_max_block_output(C.size_t(maxLen)))
	out := make([]byte, outStride*uint64(n))
	bits := make([]C.uint64_t, n)
	status := make([]C.int32_t, n)
	hashType := kanzi.EVT_HASH_NONE
	if ckBits == 32 {
		hashType = kanzi.EVT_HASH_32BITS
	} else if ckBits == 64 {
		hashType = kanzi.EVT_HASH_64BITS
	}
	t0 := time.Now()
	for i := 0; i < n && len(listeners) > 0; i++ { // the checksum is not known yet: the reference computes it before this event, here it follows
		notify(listeners, kanzi.NewEvent(kanzi.EVT_BEFORE_TRANSFORM, firstBlockID+i, int64(lens[i]), 0, kanzi.EVT_HASH_NONE, t0))
	}
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.kz_encode_blocks(c.h, C.uint64_t(transformType), C.uint32_t(entropyType), C.uint32_t(ckBits),
		(*C.uint8_t)(unsafe.Pointer(&slab[0])), C.uint64_t(stride), (*C.uint32_t)(unsafe.Pointer(&lens[0])), C.uint32_t(n),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(outStride), &bits[0], &status[0])
	if rc != 0 {
		return nil, c.err(rc)
	}
	t1 := time.Now()
	res := make([]EncodedBlock, n)
	for i := 0; i < n; i++ {
		if status[i] != 0 {
			return nil, c.err(C.int(status[i]))
		}
		b := out[uint64(i)*outStride : uint64(i+1)*outStride]
		res[i] = EncodedBlock{Bits: b, N: uint64(bits[i])}
		if len(listeners) == 0 {
			continue
		}
		post, ck := parseBlockHead(b, ckBits)
		notify(listeners, kanzi.NewEvent(kanzi.EVT_AFTER_TRANSFORM, firstBlockID+i, int64(post), ck, hashType, t1))
		notify(listeners, kanzi.NewEvent(kanzi.EVT_BEFORE_ENTROPY, firstBlockID+i, int64(post), ck, hashType, t1))
		notify(listeners, kanzi.NewEvent(kanzi.EVT_AFTER_ENTROPY, firstBlockID+i, int64((uint64(bits[i])+7)>>3), ck, hashType, t1))
	}
	return res, nil
}

// Commit is the ordered section of encodingTask.encode (:951-976), unchanged.
func Commit(obs kanzi.OutputBitStream, blocks []EncodedBlock) {
	for _, b := range blocks {
		lw := uint(3)
		if b.N >= 8 {
			lw = uint(internal.Log2NoCheck(uint32(b.N>>3)) + 4)
		}
		obs.WriteBits(uint64(lw-3), 5)
		obs.WriteBits(b.N, lw)
		obs.WriteArray(b.Bits, uint(b.N))
	}
}

// DecodeBlocks replaces the decode goroutines of Reader.processBlock: in holds the block bit strings extracted by the serial
// section (:1816-1852), byte aligned at off[i], bits[i] long.
func (c *Context) DecodeBlocks(transformType uint64, entropyType uint32, ckBits uint, in []byte, off, bits []uint64, blockSize uint32,
	out []byte, outStride uint64) ([]uint32, error) {
	n := len(off)
	outLen := make([]uint32, n)
	status := make([]C.int32_t, n)
	c.mu.Lock()
	defer c.mu.Unlock()
	rc := C.kz_decode_blocks(c.h, C.uint64_t(transformType), C.uint32_t(entropyType), C.uint32_t(ckBits), (*C.uint8_t)(unsafe.Pointer(&in[0])),
		(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint64_t)(unsafe.Pointer(&bits[0])), C.uint32_t(n), C.uint32_t(blockSize),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(outStride), (*C.uint32_t)(unsafe.Pointer(&outLen[0])), &status[0])
	if rc != 0 {
		return nil, c.err(rc)
	}
	for i := 0; i < n; i++ {
		if status[i] != 0 {
			return nil, c.err(C.int(status[i]))
		}
	}
	return outLen, nil
}

// parseBlockHead reads post-transform length and checksum back from a block-local bit string (:869-886).
func parseBlockHead(b []byte, ckBits uint) (uint64, uint64) {
	mode := b[0]
	p := 1
	if mode&0x80 == 0 && mode&0x10 != 0 { // not a copy block, _TRANSFORMS_MASK: a skip-flag byte follows
		p++
	}
	dataSize := int((mode>>5)&3) + 1
	var post uint64
	for i := 0; i < dataSize; i++ {
		post = post<<8 | uint64(b[p+i])
	}
	p += dataSize
	var ck uint64
	for i := 0; i < int(ckBits/8); i++ {
		ck = ck<<8 | uint64(b[p+i])
	}
	return post, ck
}

func notify(listeners []kanzi.Listener, evt *kanzi.Event) {
	defer func() { _ = recover() }() // panics in block listeners are ignored (:979-991)
	for _, l := range listeners {
		l.ProcessEvent(evt)
	}
}
