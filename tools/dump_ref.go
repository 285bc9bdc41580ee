// dump_ref.go — writes reference-generated golden streams for tests/golden/ (run it where a Go toolchain and the
// kanzi-go sources exist; neither is available in the build image, see DESIGN.md §2 "parity unpinned").
//
//   cd /path/to/kanzi-go/v2 && go run /path/to/repo/tools/dump_ref.go /path/to/repo/tests/golden
//
// For every (transform, entropy, block size, checksum) combination of tests/test_oracle_pins.py::GOLDEN_CASES it compresses
// the SAME deterministic input the tests build (xorshift32 bytes shaped by a small Zipf table) with the reference's
// io.Writer and stores `ref_<case>.knz`. tests/test_oracle_pins.py compares the oracle's stream with every file it finds
// there, byte for byte: that is the pin that turns "parity unpinned" into "pinned" without touching any other code.
package main

import (
	"bytes"
	"fmt"
	"os"
	"path/filepath"

	kio "github.com/flanglet/kanzi-go/v2/io"
)

type goldenCase struct {
	name      string
	transform string
	entropy   string
	blockSize uint
	checksum  uint
	size      int
}

// keep in sync with GOLDEN_CASES in tests/test_oracle_pins.py
var cases = []goldenCase{
	{"none_ans0", "NONE", "ANS0", 65536, 0, 300000},
	{"none_ans1", "NONE", "ANS1", 65536, 0, 300000},
	{"none_huffman", "NONE", "HUFFMAN", 65536, 32, 300000},
	{"none_range", "NONE", "RANGE", 65536, 64, 300000},
	{"bwt_ans0", "BWT", "ANS0", 65536, 0, 300000},
	{"bwts_ans0", "BWTS", "ANS0", 65536, 0, 300000},
	{"lz_none", "LZ", "NONE", 65536, 0, 300000},
	{"lzx_huffman", "LZX", "HUFFMAN", 65536, 0, 300000},
	{"rolz_none", "ROLZ", "NONE", 65536, 0, 300000},
	{"bwt_rank_zrlt_ans0", "BWT+RANK+ZRLT", "ANS0", 65536, 0, 300000},
	{"bwt_mtft_zrlt_huffman", "BWT+MTFT+ZRLT", "HUFFMAN", 65536, 0, 300000},
	// the complete level chains of app/BlockCompressor.go:665-700
	{"l1", "LZX", "NONE", 65536, 0, 300000},
	{"l2", "DNA+LZ", "HUFFMAN", 65536, 0, 300000},
	{"l3", "TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 65536, 0, 300000},
	{"l4", "TEXT+UTF+EXE+PACK+MM+ROLZ", "NONE", 65536, 0, 300000},
	{"l5", "TEXT+UTF+BWT+RANK+ZRLT", "ANS0", 65536, 32, 300000},
	{"text_none", "TEXT", "NONE", 65536, 0, 300000},
	{"pack_none", "PACK", "NONE", 65536, 0, 300000},
	{"mm_none", "MM", "NONE", 65536, 0, 300000},
}

// goldenInput must match golden_input() in tests/test_oracle_pins.py: xorshift32 seeded with 0x4B414E5A, each output byte is
// words[(x >> 8) % 64][k] walking through a 64-entry table of short pseudo-words, which gives LZ / BWT / entropy stages something to do.
func goldenInput(n int) []byte {
	x := uint32(0x4B414E5A)
	next := func() uint32 {
		x ^= x << 13
		x ^= x >> 17
		x ^= x << 5
		return x
	}
	words := make([][]byte, 64)
	for i := range words {
		l := 2 + int(next()%7)
		w := make([]byte, l)
		for j := range w {
			w[j] = byte('a' + next()%26)
		}
		words[i] = w
	}
	out := make([]byte, 0, n+16)
	for len(out) < n {
		r := next()
		w := words[(r>>8)%64]
		if r%5 == 0 {
			w = words[(r>>16)%8]
		}
		out = append(out, w...)
		out = append(out, ' ')
	}
	return out[:n]
}

func main() {
	if len(os.Args) < 2 {
		fmt.Println("usage: go run dump_ref.go <output dir>")
		os.Exit(1)
	}
	dir := os.Args[1]
	for _, c := range cases {
		src := goldenInput(c.size)
		var buf bytes.Buffer
		w, err := kio.NewWriter(&nopCloser{&buf}, c.transform, c.entropy, c.blockSize, 4, c.checksum, int64(len(src)), false)
		if err != nil {
			panic(err)
		}
		if _, err = w.Write(src); err != nil {
			panic(err)
		}
		if err = w.Close(); err != nil {
			panic(err)
		}
		path := filepath.Join(dir, "ref_"+c.name+".knz")
		if err = os.WriteFile(path, buf.Bytes(), 0o644); err != nil {
			panic(err)
		}
		fmt.Printf("%s: %d -> %d bytes\n", path, len(src), buf.Len())
	}
}

type nopCloser struct{ *bytes.Buffer }

func (nopCloser) Close() error { return nil }
