// NEXT ROUND — not part of the build (kanzi-go_b200/build.py does not list this directory).
//
// TEXT forward (kz_text_core.cuh forward_walk) split into the three phases a GPU wants; tests/test_text_par_host.py checks on the CPU
// that the phases reproduce forward_walk byte for byte, so that the kernels of the next round only have to wrap them:
//   A. per byte, data parallel: the candidate words — maximal letter runs of 2..31 letters that end on a delimiter — with both hashes
//      (as written, first letter case-flipped). A word is identified by the position p of the delimiter that ends it.
//   B. per candidate, in order, one thread per block: the dictionary walk (look-up, insert, recycle). This is the only serial part, and it
//      touches one map slot, one entry and at most 31 bytes per candidate instead of every byte of the block; its inputs are known
//      ahead of time, so the map slots of the next candidates can be prefetched.
//   C. per found word, data parallel after a prefix sum: the literal span in front of the word (with the escapes of emitSymbols and the
//      "single space between two dictionary words is implied" rule), the 1-3 byte word index, and the size checks of Forward :1338-1344,
//      :1382-1389.
#pragma once
#include "../kz_text_core.cuh"

namespace kz {
namespace textc {

struct Candidate {
    int32_t end;      // position of the delimiter that ends the word
    int32_t start;    // first letter (delim_anchor + 1)
    uint32_t h1, h2;  // hash as written / with the case of the first letter flipped
};
struct Found {
    int32_t end;     // delimiter position of the candidate
    int32_t start;
    int32_t length;
    int32_t index;   // dictionary index
    int32_t flip;    // 1 when the match was found through h2
};

// ---- phase A. prev_nontext[p] = last position < p holding a non-letter, -1 if none (a max-scan on the GPU).
// Returns true and fills c when position p ends a candidate word.
KZ_TEXT_HD bool candidate_at(const uint8_t* src, int count, int p, int prev_nontext, Candidate* c) {
    if (p >= count || is_text(src[p])) return false;
    const int a = prev_nontext;  // delim_anchor of forward_walk when it reaches p
    if (!(p > a + 2 && is_delimiter(src[p]))) return false;
    const int length = p - a - 1;
    if (length > MAX_WORD_LENGTH) return false;
    const uint8_t val = src[a + 1];
    uint32_t h1 = hash_step(HASH1, val);
    uint32_t h2 = hash_step(HASH1, (uint32_t)val ^ 0x20u);
    for (int i = a + 2; i < p; i++) {
        const uint32_t h = (uint32_t)src[i] * HASH2;
        h1 = (h1 * HASH1) ^ h;
        h2 = (h2 * HASH1) ^ h;
    }
    c->end = p;
    c->start = a + 1;
    c->h1 = h1;
    c->h2 = h2;
    return true;
}

// ---- phase B: the dictionary pass over the candidates of a block, in order. Writes the found words; returns their number.
KZ_TEXT_HD int dictionary_pass(const uint8_t* src, const Candidate* cand, int ncand, Dict& D, Found* found) {
    int words = D.static_size;
    int nf = 0;
    for (int k = 0; k < ncand; k++) {
        const Candidate& c = cand[k];
        const int32_t length = c.end - c.start;
        int pe = -1;
        const int pe1 = D.map[c.h1 & D.hash_mask];
        if (pe1 >= 0 && D.list[pe1].hash == (int32_t)c.h1 && (D.list[pe1].data >> 24) == length) {
            pe = pe1;
        } else {
            const int pe2 = D.map[c.h2 & D.hash_mask];
            if (pe2 >= 0 && D.list[pe2].hash == (int32_t)c.h2 && (D.list[pe2].data >> 24) == length) pe = pe2;
        }
        if (pe >= 0 && !same_words(entry_bytes(D, D.list[pe], src) + 1, src + c.start + 1, length - 1)) pe = -1;
        if (pe < 0) {
            if ((length > 3 || (length == 3 && words < THRESHOLD2)) && pe1 < 0) add_word(D, words, c.start, length, c.h1);
        } else {
            Found f;
            f.end = c.end;
            f.start = c.start;
            f.length = (int32_t)(D.list[pe].data >> 24);
            f.index = (int32_t)(D.list[pe].data & MASK_LENGTH);
            f.flip = pe != pe1 ? 1 : 0;
            found[nf++] = f;
        }
    }
    return nf;
}

// ---- phase C helpers
// bytes emitSymbols writes for one input byte (:1415-1487)
KZ_TEXT_HD int symbol_cost(uint8_t cur, bool is_crlf) {
    if (cur == ESCAPE_TOKEN1) return 2;
    if (cur == CR) return is_crlf ? 0 : 1;
    return cur >= 0x80 ? 2 : 1;
}
KZ_TEXT_HD int index_cost(int widx) {
    widx++;
    return widx >= THRESHOLD3 ? (widx >= THRESHOLD4 ? 3 : 2) : 1;
}
// literal span in front of found word k: [lit_from, lit_to) of the source; the reference emits src[emit_anchor : delim_anchor + 1] unless
// it is exactly one space (emit_anchor == delim_anchor and that byte is a space), delim_anchor = start - 1
KZ_TEXT_HD void literal_span(const uint8_t* src, int emit_anchor, const Found& f, int* lit_from, int* lit_to) {
    const int delim_anchor = f.start - 1;
    *lit_from = emit_anchor;
    *lit_to = delim_anchor + 1;
    if (emit_anchor == delim_anchor && src[delim_anchor] == ' ') *lit_to = emit_anchor;  // the single space is implied
}

// Sequential statement of phase C (the GPU version replaces the running sums by prefix sums over `cost_prefix` and over the found words):
// cost_prefix[i] = sum of symbol_cost(src[j]) for j < i. Returns the encoded length or -1 (skip), writing dst.
KZ_TEXT_HD int emit_pass(const uint8_t* src, int count, uint8_t mode, const Found* found, int nf, uint8_t* dst) {
    const bool is_crlf = (mode & MASK_CRLF) != 0;
    const int dst_end = count, dst_end3 = dst_end - 3;
    dst[0] = mode;
    int src_idx = 0, dst_idx = 1, emit_anchor = 0;
    while (src_idx < count && src[src_idx] == ' ') {  // leading spaces are copied verbatim (:1268-1273)
        dst[dst_idx++] = ' ';
        src_idx++;
        emit_anchor++;
    }
    if (src_idx >= count) return -1;
    for (int k = 0; k < nf; k++) {
        const Found& f = found[k];
        int from, to;
        literal_span(src, emit_anchor, f, &from, &to);
        if (to > from) dst_idx += emit_symbols(src + from, to - from, dst + dst_idx, dst_end - dst_idx, is_crlf);
        if (dst_idx >= dst_end3) return -1;
        if (f.flip) dst[dst_idx++] = MASK_FLIP_CASE;
        dst_idx += emit_word_index2(dst + dst_idx, f.index);
        emit_anchor = f.start + f.length;
    }
    dst_idx += emit_symbols(src + emit_anchor, count - emit_anchor, dst + dst_idx, dst_end - dst_idx, is_crlf);
    if (dst_idx > dst_end) return -1;
    return dst_idx;
}

}  // namespace textc
}  // namespace kz
