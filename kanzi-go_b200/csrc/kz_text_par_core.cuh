// TEXT forward (kz_text_core.cuh forward_walk) split into the phases a GPU wants; tests/test_text_par_host.py checks on the CPU
// that the phases reproduce forward_walk byte for byte; the kernels of kz_text_par.cu wrap them:
//   A. per byte, data parallel: the candidate words — maximal letter runs of 2..31 letters that end on a delimiter — with both hashes
//      (as written, first letter case-flipped). A word is identified by the position p of the delimiter that ends it.
//   B. the dictionary. Serial statement: dictionary_pass (look-up, insert, recycle, one candidate after the other). Parallel statement
//      (dictionary rounds, below): as long as the entry list does not wrap around (fewer than 2^19 - 1024 inserts in a block) a map slot
//      that has been filled keeps its word for ever (add_word only fires on an EMPTY slot), so "which word owns slot s" is the first
//      candidate of that slot that is inserted at all, and a candidate is inserted iff it qualifies by length, its slot is still empty
//      and it is not found through its case-flipped hash. Those conditions refer to EARLIER candidates only, so the true outcome is the
//      unique fixed point of  ins' = F(ins)  where F looks every candidate up in the slot owners that `ins` implies; every round is data
//      parallel (atomicMin per slot, one thread per candidate) and round r is right for every candidate whose decision hangs on a chain
//      of at most r earlier decisions (in text: a handful). Two quirks of the reference are part of F: fresh list entries carry hash 0,
//      so EVERY insert empties map slot 0 (:1302-1310 via add_word), and words of 3 letters are only inserted while fewer than
//      16384 - 1024 words were (THRESHOLD2, :1303).
//   C. per found word, data parallel after a prefix sum: the literal span in front of the word (with the escapes of emitSymbols and the
//      "single space between two dictionary words is implied" rule), the 1-3 byte word index, and the size checks of Forward :1338-1344,
//      :1382-1389.
#pragma once
#include "kz_text_core.cuh"

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

// ---- phase B, parallel statement: dictionary rounds ------------------------------------------------------------------------------------------------
static const int32_t OCC_EMPTY = 0x7FFFFFFF;
static const int Z0_MAX = 16;
struct SlotZero {  // the inserted candidates whose slot is 0, with the next insert behind each of them (slot 0 is emptied by every insert)
    int32_t n;
    int32_t first_ins;  // first inserted candidate of the block (INT32_MAX: none): the static word of slot 0 lives until then
    int32_t idx[Z0_MAX];
    int32_t nx[Z0_MAX];
};
struct DictView {
    const uint8_t* src;
    const Candidate* cand;
    const int32_t* owner;  // [mask + 1]: OCC_EMPTY, a static word -(index + 1), or the first inserted candidate of the slot (slot 0: static / empty only)
    uint32_t mask;
    const Entry* sdict;
    const uint8_t* swords;
    const SlotZero* z0;
};
// what map[slot] holds when candidate k is looked up
KZ_TEXT_HD int32_t occupant(const DictView& V, uint32_t slot, int k) {
    if (slot != 0) {
        const int32_t o = V.owner[slot];
        return (o >= 0 && o >= k) ? OCC_EMPTY : o;
    }
    const int32_t st = V.owner[0];
    if (st != OCC_EMPTY && k <= V.z0->first_ins) return st;
    for (int z = 0; z < V.z0->n && z < Z0_MAX; z++)
        if (V.z0->idx[z] < k && k <= V.z0->nx[z]) return V.z0->idx[z];
    return OCC_EMPTY;
}
KZ_TEXT_HD bool occupant_is(const DictView& V, int32_t o, uint32_t hash, int32_t length) {
    if (o == OCC_EMPTY) return false;
    if (o < 0) {
        const Entry& e = V.sdict[-(o + 1)];
        return e.hash == (int32_t)hash && (e.data >> 24) == length;
    }
    const Candidate& c = V.cand[o];
    return c.h1 == hash && c.end - c.start == length;
}
KZ_TEXT_HD const uint8_t* occupant_bytes(const DictView& V, int32_t o) {
    if (o < 0) return V.swords + (~V.sdict[-(o + 1)].ptr);
    return V.src + V.cand[o].start;
}
// One candidate of one round. ins_before = inserted candidates in front of k under the previous round's decisions. Returns the new decision;
// *found_occ = the word it is (OCC_EMPTY: not in the dictionary), *flip = found through the case-flipped hash.
// inverse = the decoder's walk (Inverse :1544-1581): no look-up through the case-flipped hash, and every word qualifies by length while
// fewer than THRESHOLD2 words exist.
KZ_TEXT_HD bool eval_candidate(const DictView& V, int k, uint32_t ins_before, int32_t* found_occ, int32_t* flip, bool inverse = false) {
    const Candidate& c = V.cand[k];
    const int32_t length = c.end - c.start;
    const int32_t o1 = occupant(V, c.h1 & V.mask, k);
    int32_t pe = OCC_EMPTY;
    *flip = 0;
    if (occupant_is(V, o1, c.h1, length)) {
        pe = o1;
    } else if (!inverse) {
        const int32_t o2 = occupant(V, c.h2 & V.mask, k);
        if (occupant_is(V, o2, c.h2, length)) {
            pe = o2;
            *flip = 1;
        }
    }
    if (pe != OCC_EMPTY && !same_words(occupant_bytes(V, pe) + 1, V.src + c.start + 1, length - 1)) pe = OCC_EMPTY;
    *found_occ = pe;
    if (pe != OCC_EMPTY) return false;
    const int words = STATIC_WORDS + (int)ins_before;
    if (inverse) return (length > 3 || words < THRESHOLD2) && o1 == OCC_EMPTY;
    return (length > 3 || (length == 3 && words < THRESHOLD2)) && o1 == OCC_EMPTY;
}
KZ_TEXT_HD int32_t occupant_index(int32_t o, const uint32_t* rank) { return o < 0 ? -(o + 1) : STATIC_WORDS + (int32_t)rank[o]; }

// ---- inverse (textCodec2.Inverse :1513-1718, bitstream version 6), the pieces of the data-parallel statement ---------------------------------
// The encoded stream is a sequence of tokens whose length follows from their first one or two bytes: a letter or any other byte below 0x80
// (1), the escape 0x0F + one raw byte (2), a word index of 1..3 bytes, optionally behind the flip marker 0x80. Token starts are the orbit of
// position 1 under i -> i + token_len(i): a scan over 4-state transition functions ("bytes left in the current token"). The decoder's
// dictionary grows from the letter runs of the stream exactly like the encoder's (same rounds, no flipped look-up); an index token is
// the k-th inserted word or a static word; output offsets are a prefix sum of the token output lengths.
enum { TK_LETTER = 0, TK_LIT = 1, TK_ESC = 2, TK_IDX = 3 };
KZ_TEXT_HD int idx_len(uint32_t v7) { return v7 < 64 ? 1 : (v7 < 112 ? 2 : 3); }
KZ_TEXT_HD int token_len(const uint8_t* src, int i) {  // src[i + 1] must be readable
    const uint8_t cur = src[i];
    if (cur < 0x80) return cur == ESCAPE_TOKEN1 ? 2 : 1;
    if (cur == MASK_FLIP_CASE) return 1 + idx_len(src[i + 1] & 0x7Fu);
    return idx_len(cur & 0x7Fu);
}
KZ_TEXT_HD int token_kind(uint8_t cur) {
    if (is_text(cur)) return TK_LETTER;
    if (cur >= 0x80) return TK_IDX;
    return cur == ESCAPE_TOKEN1 ? TK_ESC : TK_LIT;
}
// index token at p (all its bytes are inside the stream): the dictionary index (already decremented) and the flip mask
KZ_TEXT_HD int token_index(const uint8_t* src, int p, int* flip) {
    int q = p;
    uint8_t cur = src[q++];
    *flip = 0;
    if (cur == MASK_FLIP_CASE) {
        *flip = 1;
        cur = src[q++];
    }
    int idx = cur & 0x7F;
    if (idx >= 64) {
        if (idx >= 112) idx = ((idx & 0x0F) << 16) | ((int)src[q] << 8) | src[q + 1];
        else idx = ((idx & 0x1F) << 8) | src[q];
    }
    return idx - 1;
}
// transition of the token automaton at byte i: state = bytes still to skip before the next token start (0 = a token starts here)
KZ_TEXT_HD uint32_t token_step(const uint8_t* src, int i, uint32_t state) { return state ? state - 1 : (uint32_t)token_len(src, i) - 1; }

}  // namespace textc
}  // namespace kz
