// TEXT transform (TextCodec, encoding 2): the per-block state machine, written once for host and device.
// Reference: v2/transform/TextCodec.go — computeTextStats :187-305, detectTextType :307-396, delimiters :408-447, createDictionary :450-489,
// textCodec2.reset :1190-1223, Forward :1225-1398, expandDictionary :1400-1413, emitSymbols :1415-1487, emitWordIndex2 :1489-1511,
// Inverse :1513-1718. Encoding 2 is what the factory picks for the NONE / ANS0 / HUFFMAN / RANGE entropy codecs (Factory.go:100-120).
//
// The kernels of kz_text.cu call these functions from one thread per block (the dictionary is the state of a serial walk over the
// words of the block); tests/test_text_core_host.py compiles the same header with g++ and checks it against the oracle on the CPU,
// so the logic is confirmed even where no GPU is at hand. Nothing here allocates: the caller provides the map and the entry list.
#pragma once
#include <stdint.h>

#ifndef KZ_TEXT_HD
#ifdef __CUDACC__
#define KZ_TEXT_HD __host__ __device__ __forceinline__
#else
#define KZ_TEXT_HD inline
#endif
#endif

namespace kz {
namespace textc {

static const int THRESHOLD1 = 128, THRESHOLD2 = THRESHOLD1 * THRESHOLD1, THRESHOLD3 = 64, THRESHOLD4 = THRESHOLD3 * 128;
static const int MAX_DICT_SIZE = 1 << 19, MAX_WORD_LENGTH = 31, STATIC_WORDS = 1024;
static const int32_t MASK_LENGTH = 0x0007FFFF;
static const uint8_t ESCAPE_TOKEN1 = 0x0F, MASK_FLIP_CASE = 0x80, MASK_NOT_TEXT = 0x80, MASK_CRLF = 0x40, MASK_XML_HTML = 0x20, MASK_DT = 0x0F;
static const uint32_t HASH1 = 0x7FEB352Du, HASH2 = 0x846CA68Bu;  // the reference computes in int32: same bits modulo 2^32
static const uint8_t CR = 0x0D, LF = 0x0A;
static const int64_t NIL = INT64_MIN;

struct Entry {
    int32_t hash;
    int32_t data;  // length << 24 | index
    int64_t ptr;   // >= 0: offset of the word in the block buffer; < 0 (not NIL): ~ptr = offset in the static word list; NIL: no word yet
};

struct Dict {
    int32_t* map;          // [hash_mask + 1] slot -> entry index, -1 = empty
    uint32_t hash_mask;
    Entry* list;           // [MAX_DICT_SIZE] (the reference grows it by doubling; entry i starts as {hash 0, data i, no word})
    int dict_size;         // current logical size of the list
    int static_size;       // 1024
    const uint8_t* swords; // lower-cased letters of the static dictionary
};

KZ_TEXT_HD bool is_lower(uint8_t v) { return v >= 'a' && v <= 'z'; }
KZ_TEXT_HD bool is_upper(uint8_t v) { return v >= 'A' && v <= 'Z'; }
KZ_TEXT_HD bool is_text(uint8_t v) { return is_lower((uint8_t)(v | 0x20)); }
KZ_TEXT_HD bool is_delimiter(uint8_t v) {  // :408-447
    if (v >= ' ' && v <= '/') return true;
    if (v >= ':' && v <= '?') return true;
    return v == '\n' || v == '\r' || v == '\t' || v == '_' || v == '|' || v == '{' || v == '}' || v == '[' || v == ']';
}
KZ_TEXT_HD uint32_t hash_step(uint32_t h, uint32_t c) { return (h * HASH1) ^ (c * HASH2); }

KZ_TEXT_HD uint32_t log2_floor32(uint32_t x) {
    uint32_t r = 0;
    while (x >>= 1) r++;
    return r;
}
// log2 of the hash map size (newTextCodec2WithCtx :1143-1156) and initial list size (reset :1191-1197)
KZ_TEXT_HD uint32_t log_hash_size(uint64_t stream_block_size) {
    uint32_t lg = 13;
    if (stream_block_size >= 32) {
        lg = log2_floor32((uint32_t)(stream_block_size / 32));
        lg = lg > 24 ? 24 : lg;
        lg = lg < 13 ? 13 : lg;
    }
    return lg;
}
KZ_TEXT_HD int initial_dict_size(int count) {
    int size = 1 << 13;
    if (count >= 1024) {
        uint32_t lg = log2_floor32((uint32_t)(count / 128));
        lg = lg > 18 ? 18 : lg;
        lg = lg < 13 ? 13 : lg;
        size = 1 << lg;
    }
    return size;
}

// createDictionary :450-489 over the letters of the static word list (upper case marks a word start). words: in = mixed case,
// out = lower case. Fills entries[0..1024). Returns the number of words.
KZ_TEXT_HD int create_static_dictionary(uint8_t* words, int n, Entry* entries) {
    int anchor = 0, nb = 0;
    uint32_t h = HASH1;
    for (int i = 0; i < n && nb < STATIC_WORDS; i++) {
        if (is_upper(words[i])) {
            if (i > anchor) {
                entries[nb].ptr = ~(int64_t)anchor;
                entries[nb].hash = (int32_t)h;
                entries[nb].data = (int32_t)(((uint32_t)(i - anchor) << 24) | (uint32_t)nb);
                nb++;
                anchor = i;
                h = HASH1;
            }
            words[i] ^= 0x20;
        }
        h = hash_step(h, words[i]);
    }
    if (nb < STATIC_WORDS) {
        entries[nb].ptr = ~(int64_t)anchor;
        entries[nb].hash = (int32_t)h;
        entries[nb].data = (int32_t)(((uint32_t)(n - anchor) << 24) | (uint32_t)nb);
        nb++;
    }
    return nb;
}

// internal/Global.go:346-419 DetectSimpleType (histogram of uint32 counters)
KZ_TEXT_HD int detect_simple_type_u32(uint32_t count, const uint32_t* f) {
    if (count == 0) return 0;
    const char* DNA = "acgntuACGNTU";
    const char* NUM = "0123456789+-*/=,.:; ";
    const char* B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    uint64_t sum = 0;
    for (int i = 0; i < 12; i++) sum += f[(uint8_t)DNA[i]];
    if (sum > (uint64_t)count - (uint64_t)count / 12) return 6;
    sum = 0;
    for (int i = 0; i < 20; i++) sum += f[(uint8_t)NUM[i]];
    if (sum == (uint64_t)count) return 4;
    sum = 0;
    for (int i = 0; i < 64; i++) sum += f[(uint8_t)B64[i]];
    if (sum + f[0x3D] == (uint64_t)count) return 5;
    sum = 0;
    for (int i = 0; i < 256; i++)
        if (f[i] > 0) sum++;
    if (sum == 256) return 7;
    if (sum <= 4) return 9;
    return 0;
}

// detectTextType :307-396. freqs1[(a << 8) | b] = occurrences of byte b after byte a.
KZ_TEXT_HD uint8_t detect_text_type(const uint32_t* freqs0, const uint32_t* freqs1, int count) {
    const int dt = detect_simple_type_u32((uint32_t)count, freqs0);
    if (dt != 0) return (uint8_t)(MASK_NOT_TEXT | dt);
    uint64_t sum = (uint64_t)freqs0[0xC0] + freqs0[0xC1];
    for (int i = 0xF5; i < 256; i++) sum += freqs0[i];
    if (sum != 0) return MASK_NOT_TEXT;
    uint64_t sum2 = 0;
    for (int i = 0; i < 256; i++) {
        if (i < 0xA0 || i > 0xBF) sum += freqs1[(0xE0 << 8) | i];
        if (i < 0x80 || i > 0x9F) sum += freqs1[(0xED << 8) | i];
        if (i < 0x90 || i > 0xBF) sum += freqs1[(0xF0 << 8) | i];
        if (i < 0x80 || i > 0x8F) sum += freqs1[(0xF4 << 8) | i];
        if (i < 0x80 || i > 0xBF) {
            for (int j = 0xC2; j <= 0xDF; j++) sum += freqs1[(j << 8) | i];
            for (int j = 0xE1; j <= 0xEC; j++) sum += freqs1[(j << 8) | i];
            sum += (uint64_t)freqs1[(0xF1 << 8) | i] + freqs1[(0xF2 << 8) | i] + freqs1[(0xF3 << 8) | i];
            sum += freqs1[(0xEE << 8) | i];
            sum += freqs1[(0xEF << 8) | i];
        } else {
            sum2 += freqs0[i];
        }
        if (sum != 0) return MASK_NOT_TEXT;
    }
    return sum2 >= (uint64_t)(count / 8) ? (uint8_t)(MASK_NOT_TEXT | 8 /*DT_UTF8*/) : MASK_NOT_TEXT;
}

// computeTextStats :187-305 (strict = false) after the two histograms have been counted; has_magic: GetMagicType(block) != NO_MAGIC
KZ_TEXT_HD uint8_t text_stats_mode(const uint32_t* freqs0, const uint32_t* freqs1, int count, bool has_magic) {
    if (has_magic) return MASK_NOT_TEXT;
    int64_t nb_text = (int64_t)freqs0[CR] + freqs0[LF];
    int64_t nb_ascii = 0;
    for (int i = 0; i < 128; i++) {
        if (is_text((uint8_t)i)) nb_text += freqs0[i];
        nb_ascii += freqs0[i];
    }
    const int64_t nb_bin = (int64_t)count - nb_ascii;
    bool not_text;
    if (nb_bin > (count >> 2)) {
        not_text = true;
    } else {
        not_text = nb_text < count / 4;
        not_text = not_text || (int64_t)freqs0[32] < count / 50;
    }
    uint8_t res = 0;
    if (not_text) return (uint8_t)(res | detect_text_type(freqs0, freqs1, count));
    if (nb_bin <= count - count / 10) {
        const int64_t f1 = freqs0['<'], f2 = freqs0['>'];
        const int64_t f3 = (int64_t)freqs1[('&' << 8) | 'a'] + freqs1[('&' << 8) | 'g'] + freqs1[('&' << 8) | 'l'] + freqs1[('&' << 8) | 'q'];
        int64_t min_freq = ((int64_t)count - nb_bin) >> 9;
        if (min_freq < 2) min_freq = 2;
        if (f1 >= min_freq && f2 >= min_freq && f3 > 0) {
            if (f1 < f2) {
                if (f1 >= f2 - f2 / 100) res |= MASK_XML_HTML;
            } else if (f2 < f1) {
                if (f2 >= f1 - f1 / 100) res |= MASK_XML_HTML;
            } else {
                res |= MASK_XML_HTML;
            }
        }
    }
    if (freqs0[CR] != 0 && freqs0[CR] == freqs0[LF]) {
        bool is_crlf = true;
        for (int i = 0; i < 256; i++) {
            if (i != LF && freqs1[((int)CR << 8) | i] != 0) {
                is_crlf = false;
                break;
            }
            if (i != CR && freqs1[(i << 8) | LF] != 0) {
                is_crlf = false;
                break;
            }
        }
        if (is_crlf) res |= MASK_CRLF;
    }
    return res;
}

KZ_TEXT_HD const uint8_t* entry_bytes(const Dict& D, const Entry& e, const uint8_t* base) { return e.ptr >= 0 ? base + e.ptr : D.swords + (~e.ptr); }

KZ_TEXT_HD bool same_words(const uint8_t* a, const uint8_t* b, int n) {
    for (int i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

KZ_TEXT_HD bool expand_dictionary(Dict& D) {  // :1400-1413 (the entries are pre-initialised, only the logical size moves)
    if (D.dict_size >= MAX_DICT_SIZE) return false;
    D.dict_size <<= 1;
    return true;
}

// the "word not found: maybe remember it" step shared by Forward :1302-1324 and Inverse :1559-1581
KZ_TEXT_HD void add_word(Dict& D, int& words, int64_t off, int32_t length, uint32_t h1) {
    Entry& pe = D.list[words];
    if ((pe.data & MASK_LENGTH) >= D.static_size) {
        D.map[(uint32_t)pe.hash & D.hash_mask] = -1;
        pe.ptr = off;
        pe.hash = (int32_t)h1;
        pe.data = (length << 24) | words;
    }
    D.map[h1 & D.hash_mask] = words;
    words++;
    if (words >= D.dict_size) {
        if (!expand_dictionary(D)) words = D.static_size;
    }
}

// emitSymbols :1415-1487. Returns the number of bytes written, or dst_len + 1 on overflow.
KZ_TEXT_HD int emit_symbols(const uint8_t* src, int len, uint8_t* dst, int dst_len, bool is_crlf) {
    int d = 0;
    const bool checked = !(2 * (int64_t)len < (int64_t)dst_len);
    for (int i = 0; i < len; i++) {
        const uint8_t cur = src[i];
        if (cur == ESCAPE_TOKEN1) {
            if (checked && d + 1 >= dst_len) return dst_len + 1;
            dst[d++] = ESCAPE_TOKEN1;
            dst[d++] = ESCAPE_TOKEN1;
        } else if (cur == CR) {
            if (!is_crlf) {
                if (checked && d >= dst_len) return dst_len + 1;
                dst[d++] = cur;
            }
        } else {
            if (cur >= 0x80) {
                if (checked && d >= dst_len) return dst_len + 1;
                dst[d++] = ESCAPE_TOKEN1;
            }
            if (checked && d >= dst_len) return dst_len + 1;
            dst[d++] = cur;
        }
    }
    return d;
}

KZ_TEXT_HD int emit_word_index2(uint8_t* dst, int widx) {  // :1489-1511
    widx++;
    if (widx >= THRESHOLD3) {
        if (widx >= THRESHOLD4) {
            dst[0] = (uint8_t)(0xF0 | (widx >> 16));
            dst[1] = (uint8_t)(widx >> 8);
            dst[2] = (uint8_t)widx;
            return 3;
        }
        dst[0] = (uint8_t)(0xC0 | (widx >> 8));
        dst[1] = (uint8_t)widx;
        return 2;
    }
    dst[0] = (uint8_t)(0x80 | widx);
    return 1;
}

// textCodec2.Forward :1257-1398 after the statistics: the walk over the block. D.map / D.list are initialised (static entries mapped).
// Returns the encoded length, or -1 when the reference returns an error (the sequence then skips the transform).
KZ_TEXT_HD int forward_walk(const uint8_t* src, int count, uint8_t* dst, uint8_t mode, Dict& D) {
    const int src_end = count, dst_end = count, dst_end3 = dst_end - 3;  // MaxEncodedLen = count (:1720-1724)
    int emit_anchor = 0;
    int words = D.static_size;
    const bool is_crlf = (mode & MASK_CRLF) != 0;
    dst[0] = mode;
    int src_idx = 0, dst_idx = 1;
    while (src_idx < src_end && src[src_idx] == ' ') {
        dst[dst_idx++] = ' ';
        src_idx++;
        emit_anchor++;
    }
    if (src_idx >= src_end) return -1;  // all spaces: the reference would index past the slice; never text anyway
    int delim_anchor = src_idx;
    if (is_text(src[src_idx])) delim_anchor = src_idx - 1;
    while (src_idx < src_end) {
        if (is_text(src[src_idx])) {
            src_idx++;
            continue;
        }
        if (src_idx > delim_anchor + 2 && is_delimiter(src[src_idx])) {
            const int32_t length = src_idx - delim_anchor - 1;
            if (length <= MAX_WORD_LENGTH) {
                const uint8_t val = src[delim_anchor + 1];
                uint32_t h1 = hash_step(HASH1, val);
                uint32_t h2 = hash_step(HASH1, (uint32_t)val ^ 0x20u);
                for (int i = delim_anchor + 2; i < src_idx; i++) {
                    const uint32_t h = (uint32_t)src[i] * HASH2;
                    h1 = (h1 * HASH1) ^ h;
                    h2 = (h2 * HASH1) ^ h;
                }
                int pe = -1;
                const int pe1 = D.map[h1 & D.hash_mask];
                if (pe1 >= 0 && D.list[pe1].hash == (int32_t)h1 && (D.list[pe1].data >> 24) == length) {
                    pe = pe1;
                } else {
                    const int pe2 = D.map[h2 & D.hash_mask];
                    if (pe2 >= 0 && D.list[pe2].hash == (int32_t)h2 && (D.list[pe2].data >> 24) == length) pe = pe2;
                }
                if (pe >= 0 && !same_words(entry_bytes(D, D.list[pe], src) + 1, src + delim_anchor + 2, length - 1)) pe = -1;
                if (pe < 0) {
                    if ((length > 3 || (length == 3 && words < THRESHOLD2)) && pe1 < 0) add_word(D, words, delim_anchor + 1, length, h1);
                } else {
                    if (emit_anchor != delim_anchor || src[delim_anchor] != ' ')
                        dst_idx += emit_symbols(src + emit_anchor, delim_anchor + 1 - emit_anchor, dst + dst_idx, dst_end - dst_idx, is_crlf);
                    if (dst_idx >= dst_end3) return -1;
                    if (pe != pe1) dst[dst_idx++] = MASK_FLIP_CASE;
                    const Entry& e = D.list[pe];
                    dst_idx += emit_word_index2(dst + dst_idx, (int)(e.data & MASK_LENGTH));
                    emit_anchor = delim_anchor + 1 + (int)(e.data >> 24);
                }
            }
        }
        delim_anchor = src_idx;
        src_idx++;
    }
    dst_idx += emit_symbols(src + emit_anchor, src_end - emit_anchor, dst + dst_idx, dst_end - dst_idx, is_crlf);
    if (dst_idx > dst_end) return -1;
    return dst_idx;
}

// textCodec2.Inverse :1513-1718 (bitstream version 6). cap = len(dst). D is initialised; the list size follows reset(len(dst)).
// Returns the decoded length, or -1 on invalid data.
KZ_TEXT_HD int64_t inverse_walk(const uint8_t* src, int len, uint8_t* dst, int64_t cap, Dict& D) {
    int words = D.static_size;
    bool word_run = false;
    const bool is_crlf = (src[0] & MASK_CRLF) != 0;
    int src_idx = 1;
    int64_t dst_idx = 0;
    const int src_end = len;
    const int64_t dst_end = cap;
    int delim_anchor = src_idx;
    if (is_text(src[src_idx])) delim_anchor = src_idx - 1;
    while (src_idx < src_end && dst_idx < dst_end) {
        uint8_t cur = src[src_idx];
        if (is_text(cur)) {
            dst[dst_idx++] = cur;
            src_idx++;
            continue;
        }
        if (src_idx > delim_anchor + 3 && is_delimiter(cur)) {
            const int32_t length = src_idx - delim_anchor - 1;
            if (length <= MAX_WORD_LENGTH) {
                uint32_t h1 = HASH1;
                for (int i = delim_anchor + 1; i < src_idx; i++) h1 = hash_step(h1, src[i]);
                int pe = -1;
                const int pe1 = D.map[h1 & D.hash_mask];
                if (pe1 >= 0) {
                    const Entry& e = D.list[pe1];
                    if (e.hash == (int32_t)h1 && (e.data >> 24) == length && same_words(entry_bytes(D, e, src) + 1, src + delim_anchor + 2, length - 1)) pe = pe1;
                }
                if (pe < 0) {
                    if ((length > 3 || words < THRESHOLD2) && pe1 < 0) add_word(D, words, delim_anchor + 1, length, h1);
                }
            }
        }
        src_idx++;
        uint8_t flip_mask = 0;
        if (cur >= 128) {
            if (cur == MASK_FLIP_CASE) {
                flip_mask = 0x20;
                if (src_idx >= src_end) return -1;
                cur = src[src_idx++];
            }
            int idx = cur & 0x7F;
            if (idx >= 64) {
                if (idx >= 112) {
                    if (src_idx + 1 >= src_end) return -1;
                    idx = ((idx & 0x0F) << 16) | ((int)src[src_idx] << 8) | src[src_idx + 1];
                    src_idx += 2;
                } else {
                    if (src_idx >= src_end) return -1;
                    idx = ((idx & 0x1F) << 8) | src[src_idx];
                    src_idx++;
                }
                if (idx > D.dict_size) return -1;
            } else if (idx == 0) {
                return -1;
            }
            idx--;
            if (idx < 0 || idx >= D.dict_size) return -1;  // the reference indexes the list slice: beyond its length it panics
            const Entry& e = D.list[idx];
            const int length = (e.data >> 24) & 0xFF;
            if (length > 1) {
                if (word_run) dst[dst_idx++] = ' ';
                word_run = true;
                delim_anchor = src_idx;
            } else {
                word_run = false;
                delim_anchor = src_idx - 1;
            }
            if (e.ptr == NIL || dst_idx + length >= dst_end) return -1;
            const uint8_t* w = entry_bytes(D, e, src);
            for (int i = 0; i < length; i++) dst[dst_idx + i] = w[i];
            dst[dst_idx] ^= flip_mask;
            dst_idx += length;
        } else {
            if (cur == ESCAPE_TOKEN1) {
                if (src_idx >= src_end) return -1;
                dst[dst_idx++] = src[src_idx++];
            } else {
                if (is_crlf && cur == LF) {
                    dst[dst_idx++] = CR;
                    if (dst_idx >= dst_end) return -1;
                }
                dst[dst_idx++] = cur;
            }
            word_run = false;
            delim_anchor = src_idx - 1;
        }
    }
    if (src_idx != src_end) return -1;
    return dst_idx;
}

}  // namespace textc
}  // namespace kz
