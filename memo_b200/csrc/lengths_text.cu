// MONI *.lengths / *.lengths.vert text -> int32 lengths, on the HOST (no device work, no CUDA call).
//
// The reference flattens every per-genome MONI output with `grep -v '^>' | tr ' ' '\n' | grep .`
// (src/index.sh:79), pastes the columns into dap.txt (index.sh:83) and parses that text again with
// `map(int, row.split(' '))` (src/dap_to_bed.py:87).  The `--lengths` extension of
// memo_b200.dap_to_bed reads the per-genome files side by side instead; this is its tokenizer:
// one call parses a run of bytes of ONE file into one strided column of the [rows, C] DAP block
// that goes to the device.  memo_lengths_block_parse walks a group of neighbouring columns through
// one block, tile of rows by tile of rows, refilling each file's buffer with read(2) as it goes:
// one call per block and host thread (ctypes releases the GIL).  Plain byte loops, no shared state.
#include <errno.h>
#include <string.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "common.cuh"

namespace memo {
namespace {

enum : int32_t { AT_LINE_START = 0, IN_LINE = 1, IN_HEADER = 2 };

inline bool is_space(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }   // bytes `\s`

#if defined(__SSE2__)
#define MEMO_LENGTHS_SIMD 1
// 64 bytes -> bit masks of the digits and of the white space among them
inline void classify64(const uint8_t* p, uint64_t& digits, uint64_t& spaces) {
    const __m128i zero = _mm_set1_epi8('0'), nine = _mm_set1_epi8(9), tab = _mm_set1_epi8('\t'),
                  four = _mm_set1_epi8(4), blank = _mm_set1_epi8(' ');
    digits = spaces = 0;
    for (int k = 0; k < 4; ++k) {
        const __m128i v = _mm_loadu_si128((const __m128i*)(p + 16 * k));
        const __m128i d = _mm_sub_epi8(v, zero), t = _mm_sub_epi8(v, tab);
        const __m128i is_d = _mm_cmpeq_epi8(_mm_min_epu8(d, nine), d);
        const __m128i is_s = _mm_or_si128(_mm_cmpeq_epi8(_mm_min_epu8(t, four), t), _mm_cmpeq_epi8(v, blank));
        digits |= uint64_t(uint32_t(_mm_movemask_epi8(is_d))) << (16 * k);
        spaces |= uint64_t(uint32_t(_mm_movemask_epi8(is_s))) << (16 * k);
    }
}
#endif

}  // namespace
}  // namespace memo

extern "C" int memo_lengths_text_parse(const uint8_t* text, int64_t n_bytes, int32_t final_block,
                                       int32_t* state, int32_t* out, int64_t out_stride,
                                       int64_t max_vals, int64_t* result) {
    using namespace memo;
    if (!state || !result || n_bytes < 0 || max_vals < 0 || (n_bytes > 0 && !text) ||
        (max_vals > 0 && !out) || *state < AT_LINE_START || *state > IN_HEADER) {
        set_error("memo_lengths_text_parse: bad argument");
        return MEMO_ERR_ARG;
    }
    int32_t st = *state;
    int64_t i = 0, count = 0, errors = 0;
    const int64_t n = n_bytes;
    while (i < n) {
#ifdef MEMO_LENGTHS_SIMD
        // 64 bytes of nothing but numbers and white space: the numbers that end inside them are
        // found from the bit masks (two independent chains of blsr / tzcnt) and converted eight
        // bytes at a time; anything else -- headers, long numbers, the last bytes -- takes the byte
        // loop below
        while (st != IN_HEADER && i + 72 <= n && count < max_vals) {
            uint64_t dm, sm;
            classify64(text + i, dm, sm);
            if (~(dm | sm)) break;                           // '>' or a byte int() would reject
            uint64_t starts = dm & ~(dm << 1), ends = dm & ~(dm >> 1) & 0x7FFFFFFFFFFFFFFFull;
            if (dm & 1) {                                    // the chunk begins inside digits: only if they begin here
                if (i > 0 ? uint8_t(text[i - 1] - '0') < 10 : false) break;
            }
            int64_t todo = __builtin_popcountll(ends);       // numbers followed by a byte of the chunk
            if (todo == 0) {
                if (dm >> 63) break;                         // one number across the whole chunk
                st = text[i + 63] == '\n' ? AT_LINE_START : IN_LINE;
                i += 64;
                continue;
            }
            if (todo > max_vals - count) todo = max_vals - count;
            int64_t last_end = 0;
            bool is_long = false;
            for (int64_t t = 0; t < todo; ++t) {
                const int a = __builtin_ctzll(starts), b = __builtin_ctzll(ends);
                starts &= starts - 1;
                ends &= ends - 1;
                const int len = b - a + 1;
                if (len > 8) { is_long = true; last_end = a; break; }   // nine and more digits: byte loop
                uint64_t w;
                memcpy(&w, text + i + a, 8);
                w ^= 0x3030303030303030ull;
                w <<= 64 - 8 * len;
                w = (w * 2561) >> 8;
                w = ((w & 0x00FF00FF00FF00FFull) * 6553601) >> 16;
                out[count * out_stride] = int32_t(((w & 0x0000FFFF0000FFFFull) * 42949672960001ull) >> 32);
                ++count;
                last_end = b + 1;
            }
            if (is_long) {                                   // i at the long number's first digit
                i += last_end;
                st = IN_LINE;
                break;
            }
            if (ends == 0 && count < max_vals) {             // every number of the chunk that ends in it is done
                const int64_t step = (dm >> 63) ? __builtin_ctzll(starts) : 64;   // an open number: restart at it
                st = text[i + step - 1] == '\n' ? AT_LINE_START : IN_LINE;
                i += step;
            } else {                                         // stopped at max_vals: just behind the last number
                i += last_end;
                st = IN_LINE;
            }
        }
        if (i >= n) break;
#endif
        const uint8_t c = text[i];
        if (st != IN_HEADER && uint8_t(c - '0') < 10) {      // a number starts here
            if (count == max_vals) break;                    // the column has what was asked for
            int64_t j;
            uint64_t v;
            uint64_t w = 0, nd = 0;
            if (i + 8 <= n) {                                // eight bytes at once: where does it end?
                memcpy(&w, text + i, 8);
                w ^= 0x3030303030303030ull;                  // digits -> 0 .. 9
                nd = (((w & 0x7F7F7F7F7F7F7F7Full) + 0x7676767676767676ull) | w) & 0x8080808080808080ull;
            }
            if (nd) {                                        // 1 .. 7 digits and the byte after them
                const int len = __builtin_ctzll(nd) >> 3;
                const uint8_t sep = uint8_t(w >> (8 * len)) ^ 0x30;
                j = i + len;
                if (sep == ' ') ++j;                         // the usual separator goes with its number
                w <<= 64 - 8 * len;                          // leading zero digits, last digit in the top byte
                w = (w * 2561) >> 8;                         // pairs
                w = ((w & 0x00FF00FF00FF00FFull) * 6553601) >> 16;
                v = ((w & 0x0000FFFF0000FFFFull) * 42949672960001ull) >> 32;
            } else {                                         // long numbers and the last bytes of a block
                j = i;
                v = 0;
                while (j < n && uint8_t(text[j] - '0') < 10) {
                    v = v * 10 + (text[j] - '0');
                    if (v > 0xFFFFFFFFFFull) v = 0xFFFFFFFFFFull;    // saturate: flagged below
                    ++j;
                }
                if (j == n && !final_block) break;           // the number may go on in the next block
                if (v > 0x7FFFFFFFull) { errors |= 8; break; }
            }
            out[count * out_stride] = int32_t(v);
            ++count;
            st = IN_LINE;
            i = j;
            continue;
        }
        if (st == IN_HEADER) {                               // a '>' line: dropped up to its newline
            const void* nl = memchr(text + i, '\n', size_t(n - i));
            if (!nl) { i = n; break; }
            i = (const uint8_t*)nl - text + 1;
            st = AT_LINE_START;
            continue;
        }
        if (c == '\n') { st = AT_LINE_START; ++i; continue; }
        if (is_space(c)) { st = IN_LINE; ++i; continue; }
        if (c == '>' && st == AT_LINE_START) { st = IN_HEADER; ++i; continue; }
        errors |= 1;                                         // int() raises in the reference
        break;
    }
    *state = st;
    result[0] = count;
    result[1] = errors;
    result[2] = i;
    return MEMO_OK;
}

extern "C" int memo_lengths_block_parse(memo_lengths_file_t* files, int32_t n_files, int32_t* out,
                                        int64_t out_stride, int64_t rows, int64_t tile_rows) {
    using namespace memo;
    if (!files || n_files < 0 || rows < 0 || tile_rows <= 0 || (rows > 0 && n_files > 0 && !out)) {
        set_error("memo_lengths_block_parse: bad argument");
        return MEMO_ERR_ARG;
    }
    for (int32_t j = 0; j < n_files; ++j)
        if (!files[j].buf || files[j].cap <= 0 || files[j].lo < 0 || files[j].lo > files[j].hi ||
            files[j].hi > files[j].cap) {
            set_error("memo_lengths_block_parse: bad buffer of file %d", j);
            return MEMO_ERR_ARG;
        }
    int bad = 0;
    for (int64_t r0 = 0; r0 < rows; r0 += tile_rows) {
        const int64_t want = rows - r0 < tile_rows ? rows - r0 : tile_rows;
        for (int32_t j = 0; j < n_files; ++j) {
            memo_lengths_file_t& f = files[j];
            if (f.ended || f.error) continue;
            int32_t* dst = out + r0 * out_stride + j;
            int64_t got = 0;
            while (got < want) {
                int64_t res[3];
                memo_lengths_text_parse(f.buf + f.lo, f.hi - f.lo, f.eof, &f.state, dst + got * out_stride,
                                        out_stride, want - got, res);
                got += res[0];
                f.lo += res[2];
                if (res[1]) { f.error = res[1]; break; }
                if (got == want) break;
                if (f.eof) { f.ended = 1; break; }           // the file has nothing left
                const int64_t rest = f.hi - f.lo;             // out of bytes: keep the cut number, read on
                if (rest == f.cap) { f.error = 1; break; }   // a "number" as long as the buffer
                if (rest && f.lo) memmove(f.buf, f.buf + f.lo, size_t(rest));
                f.lo = 0;
                f.hi = rest;
                ssize_t n;
                do n = read(f.fd, f.buf + rest, size_t(f.cap - rest)); while (n < 0 && errno == EINTR);
                if (n < 0) { f.error = 32; break; }
                if (n == 0) f.eof = 1;
                f.hi = rest + n;
            }
            f.count += got;
            if (f.error) bad = 1;
        }
    }
    if (bad) {
        set_error("memo_lengths_block_parse: a file could not be parsed (see the error bits)");
        return MEMO_ERR_ARG;
    }
    return MEMO_OK;
}
