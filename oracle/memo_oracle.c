/* C restatement of the reference's hot path -- TEST INFRASTRUCTURE, not product.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  Parity status: PINNED (tests/test_oracle_c.py
 * checks it against the numpy oracle, which is pinned to outputs of the
 * reference scripts in tests/golden/).
 *
 * Reference lines followed (paths relative to the reference root):
 *   mo_index_build   src/dap_to_bed.py:85-134  row-at-a-time: sort the row
 *                    (:89-90), flag prev <= curr (:123), overlap with the stored
 *                    MEM of the column (:93-109), chr-end rows (:126-128,:133-134)
 *   mo_query         src/memo_query.py:42-71   shadow cast + clip (:46-49),
 *                    paint a [W, N(+1)] byte matrix (:57-63), argmax (:70)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t row_begin, n_rows;
    int32_t pos0, rec_len, rec_id, flags;   /* same layout as memo_segment_t */
} mo_segment_t;

static void sort_desc(int64_t* a, int n) {
    /* insertion sort: rows are short (n = genomes - 1) */
    for (int i = 1; i < n; ++i) {
        int64_t v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] < v) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
    }
}

typedef struct { int64_t *start, *end, *col, *rec; int64_t n, cap; } mo_out_t;

static inline void put(mo_out_t* o, int64_t* ps, int64_t* pe, char* has, int j,
                       int64_t start, int64_t end, int64_t rec) {
    if (has[j]) {                                   /* dap_to_bed.py:104-106 */
        int64_t s = ps[j] > start ? ps[j] : start;
        int64_t e = pe[j] < end ? pe[j] : end;
        if (e >= s) {
            if (o->n < o->cap) {
                o->start[o->n] = s; o->end[o->n] = e; o->col[o->n] = j + 1; o->rec[o->n] = rec;
            }
            o->n++;
        }
    }
    ps[j] = start; pe[j] = end; has[j] = 1;         /* :107 */
}

/* Returns the number of index rows the input produces (rows past `cap` are
 * counted, not stored).  A run without flag bit 0 (primed) continues a record:
 * row row_begin-1 is its halo and the stored MEM of each column is taken to be
 * (pos0-1, pos0-1+S[halo]) -- exact for valid matching statistics (used only by
 * the multi-threaded baseline; the single-run form is exact for any input). */
int64_t mo_index_build(const int32_t* dap, int64_t rows, int32_t C, int32_t ld,
                       const mo_segment_t* segs, int32_t n_seg, int32_t order,
                       int64_t* out_rec, int64_t* out_start, int64_t* out_end, int64_t* out_col,
                       int64_t cap) {
    (void)rows;
    mo_out_t o = {out_start, out_end, out_col, out_rec, 0, cap};
    int64_t* prev = malloc(sizeof(int64_t) * C);
    int64_t* cur = malloc(sizeof(int64_t) * C);
    int64_t* ps = malloc(sizeof(int64_t) * C);
    int64_t* pe = malloc(sizeof(int64_t) * C);
    char* has = malloc(C);
    for (int s = 0; s < n_seg; ++s) {
        const mo_segment_t* g = &segs[s];
        memset(has, 0, C);                              /* :129 */
        int64_t r0 = 0;
        if (!(g->flags & 1)) {
            const int32_t* h = dap + (g->row_begin - 1) * (int64_t)ld;
            for (int j = 0; j < C; ++j) prev[j] = h[j];
            if (order) sort_desc(prev, C);
            for (int j = 0; j < C; ++j) { ps[j] = g->pos0 - 1; pe[j] = g->pos0 - 1 + prev[j]; has[j] = 1; }
        } else {
            const int32_t* h = dap + g->row_begin * (int64_t)ld;
            for (int j = 0; j < C; ++j) prev[j] = h[j];
            if (order) sort_desc(prev, C);
            for (int j = 0; j < C; ++j) put(&o, ps, pe, has, j, g->pos0, g->pos0 + prev[j], g->rec_id); /* :130 */
            r0 = 1;
        }
        for (int64_t r = r0; r < g->n_rows; ++r) {
            const int32_t* row = dap + (g->row_begin + r) * (int64_t)ld;
            const int64_t p = g->pos0 + r;
            for (int j = 0; j < C; ++j) cur[j] = row[j];
            if (order) sort_desc(cur, C);               /* :89-90 */
            for (int j = 0; j < C; ++j)
                if (prev[j] <= cur[j]) put(&o, ps, pe, has, j, p, p + cur[j], g->rec_id);  /* :123-124 */
            int64_t* t = prev; prev = cur; cur = t;
        }
        if (g->flags & 2) {                             /* :126-128, :133-134 */
            const int64_t n = g->rec_len;
            for (int j = 0; j < C; ++j) put(&o, ps, pe, has, j, n, 2 * n, g->rec_id);
        }
    }
    free(prev); free(cur); free(ps); free(pe); free(has);
    return o.n;
}

/* Query of one record's rows over [q_start, q_end).  out: conservation int64[W],
 * membership uint8[W * n_docs].  Returns 0, or -1 when a row's order is out of
 * range for n_docs (the reference would write out of bounds). */
int mo_query(const int64_t* f1, const int64_t* f2, const int64_t* f3, int64_t n_rows,
             int64_t q_start, int64_t q_end, int32_t k, int32_t n_docs, int32_t membership,
             void* out) {
    const int64_t W = q_end - q_start;
    const int width = membership ? n_docs : n_docs + 1;
    unsigned char* rec = malloc((size_t)(W > 0 ? W : 1) * width);
    if (membership) memset(rec, 1, (size_t)W * width);                   /* :51 */
    else {
        memset(rec, 0, (size_t)W * width);                               /* :53-54 */
        for (int64_t p = 0; p < W; ++p) rec[p * width + n_docs] = 1;
    }
    const unsigned char bit = membership ? 0 : 1;
    for (int64_t i = 0; i < n_rows; ++i) {
        if (!(f1[i] > q_start && f1[i] < q_end + k)) continue;           /* :25-27 */
        int64_t start = f1[i] - q_start;                                 /* :46 */
        int64_t cend = f2[i] - q_start - (k - 1);                        /* :47 */
        if (start < 0) start = 0; if (start > W) start = W;              /* :48 */
        if (cend < 0) cend = 0; if (cend > W) cend = W;
        if (!(cend < start)) continue;                                   /* :49 */
        if (f3[i] < 0 || f3[i] >= width) { free(rec); return -1; }
        for (int64_t p = cend; p < start; ++p) rec[p * width + f3[i]] = bit;   /* :61-62 */
    }
    if (membership) memcpy(out, rec, (size_t)W * width);                 /* :68 */
    else {
        int64_t* o = out;
        for (int64_t p = 0; p < W; ++p) {                                /* :70 argmax */
            int j = 0;
            while (!rec[p * width + j]) ++j;
            o[p] = j;
        }
    }
    free(rec);
    return 0;
}

/* Synthetic HPRC-shaped DAP (measurement input; same integers as
 * oracle/memo_oracle.py:synth_dap, which tests/test_oracle_c.py checks): rows
 * [row0, row0 + rows) of one record of length rec_len into dap[rows, C].
 *   h = splitmix64(seed ^ p*K1 ^ c*K2); short draw 12 + ctz(h | 2^20); with
 *   probability 2^-9 a long draw ((ctz((h>>32) | 2^16) + 1) << 10) + ((h>>48) & 1023);
 *   MS[p][c] = min(max_{q<=p}(d[q][c] + q) - p, rec_len - p). */
static inline uint64_t mo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void mo_synth_dap(int32_t* dap, int64_t row0, int64_t rows, int32_t C, int64_t rec_len,
                  uint64_t seed, int32_t dense) {
    const int64_t LOOKBACK = 32768;                 /* draws are < 2^15 */
    const int64_t lo = row0 > LOOKBACK ? row0 - LOOKBACK : 0;
    int64_t* reach = malloc(sizeof(int64_t) * C);
    for (int c = 0; c < C; ++c) reach[c] = 0;
    for (int64_t p = lo; p < row0 + rows; ++p) {
        const uint64_t pp = (uint64_t)p * 0x9E3779B97F4A7C15ull;
        int32_t* out = p >= row0 ? dap + (p - row0) * (int64_t)C : NULL;
        for (int c = 0; c < C; ++c) {
            const uint64_t h = mo_splitmix64(seed ^ pp ^ ((uint64_t)c * 0xC2B2AE3D27D4EB4Full));
            int64_t d = 12 + __builtin_ctzll(h | (1ull << 20));
            if (dense) {
                if (((h >> 20) & 1023) == 0) {
                    const int64_t dl = (((int64_t)__builtin_ctzll((h >> 32) | (1ull << 16)) + 1) << 8) +
                                       (int64_t)((h >> 48) & 255);
                    if (dl > d) d = dl;
                }
            } else if (((h >> 20) & 511) == 0) {
                const int64_t dl = (((int64_t)__builtin_ctzll((h >> 32) | (1ull << 16)) + 1) << 10) +
                                   (int64_t)((h >> 48) & 1023);
                if (dl > d) d = dl;
            }
            if (d + p > reach[c]) reach[c] = d + p;
            if (out) {
                int64_t ms = reach[c] - p;
                if (ms > rec_len - p) ms = rec_len - p;
                out[c] = (int32_t)ms;
            }
        }
    }
    free(reach);
}
