// DAP -> MEMO index rows on sm_100a: single-pass, single-KERNEL build for wide rows
// (any n_cols <= 512; the 94-genome configurations).  One warp per strip of rows,
// rows staged by bulk async copies (TMA engine), index rows written straight to
// their place in the ordered output (decoupled look-back over strips).
//
// Replaces the hot loop of the reference's src/dap_to_bed.py (--mem --overlap
// [--order]): get_new_record :85-91, dap_to_mem :116-134, print_interval /
// overlaps :93-109.  Mathematics as in index_build.cu (DESIGN.md "index build"):
// with E[r][c] = p(r) + v[r][c] and A[r] = E[r] sorted descending (--order) or
// E[r] itself, row r emits (p, A[r-1][j], j+1) for every j with A[r][j] >
// A[r-1][j] and A[r-1][j] >= p, provided no E decreases down a column (matching
// statistics; checked on the way, result[MEMO_RES_IRREGULAR]).
//
// A strip of R consecutive rows streams through the warp's shared memory in
// chunks of T rows (one 1-D bulk copy per chunk from the 16-byte aligned address
// below its first row; the row before the chunk sits right in front of it, so
// that "the cell one row up" is always ld words back).
//   phase A  the chunk is scanned FLAT: lane l takes the aligned quads l, l + 32,
//            ... of the chunk (one 128-bit LDS) and compares them with the four
//            words ld back.  A row of 4 n_cols bytes is not 16-byte aligned unless
//            n_cols % 4 == 0 (93 columns: 372 B) -- TMA tensor copies cannot realign
//            it (the inner box coordinate must itself be 16-byte aligned) -- so the
//            words ld back come from two more aligned 128-bit loads and a
//            compile-time word shift (ld % 4, one code copy per shift).  A quad with
//            v[r][c] != v[r-1][c] - 1 flags the one or two rows it belongs to:
//            ~10 warp instructions per row instead of a warp-wide compare per row;
//   phase B  only flagged rows (30 % on HPRC-shaped data) enter the warp-wide
//            path: the sorted MEM ends A of the previous row live in registers
//            (lane l: positions l KPL ..), a cell whose MEM end moved x -> y is
//            a delete/insert (positions holding x <= A <= y shift by one), the
//            positions where A changed are the row's index rows.  A is sorted from
//            scratch once per strip (warp bitonic merge-split network).
// Index rows are staged in the warp's shared memory (8-byte records {end, row | order};
// a dense strip spills to a per-warp global overflow area) and written straight to
// their place in the ordered output: no scratch round trip, no scan or gather kernel.
// The place of a strip = rows of all earlier strips, resolved on two levels (a chain
// over 2 M strips, 32 per L2 round trip, would be slower than the build itself):
//   * a strip publishes its row count when its rows are through (status word) and counts
//     itself into its group of 32 consecutive strips; the group's last finisher sums the
//     group, looks back over the earlier GROUPS (aggregate / inclusive prefix per group,
//     decoupled look-back) and publishes the group's inclusive prefix;
//   * a strip's offset = inclusive prefix of the previous group + counts of the earlier
//     strips of its own group.
// The staging area is double buffered: a strip is resolved and copied out (coalesced
// stores) after the warp's NEXT strip is through, when its predecessors have normally
// long finished.  No warp ever waits: a strip whose predecessors are still running by
// then (behind a dense stretch of the DAP) is PARKED -- its staged records go to a
// global parking area -- and wide2_finish_kernel, which also writes the per-run row
// totals, copies the parked strips to their places afterwards.
#include <mutex>
#include <vector>

#include "index_fast.cuh"
#include "warp_sort.cuh"

namespace memo {
namespace {

constexpr int W2_WARPS = 4;            // warps per CTA (independent streams)
constexpr int W2_INLINE_SEGS = 8;      // record runs passed in the kernel parameters
constexpr int W2_MAX_T = 28;           // rows per chunk (row flags of a chunk + the rows around it: 32 bits)
constexpr int W2_MAX_WORDS = 3968;     // words per chunk (phase A: one flag bit per 128 words)
constexpr int W2_CTAS = 5;             // CTAs per SM the shared-memory budget is cut for
constexpr int W2_FIFO = 16;            // finished strips a warp can keep waiting in its ring

struct Wide2Params {
    const int32_t* dap;
    long long total_bytes;             // rows * ld * 4
    int32_t C, ld;
    int32_t T;                         // rows per chunk
    int32_t R;                         // compare rows per strip (the first chunk holds the predecessor row + T - 1)
    uint32_t magic;                    // ceil(2^32 / ld): word index -> row by one multiply
    int32_t n_seg;
    const memo_segment_t* segs;        // device tables (n_seg > W2_INLINE_SEGS)
    const long long* seg_unit_start;
    memo_segment_t isegs[W2_INLINE_SEGS];
    long long iunit[W2_INLINE_SEGS + 1];
    long long n_units;
    uint32_t warp_smem, off_stage, off_stg, off_bar, off_fifo;   // bytes inside the warp's region
    uint32_t cap;                      // staged index rows per warp (shared memory)
    uint2* ring;                       // per-warp ring of finished strips' records (global, L2 resident)
    uint32_t ring_cap;                 // records per warp, a power of two
    unsigned long long* status;        // [n_units] (rows of the strip << 1) | 1 once published
    unsigned int* gdone;               // [n_groups] published strips of the group
    unsigned long long* gstatus;       // [n_groups] (value << 2) | {0 none, 1 group total, 2 inclusive prefix}
    unsigned long long* strip_counter;
    unsigned long long* park_cursor;   // next free record of the parking area
    unsigned long long* park_off;      // [n_units] 1 + first parked record of the strip, 0 = written in place
    uint2* park;                       // parking area
    long long park_cap;
    uint2* ovf;                        // per-warp overflow of the staging area
    long long ovf_cap;
    int32_t* out_start;
    uint32_t* out_end;
    int32_t* out_order;
    long long out_cap;
    int64_t* seg_out_end;
    int64_t* result;
};

// Phase A over one chunk.  `st` = the stage (16-byte aligned; the chunk's first row starts
// `off` words in, the row before it ld words further back), n = rows of the chunk.  Returns
// the flags of the rows that hold a cell with v != (cell one row up) - 1: bit j + 1 = row j
// of the chunk (bit 0: the row before the chunk; bits above n: whatever follows the chunk).
// SHIFT = ld % 4: the four words ld back start SHIFT words before an aligned quad boundary.
template <int SHIFT>
__device__ __forceinline__ uint32_t scan_chunk(const uint32_t* st, int off, int n, int ld, uint32_t magic, int lane) {
    const int n_quads = (off + n * ld + 3) >> 2;
    const int back = (ld + 3) >> 2;                              // quads back to the first needed word's quad
    // pass 1, branch free: bit (32 - iterations + i) of `hit` = the lane's quad of iteration i
    // holds a changed cell
    uint32_t hit = 0;
    int iters = 0;
#pragma unroll 4
    for (int q = lane; q < n_quads; q += 32, ++iters) {
        const uint4 c = *reinterpret_cast<const uint4*>(st + 4 * q);
        uint32_t p0, p1, p2, p3;
        if (SHIFT == 0) {
            const uint4 a = *reinterpret_cast<const uint4*>(st + 4 * (q - back));
            p0 = a.x; p1 = a.y; p2 = a.z; p3 = a.w;
        } else {
            const uint4 a = *reinterpret_cast<const uint4*>(st + 4 * (q - back));
            const uint4 b = *reinterpret_cast<const uint4*>(st + 4 * (q - back + 1));
            if (SHIFT == 1) { p0 = a.w; p1 = b.x; p2 = b.y; p3 = b.z; }
            else if (SHIFT == 2) { p0 = a.z; p1 = a.w; p2 = b.x; p3 = b.y; }
            else { p0 = a.y; p1 = a.z; p2 = a.w; p3 = b.x; }
        }
        const uint32_t d = (c.x + 1u - p0) | (c.y + 1u - p1) | (c.z + 1u - p2) | (c.w + 1u - p3);
        hit = __funnelshift_r(hit, min(d, 1u), 1);
    }
    // pass 2 (few lanes, few quads): the rows of the flagged quads.  Words 4q .. 4q+3 of the
    // stage = words u .. u+3 counted from the start of the row before the chunk; they belong
    // to row u / ld and (a quad across a row end) (u+3) / ld
    uint32_t bits = 0;
    const int first_bit = 32 - iters;
    while (hit) {
        const int b = __ffs(hit) - 1;
        hit &= hit - 1;
        const uint32_t u = (uint32_t)(4 * (lane + 32 * (b - first_bit)) - off + ld);
        if (ld >= 3) {
            bits |= (1u << __umulhi(u, magic)) | (1u << __umulhi(u + 3u, magic));
        } else {
            // one- and two-column rows: a quad spans up to four rows (and 2^32 / 1 does not
            // fit the multiplier)
#pragma unroll
            for (uint32_t i = 0; i < 4u; ++i) bits |= 1u << ((u + i) >> (ld - 1));
        }
    }
    return __reduce_or_sync(FULL, bits);
}

__device__ __forceinline__ long long w2_unit_start(const Wide2Params& P, int i) {
    return P.n_seg <= W2_INLINE_SEGS ? P.iunit[i] : P.seg_unit_start[i];
}
// record run of strip s (binary search over the runs' first strips)
__device__ __forceinline__ int w2_find_run(const Wide2Params& P, long long s) {
    int a = 0, b = P.n_seg - 1;
    while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (w2_unit_start(P, mid) <= s) a = mid; else b = mid - 1;
    }
    return a;
}

// Rows of all strips before strip s: inclusive prefix of the previous group + the earlier
// strips of the strip's own group.  Never waits: returns false when a strip or group it needs
// is not published yet.  The group prefix is materialised lazily: whoever needs it walks back
// over the group totals to the nearest inclusive prefix (32 groups per step) and leaves the
// result behind for the others.
__device__ __forceinline__ bool w2_resolve(const Wide2Params& P, long long s, int lane, unsigned long long* excl) {
    volatile unsigned long long* const st = P.status;
    volatile unsigned long long* const gs = P.gstatus;
    const long long g = s >> 5;
    const int k = (int)(s & 31);
    unsigned long long v = 0;
    bool ok = true;
    if (lane < k) {
        const unsigned long long w = st[(g << 5) + lane];
        ok = (w & 1ull) != 0ull;
        v = w >> 1;
    }
    if (!__all_sync(FULL, ok)) return false;
    if (g > 0) {
        long long idx = g - 1;
        bool first = true;
        unsigned long long base = 0;
        for (;;) {
            const long long j = idx - lane;
            const unsigned long long w = j >= 0 ? gs[j] : 2ull;          // before the first group: inclusive 0
            const unsigned inc = __ballot_sync(FULL, (w & 3ull) == 2ull);
            const unsigned none = __ballot_sync(FULL, (w & 3ull) == 0ull);
            const int f = inc ? __ffs(inc) - 1 : 31;                     // walk ends at the nearest inclusive prefix
            if (none & (0xFFFFFFFFu >> (31 - f))) return false;          // a group on the way is not complete
            unsigned long long t = lane <= f ? (w >> 2) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
            base += t;
            if (inc) {
                if (!(first && f == 0) && lane == 0) gs[g - 1] = (base << 2) | 2ull;
                break;
            }
            first = false;
            idx -= 32;
        }
        if (lane == 0) v += base;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    *excl = v;
    return true;
}

// n staged records -> out_* at `excl`; p0 = position of the strip's first compare row, plen =
// its record's length (chr-end rows); records [0, ns) from sb, the rest from ob
__device__ __forceinline__ void w2_copy_out(const Wide2Params& P, const uint2* sb, const uint2* ob, uint32_t ns,
                                            unsigned long long excl, uint32_t total, uint32_t p0, uint32_t plen,
                                            int lane) {
    const unsigned long long ocap = (unsigned long long)P.out_cap;
    for (uint32_t i = lane; i < total; i += 32) {
        const uint2 e = i < ns ? sb[i] : ob[i - ns];
        const unsigned long long gi = excl + i;
        if (gi < ocap) {
            const uint32_t rc = e.y >> 16;
            P.out_start[gi] = (int32_t)(rc == 0xFFFFu ? plen : p0 + rc);
            P.out_end[gi] = e.x;
            P.out_order[gi] = (int32_t)(e.y & 0xFFFFu);
        }
    }
}

template <int KPL, bool ORDER>
__global__ void __launch_bounds__(W2_WARPS * 32, W2_CTAS) wide2_kernel(const Wide2Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int C = P.C, ld = P.ld, T = P.T;
    const long long R = P.R;

    unsigned char* const wbase = smem_raw + (size_t)warp * P.warp_smem;
    uint32_t* const stage = reinterpret_cast<uint32_t*>(wbase + P.off_stage);    // (ld + 4 words of room in front)
    uint2* const stg = reinterpret_cast<uint2*>(wbase + P.off_stg);              // staged records of the strip
    uint64_t* const bar = reinterpret_cast<uint64_t*>(wbase + P.off_bar);
    uint4* const fifo = reinterpret_cast<uint4*>(wbase + P.off_fifo);            // W2_FIFO x {s lo, s hi, total, p0} {plen, start, -, -}
    const uint32_t cap = P.cap;
    const long long gwarp = (long long)blockIdx.x * W2_WARPS + warp;
    uint2* const ovf = P.ovf + gwarp * P.ovf_cap;
    uint2* const ring = P.ring + gwarp * (long long)P.ring_cap;
    const uint32_t ring_mask = P.ring_cap - 1u;
    uint32_t ring_head = 0, ring_tail = 0;       // records appended / retired so far
    uint32_t fifo_head = 0, fifo_n = 0;          // strips pushed so far / waiting
    const unsigned char* const src_bytes = reinterpret_cast<const unsigned char*>(P.dap);

    // phase B: slot k of lane l is DAP column l + 32 k (only the last slot can lie past the
    // last column; it reads whatever follows the row and is masked)
    bool cvalid[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) cvalid[k] = lane + 32 * k < C;
    uint32_t vm_last = cvalid[KPL - 1] ? 0xFFFFFFFFu : 0u;
    asm volatile("" : "+r"(vm_last));            // keep it a mask: one LOP3 in the row path
    const int ibase = lane * KPL;                // ORDER: first sorted position of the lane

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // ---------------- strip bookkeeping (warp uniform)
    long long c_lo = 0, c_hi = 0;                // units of the cached record run
    int run = 0;
    memo_segment_t seg;
    seg.row_begin = seg.n_rows = 0;
    seg.pos0 = seg.rec_len = seg.rec_id = seg.flags = 0;
    long long r0 = 0, r1 = 0;                    // compare rows of the strip (buffer rows)
    uint32_t pos_r0 = 0, rec_len = 0;
    bool s_last = false, s_chr = false;
    int off = 0;                                 // word offset of the chunk's first row inside the stage
    uint32_t parity = 0, irr = 0;

    auto locate = [&](long long s) {
        if (s < c_lo || s >= c_hi) {
            run = w2_find_run(P, s);
            c_lo = w2_unit_start(P, run);
            c_hi = w2_unit_start(P, run + 1);
            seg = P.n_seg <= W2_INLINE_SEGS ? P.isegs[run] : P.segs[run];
        }
        const long long primed = (seg.flags & MEMO_SEG_PRIMED) ? 1 : 0;
        const long long fc = seg.row_begin + primed;             // first compare row of the run
        const long long end = seg.row_begin + seg.n_rows;
        r0 = fc + (s - c_lo) * R;
        r1 = min(end, r0 + R);
        if (r1 < r0) r1 = r0;                                     // a one-row run: chr-end rows only
        pos_r0 = (uint32_t)seg.pos0 + (uint32_t)(r0 - seg.row_begin);
        rec_len = (uint32_t)seg.rec_len;
        s_last = s + 1 == c_hi;
        s_chr = s_last && (seg.flags & MEMO_SEG_CHR_END);
    };
    // rows [first, first + n) of the buffer into the stage by a 1-D bulk copy from the 16-byte
    // aligned address below the first row (the last < 16 bytes of the buffer are copied by hand)
    auto issue_rows = [&](long long first, long long n) {
        const long long start = first * (long long)ld * 4;
        const long long end = (first + n - 1) * (long long)ld * 4 + (long long)C * 4;
        const long long a0 = start & ~15ll;
        long long a1 = (end + 15) & ~15ll;
        const long long lim = P.total_bytes & ~15ll;
        if (a1 > lim) a1 = lim;
        if (a1 < a0) a1 = a0;
        off = (int)((start - a0) >> 2);
        if (lane == 0) {
            unsigned char* d = reinterpret_cast<unsigned char*>(stage);
            for (long long b = a1; b < end; b += 4)
                *reinterpret_cast<uint32_t*>(d + (b - a0)) = *reinterpret_cast<const uint32_t*>(src_bytes + b);
            if (a1 > a0) {
                mbar_arrive_expect_tx(bar, (uint32_t)(a1 - a0));
                bulk_g2s(d, src_bytes + a0, (uint32_t)(a1 - a0), bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };

    // ---------------- consumer state
    uint32_t A[KPL];                   // ORDER: sorted MEM ends of the previous row, position ibase + k
#pragma unroll
    for (int k = 0; k < KPL; ++k) A[k] = 0;
    uint32_t n_emit = 0;               // index rows of the strip so far

    uint32_t ordk[KPL];                // BED f3 of the slot: ORDER -> position ibase + k; else column lane + 32 k
#pragma unroll
    for (int k = 0; k < KPL; ++k) ordk[k] = (uint32_t)(ORDER ? ibase + k : lane + 32 * k) + 1u;
    // index rows of one DAP row: em[k] / endv[k] per slot; rowcode = the row's offset from the
    // strip's first compare row (0xFFFF: chr-end rows), from which the copy-out rebuilds the BED
    // start.  Output order = f3 ascending.
    auto emit = [&](const bool (&em)[KPL], const uint32_t (&endv)[KPL], uint32_t rowcode) {
        unsigned b[KPL];
        uint32_t total = 0, rank = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            b[k] = __ballot_sync(FULL, em[k]);
            total += __popc(b[k]);
        }
        if (total == 0) return;
        if (ORDER) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) rank += __popc(b[k] & ltmask);
        }
        const uint32_t rc = rowcode << 16;
        rank += n_emit;
        n_emit += total;
        if (n_emit <= cap) {                                          // warp uniform: the usual case
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const uint32_t rk = ORDER ? rank : rank + __popc(b[k] & ltmask);
                if (em[k]) stg[rk] = make_uint2(endv[k], rc | ordk[k]);
                if (ORDER) rank += em[k] ? 1u : 0u; else rank += __popc(b[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const uint32_t rk = ORDER ? rank : rank + __popc(b[k] & ltmask);
                if (em[k]) {
                    const uint2 rec = make_uint2(endv[k], rc | ordk[k]);
                    if (rk < cap) stg[rk] = rec; else ovf[rk - cap] = rec;
                }
                if (ORDER) rank += em[k] ? 1u : 0u; else rank += __popc(b[k]);
            }
        }
    };

    // one flagged row: rp = its raw values (the row before: ld words back), pos = its position
    auto process_row = [&](const uint32_t* rp, uint32_t pos) {
        const uint32_t rowcode = pos - pos_r0;
        const uint32_t* pp = rp - ld;
        uint32_t prev[KPL], dk[KPL], acc = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            prev[k] = pp[lane + 32 * k];
            dk[k] = rp[lane + 32 * k] + 1u - prev[k];
            if (k == KPL - 1) dk[k] &= vm_last;
            acc |= dk[k];
        }
        irr |= acc;
        bool em[KPL];
        uint32_t endv[KPL];
        if (ORDER) {
            // cells whose MEM end moved up (a decrease makes the input irregular: flagged
            // through irr, skipped here)
            unsigned cb[KPL];
            uint32_t nchg = 0;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                cb[k] = __ballot_sync(FULL, (int)dk[k] > 0);
                nchg += __popc(cb[k]);
            }
            if (nchg == 0) return;                                    // (a false positive of phase A)
            if (nchg == 1) {
                // the common case, one cell x -> y: the positions holding x <= A <= y shift
                // down by one, y lands on the first of them, and exactly those can emit
                uint32_t myx = prev[0], myd = dk[0];
                unsigned ball = cb[0];
#pragma unroll
                for (int k = 1; k < KPL; ++k)
                    if (cb[k]) { myx = prev[k]; myd = dk[k]; ball = cb[k]; }
                const int src = __ffs(ball) - 1;
                const uint32_t x = __shfl_sync(FULL, myx, src) + (pos - 1u);
                const uint32_t y = x + __shfl_sync(FULL, myd, src);
                uint32_t up = __shfl_up_sync(FULL, A[KPL - 1], 1);
                if (lane == 0) up = 0xFFFFFFFFu;
                // in place, last slot first: slot kk needs the old value of slot kk - 1
#pragma unroll
                for (int kk = KPL - 1; kk >= 0; --kk) {
                    const uint32_t before = kk == 0 ? up : A[kk - 1];
                    const uint32_t old = A[kk];
                    const bool inr = old - x <= y - x;                // x <= old <= y (y > x)
                    const uint32_t nw = inr ? min(before, y) : old;
                    em[kk] = nw > old && old >= pos && ibase + kk < C;
                    endv[kk] = old;
                    A[kk] = nw;
                }
            } else {
                uint32_t Aold[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) Aold[k] = A[k];
                // one cell per lane and round
                unsigned todo = 0;
#pragma unroll
                for (int k = 0; k < KPL; ++k) todo |= ((int)dk[k] > 0 ? 1u : 0u) << k;
                unsigned m;
                while ((m = __ballot_sync(FULL, todo != 0u)) != 0u) {
                    uint32_t myx = 0, myd = 0;
#pragma unroll
                    for (int k = KPL - 1; k >= 0; --k)
                        if (todo & (1u << k)) { myx = prev[k]; myd = dk[k]; }
                    todo &= todo - 1;
                    do {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        // delete x, insert y > x
                        const uint32_t x = __shfl_sync(FULL, myx, src) + (pos - 1u);
                        const uint32_t y = x + __shfl_sync(FULL, myd, src);
                        uint32_t up = __shfl_up_sync(FULL, A[KPL - 1], 1);
                        if (lane == 0) up = 0xFFFFFFFFu;
#pragma unroll
                        for (int kk = KPL - 1; kk >= 0; --kk) {
                            const uint32_t before = kk == 0 ? up : A[kk - 1];
                            A[kk] = (A[kk] - x <= y - x) ? min(before, y) : A[kk];
                        }
                    } while (m);
                }
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    em[k] = A[k] > Aold[k] && Aold[k] >= pos && ibase + k < C;
                    endv[k] = Aold[k];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const uint32_t e = prev[k] + (pos - 1u);
                em[k] = (int)dk[k] > 0 && e >= pos;
                endv[k] = e;
            }
        }
        emit(em, endv, rowcode);
    };

    volatile unsigned long long* const st = P.status;
    volatile unsigned long long* const gs = P.gstatus;
    // the strip's rows are through: publish its count; the last finisher of a group of 32
    // strips publishes the group's total (nobody waits for anybody)
    auto publish = [&](long long s, unsigned long long total) {
        const long long g = s >> 5;
        unsigned old = 0;
        if (lane == 0) {
            st[s] = (total << 1) | 1ull;
            __threadfence();
            old = atomicAdd(&P.gdone[g], 1u);
        }
        old = __shfl_sync(FULL, old, 0);
        const long long g_first = g << 5;
        const int g_size = (int)min(32ll, P.n_units - g_first);
        if ((int)old != g_size - 1) return;
        __threadfence();
        unsigned long long gt = lane < g_size ? (st[g_first + lane] >> 1) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gt += __shfl_xor_sync(FULL, gt, o);
        if (lane == 0) gs[g] = (gt << 2) | (g == 0 ? 2ull : 1ull);
    };
    // strips of this warp whose rows wait in its ring, oldest first: copied to their place once
    // every earlier strip is published
    auto drain = [&](bool blocking) {
        const unsigned long long ocap = (unsigned long long)P.out_cap;
        while (fifo_n) {
            const uint32_t slot = (fifo_head - fifo_n) & (W2_FIFO - 1);
            const uint4 m0 = fifo[2 * slot], m1 = fifo[2 * slot + 1];
            const long long ps = (long long)(((unsigned long long)m0.y << 32) | m0.x);
            unsigned long long excl = 0;
            if (!w2_resolve(P, ps, lane, &excl)) {
                if (!blocking) break;
                __nanosleep(500);
                continue;
            }
            const uint32_t total = m0.z, p0 = m0.w, plen = m1.x, start = m1.y;
            for (uint32_t i = lane; i < total; i += 32) {
                const uint2 e = ring[(start + i) & ring_mask];
                const unsigned long long gi = excl + i;
                if (gi < ocap) {
                    const uint32_t rcd = e.y >> 16;
                    P.out_start[gi] = (int32_t)(rcd == 0xFFFFu ? plen : p0 + rcd);
                    P.out_end[gi] = e.x;
                    P.out_order[gi] = (int32_t)(e.y & 0xFFFFu);
                }
            }
            ring_tail = start + total;
            --fifo_n;
        }
    };
    // a finished strip (records in the staging area): straight to its place if nothing of this
    // warp waits and every earlier strip is published; else into the warp's ring; a strip that
    // does not fit the ring is parked for wide2_finish_kernel.  Nothing waits for anything.
    auto place = [&](long long ps, uint32_t total, uint32_t p0, uint32_t plen) {
        __syncwarp();
        drain(false);
        if (P.out_cap == 0) return;                                   // counting run: nothing to keep
        const uint32_t ns = total < cap ? total : cap;
        unsigned long long excl = 0;
        if (fifo_n == 0 && w2_resolve(P, ps, lane, &excl)) {
            w2_copy_out(P, stg, ovf, ns, excl, total, p0, plen, lane);
        } else if (fifo_n < W2_FIFO && total <= P.ring_cap - (ring_head - ring_tail)) {
            for (uint32_t i = lane; i < total; i += 32) ring[(ring_head + i) & ring_mask] = i < ns ? stg[i] : ovf[i - ns];
            if (lane == 0) {
                const uint32_t slot = fifo_head & (W2_FIFO - 1);
                fifo[2 * slot] = make_uint4((uint32_t)ps, (uint32_t)((unsigned long long)ps >> 32), total, p0);
                fifo[2 * slot + 1] = make_uint4(plen, ring_head, 0u, 0u);
            }
            ring_head += total;
            ++fifo_head;
            ++fifo_n;
        } else {
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(P.park_cursor, (unsigned long long)total);
            at = __shfl_sync(FULL, at, 0);
            if (at + total <= (unsigned long long)P.park_cap) {
                for (uint32_t i = lane; i < total; i += 32) P.park[at + i] = i < ns ? stg[i] : ovf[i - ns];
                if (lane == 0) P.park_off[ps] = at + 1ull;
            } else {
                // (parking area full: cannot happen with park_cap = out_cap; wait as a last resort)
                while (!w2_resolve(P, ps, lane, &excl)) __nanosleep(500);
                w2_copy_out(P, stg, ovf, ns, excl, total, p0, plen, lane);
            }
        }
        __syncwarp();
    };

    // ---------------- strips
    unsigned long long look = 0;
    if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);
    long long s = (long long)__shfl_sync(FULL, look, 0);
    if (s < P.n_units) {
        locate(s);
        // first chunk: the strip's predecessor row + up to T - 1 compare rows
        issue_rows(r0 - 1, min((long long)T, r1 - r0 + 1));
    }
    const int shift = ld & 3;
    while (s < P.n_units) {
        // ---- the strip's rows
        bool claimed = false;
        if (r1 - (r0 - 1) <= T) {                    // a single chunk: the next strip is claimed now
            if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);
            claimed = true;
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        if (ORDER) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) A[k] = cvalid[k] ? stage[off + lane + 32 * k] + (pos_r0 - 1u) : 0u;
            group_sort_desc<32, KPL>(A, lane);
        }
        n_emit = 0;
        long long cur = r0 - 1;                      // first row of the chunk in the stage
        bool first = true;
        const uint32_t* lastptr = stage + off;       // raw values of the strip's last row
        for (;;) {
            const int n = (int)min((long long)T, r1 - cur);          // rows in the stage
            const int skip = first ? 1 : 0;                          // (the predecessor row is no compare row)
            if (n > skip) {
                uint32_t rowmask;
                switch (shift) {
                    case 0: rowmask = scan_chunk<0>(stage, off, n, ld, P.magic, lane); break;
                    case 1: rowmask = scan_chunk<1>(stage, off, n, ld, P.magic, lane); break;
                    case 2: rowmask = scan_chunk<2>(stage, off, n, ld, P.magic, lane); break;
                    default: rowmask = scan_chunk<3>(stage, off, n, ld, P.magic, lane); break;
                }
                // bit j + 1 = row j of the chunk
                rowmask = (rowmask >> 1) & ((1u << n) - 1u) & ~((1u << skip) - 1u);
                const uint32_t pos_c = pos_r0 - 1u + (uint32_t)(cur - (r0 - 1));
                while (rowmask) {
                    const int j = __ffs(rowmask) - 1;
                    rowmask &= rowmask - 1;
                    process_row(stage + off + j * ld, pos_c + (uint32_t)j);
                }
            }
            lastptr = stage + off + (n - 1) * ld;
            cur += n;
            first = false;
            if (cur >= r1) break;
            // the strip goes on: its last row moves in front of where the next chunk lands
            const int n_next = (int)min((long long)T, r1 - cur);
            const int off_next = (int)(((cur * (long long)ld * 4) & 15ll) >> 2);
            {
                // (words that end up at stage[0 ..) arrive with the bulk copy itself)
                const int keep = ld - off_next;
                uint32_t w[KPL + 1];
#pragma unroll
                for (int k = 0; k <= KPL; ++k) w[k] = lane + 32 * k < keep ? lastptr[lane + 32 * k] : 0u;
                __syncwarp();
#pragma unroll
                for (int k = 0; k <= KPL; ++k)
                    if (lane + 32 * k < keep) stage[off_next - ld + lane + 32 * k] = w[k];
            }
            __syncwarp();                                            // the stage is free again
            issue_rows(cur, n_next);
            if (!claimed && cur + n_next >= r1) {
                // the strip's last chunk: claim the next strip now -- late, so that few strips are
                // claimed but not started (the strips after them would have to be parked), early
                // enough to hide the atomic
                if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);
                claimed = true;
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        if (s_chr) {                                                 // chr-end rows after the run's last row
            bool em[KPL];
            uint32_t endv[KPL];
            const uint32_t pos_last = pos_r0 + (uint32_t)(r1 - r0) - 1u;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const uint32_t e = ORDER ? A[k] : lastptr[lane + 32 * k] + pos_last;
                const bool valid = ORDER ? (ibase + k < C) : cvalid[k];
                em[k] = valid && e >= rec_len;
                endv[k] = min(e, 2u * rec_len);
            }
            emit(em, endv, 0xFFFFu);
        }
        // ---- publish the strip's count and place its rows; the next strip's first load goes out
        //      before either
        const long long fin_s = s;
        const uint32_t fin_total = n_emit, fin_p0 = pos_r0, fin_len = rec_len;
        __syncwarp();                                                // the stage is free
        s = (long long)__shfl_sync(FULL, look, 0);
        if (s < P.n_units) {
            locate(s);
            issue_rows(r0 - 1, min((long long)T, r1 - r0 + 1));
        }
        publish(fin_s, fin_total);
        place(fin_s, fin_total, fin_p0, fin_len);
    }
    drain(true);
    if (irr >> 31) P.result[MEMO_RES_IRREGULAR] = 1;
}

// After the strip kernel: rows up to the end of every record run (seg_out_end), the grand
// total, and the parked strips to their places.  Every count is published by now.
__global__ void __launch_bounds__(256) wide2_finish_kernel(const Wide2Params P) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i <= P.n_seg; i += n_warps) {
        const long long last = (i < P.n_seg ? w2_unit_start(P, (int)i + 1) : P.n_units) - 1;   // i == n_seg: everything
        unsigned long long excl = 0;
        w2_resolve(P, last, lane, &excl);
        const unsigned long long incl = excl + (P.status[last] >> 1);
        if (lane == 0) {
            if (i < P.n_seg) P.seg_out_end[i] = (int64_t)incl; else P.result[MEMO_RES_N_OUT] = (int64_t)incl;
        }
    }
    if (warp == 0 && lane == 0) P.result[3] = (int64_t)*P.park_cursor;      // (stat: index rows that were parked)
    if (P.out_cap == 0) return;
    for (long long s0 = warp * 32; s0 < P.n_units; s0 += n_warps * 32) {
        const unsigned long long po = s0 + lane < P.n_units ? P.park_off[s0 + lane] : 0ull;
        unsigned m = __ballot_sync(FULL, po != 0ull);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const long long s = s0 + src;
            const unsigned long long at = __shfl_sync(FULL, po, src) - 1ull;
            const int run = w2_find_run(P, s);
            const memo_segment_t seg = P.n_seg <= W2_INLINE_SEGS ? P.isegs[run] : P.segs[run];
            const long long fc = seg.row_begin + ((seg.flags & MEMO_SEG_PRIMED) ? 1 : 0);
            const long long r0 = fc + (s - w2_unit_start(P, run)) * (long long)P.R;
            const uint32_t p0 = (uint32_t)seg.pos0 + (uint32_t)(r0 - seg.row_begin);
            const uint32_t total = (uint32_t)(P.status[s] >> 1);
            unsigned long long excl = 0;
            w2_resolve(P, s, lane, &excl);
            w2_copy_out(P, P.park + at, P.park + at, total, excl, total, p0, (uint32_t)seg.rec_len, lane);
        }
    }
}

typedef void (*wide2_kernel_t)(const Wide2Params);

wide2_kernel_t select_wide2(int kpl, bool order) {
#define MEMO_W2(KK) \
    if (kpl == KK) return order ? wide2_kernel<KK, true> : wide2_kernel<KK, false>;
    MEMO_W2(1) MEMO_W2(2) MEMO_W2(3) MEMO_W2(4) MEMO_W2(5) MEMO_W2(6) MEMO_W2(7) MEMO_W2(8)
    MEMO_W2(9) MEMO_W2(10) MEMO_W2(11) MEMO_W2(12) MEMO_W2(13) MEMO_W2(14) MEMO_W2(15) MEMO_W2(16)
#undef MEMO_W2
    return nullptr;
}

struct Wide2Plan {
    int kpl, T, R;
    uint32_t off_stage, off_stg, off_bar, off_fifo, warp_smem, cap, ring_cap;
    size_t smem;
    long long n_units, grid_warps, ovf_cap;
    long long park_cap;
    size_t off_segs, off_ustart, off_status, off_gdone, off_gstatus, off_parkoff, off_ctrl, off_ovf, off_ring, off_park, total;
};

// resident CTAs per SM of a kernel variant on a device for a dynamic shared-memory size.  The
// kernel's shared-memory limit is raised to the device maximum once per (device, kernel) --
// never lowered again, whatever sizes later launches use -- and the occupancy query is cached.
int wide2_ctas_per_sm(wide2_kernel_t kern, size_t smem) {
    struct Entry { int dev; wide2_kernel_t k; size_t smem; int n; };
    static std::mutex mu;
    static Entry cache[64];
    static int n_cache = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    bool raised = false;
    for (int i = 0; i < n_cache; ++i) {
        if (cache[i].dev != dev || cache[i].k != kern) continue;
        raised = true;
        if (cache[i].smem == smem) return cache[i].n;
    }
    if (!raised &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, W2_WARPS * 32, smem) != cudaSuccess) n = 0;
    if (n_cache == 64) n_cache = 0;                                // (start over: the limit stays raised)
    cache[n_cache++] = Entry{dev, kern, smem, n};
    return n;
}

int make_wide2_plan(int64_t rows, int32_t C, int32_t ld, int64_t out_cap, const memo_segment_t* segs, int32_t n_seg,
                    const memo_index_opts_t* opts, Wide2Plan* plan, long long* ustart) {
    plan->kpl = (C + 31) / 32;
    const long long row_bytes = (long long)ld * 4;
    // shared memory of one warp, W2_CTAS CTAs of W2_WARPS warps per SM (227 KB - 1 KB per CTA):
    //   room for the row before the chunk | stage: T rows (+ alignment slack, + what phase B
    //   reads past the last row) | barrier | table of waiting strips | staged index rows (8 B each)
    const long long per_warp = ((227 * 1024) / W2_CTAS - 1024) / W2_WARPS / 128 * 128;
    const long long front = (long long)align_up((size_t)row_bytes + 16, 128);
    const long long over = 144;                                   // phase B reads up to 31 words past a row
    const long long fixed = front + 32 + over + 16 + 32 * W2_FIFO;
    // ~65 % of the warp's share for DAP rows, the rest for staged index rows
    long long t = (opts && opts->rows_per_tile > 0) ? opts->rows_per_tile : (per_warp * 21 / 32 - fixed) / row_bytes;
    if (t > W2_MAX_T) t = W2_MAX_T;
    while (t > 2 && (fixed + t * row_bytes + 16 * 64 > 55 * 1024 || t * ld > W2_MAX_WORDS)) --t;   // (very wide rows)
    if (t < 2) t = 2;
    plan->T = (int)t;
    // about 128 rows per strip, a whole number of chunks (the predecessor row is one of them)
    long long r = (opts && opts->emit_buf_records > 0) ? opts->emit_buf_records
                                                       : (128 / t > 1 ? 128 / t : 1) * t - 1;
    if (r < 1) r = 1;
    if (r > 60000) r = 60000;                                     // (16-bit row codes in the staged records)
    plan->R = (int)r;
    size_t o = (size_t)front;
    plan->off_stage = (uint32_t)o;  o = align_up(o + (size_t)(t * row_bytes) + 32 + (size_t)over, 16);
    plan->off_bar = (uint32_t)o;    o += 16;
    plan->off_fifo = (uint32_t)o;   o += 32 * W2_FIFO;
    plan->off_stg = (uint32_t)o;
    long long cap = (long long)o + 8 * 64 <= per_warp ? (per_warp - (long long)o) / 8 : 64;
    if (cap > 2048) cap = 2048;
    cap = cap / 32 * 32;
    if (cap < 32) cap = 32;
    plan->cap = (uint32_t)cap;
    plan->warp_smem = (uint32_t)align_up(o + 8 * (size_t)cap, 128);
    plan->smem = (size_t)plan->warp_smem * W2_WARPS;
    MEMO_REQUIRE(plan->smem <= 227 * 1024, "strip configuration needs %zu B of shared memory", plan->smem);

    long long u = 0;
    int64_t prev_end = 0;
    for (int i = 0; i < n_seg; ++i) {
        const memo_segment_t& s = segs[i];
        MEMO_REQUIRE(s.n_rows > 0, "segment %d has no rows", i);
        MEMO_REQUIRE(s.row_begin >= prev_end && s.row_begin + s.n_rows <= rows,
                     "segment %d out of order or out of range", i);
        MEMO_REQUIRE((s.flags & MEMO_SEG_PRIMED) || (s.row_begin >= 1 && s.pos0 >= 1),
                     "segment %d: continuation run needs a halo row before it", i);
        MEMO_REQUIRE(s.pos0 >= 0 && s.rec_len >= 1 && (int64_t)s.pos0 + s.n_rows <= 2147483647LL,
                     "segment %d: positions exceed int32", i);
        prev_end = s.row_begin + s.n_rows;
        if (ustart) ustart[i] = u;
        const long long m = s.n_rows - ((s.flags & MEMO_SEG_PRIMED) ? 1 : 0);      // compare rows
        const long long nt = (m + r - 1) / r;
        u += nt > 0 ? nt : 1;                                     // a one-row run still owns its chr-end rows
    }
    if (ustart) ustart[n_seg] = u;
    plan->n_units = u;
    const long long sm_warps = (long long)device_sm_count() * W2_CTAS * W2_WARPS;
    const long long unit_warps = (u + W2_WARPS - 1) / W2_WARPS * W2_WARPS;
    plan->grid_warps = unit_warps < sm_warps ? unit_warps : sm_warps;
    // a strip emits at most (rows of the strip + 1) * C index rows
    const long long strip_rows = rows < r ? rows : r;
    plan->ovf_cap = (strip_rows + 1) * C;
    size_t off = 0;
    plan->off_segs = off;    off = align_up(off + sizeof(memo_segment_t) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
    plan->off_ustart = off;  off = align_up(off + sizeof(long long) * (size_t)(n_seg + 1), 256);
    plan->off_ctrl = off;    off = align_up(off + 256, 256);
    const size_t ng = (size_t)((u + 31) / 32 + 1);
    plan->off_status = off;  off = align_up(off + 8 * (size_t)(u > 0 ? u : 1), 256);
    plan->off_gdone = off;   off = align_up(off + 4 * ng, 256);
    plan->off_gstatus = off; off = align_up(off + 8 * ng, 256);
    plan->off_parkoff = off; off = align_up(off + 8 * (size_t)(u > 0 ? u : 1), 256);
    plan->off_ovf = off;     off = align_up(off + 8 * (size_t)plan->ovf_cap * (size_t)plan->grid_warps, 256);
    // ring of a warp: ~10 average strips (it stays in L2: written and read back within microseconds)
    plan->ring_cap = 2048;
    plan->off_ring = off;    off = align_up(off + 8 * (size_t)plan->ring_cap * (size_t)plan->grid_warps, 256);
    // parking area: strips whose predecessors were late (at most every index row once)
    plan->park_cap = out_cap;
    plan->off_park = off;    off = align_up(off + 8 * (size_t)plan->park_cap + 256, 256);
    plan->total = off;
    return MEMO_OK;
}

}  // namespace

bool wide2_supported(int32_t n_cols, int32_t ld) {
    return n_cols >= 1 && n_cols <= 512 && ld >= n_cols && ld <= 2 * n_cols + 8;
}

size_t wide2_workspace_bytes(int64_t rows, int32_t n_cols, int32_t ld, int64_t out_cap, const memo_segment_t* segs,
                             int32_t n_seg, const memo_index_opts_t* opts) {
    Wide2Plan plan;
    if (make_wide2_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, nullptr) != MEMO_OK) return 0;
    return plan.total;
}

int launch_wide2(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld, const memo_segment_t* segs,
                 int32_t n_seg, const memo_index_opts_t* opts, int32_t* out_start, uint32_t* out_end,
                 int32_t* out_order, int64_t out_cap, int64_t* seg_out_end, int64_t* result,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    Wide2Plan plan;
    std::vector<long long> ustart((size_t)n_seg + 1);
    int rc = make_wide2_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, ustart.data());
    if (rc != MEMO_OK) return rc;
    if (workspace_bytes < plan.total || workspace == nullptr) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, plan.total);
        return MEMO_ERR_WORKSPACE;
    }
    MEMO_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(int64_t) * MEMO_RES_SLOTS, stream));
    if (n_seg == 0 || plan.n_units == 0) return MEMO_OK;
    const bool order = opts ? (opts->order_mode != 0) : true;
    wide2_kernel_t kern = select_wide2(plan.kpl, order);
    if (!kern) {
        set_error("no strip kernel for n_cols=%d", n_cols);
        return MEMO_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);

    Wide2Params P{};
    P.dap = dap;
    P.total_bytes = (long long)rows * ld * 4;
    P.C = n_cols; P.ld = ld; P.T = plan.T; P.R = plan.R;
    P.magic = (uint32_t)(((1ull << 32) + (unsigned long long)ld - 1ull) / (unsigned long long)ld);
    P.n_seg = n_seg;
    P.n_units = plan.n_units;
    if (n_seg <= W2_INLINE_SEGS) {
        for (int i = 0; i < n_seg; ++i) { P.isegs[i] = segs[i]; P.iunit[i] = ustart[i]; }
        P.iunit[n_seg] = ustart[n_seg];
    } else {
        // pageable host -> device copies are staged by the runtime before returning
        MEMO_CUDA_TRY(cudaMemcpyAsync(ws + plan.off_segs, segs, sizeof(memo_segment_t) * (size_t)n_seg,
                                      cudaMemcpyHostToDevice, stream));
        MEMO_CUDA_TRY(cudaMemcpyAsync(ws + plan.off_ustart, ustart.data(), sizeof(long long) * (size_t)(n_seg + 1),
                                      cudaMemcpyHostToDevice, stream));
        P.segs = reinterpret_cast<const memo_segment_t*>(ws + plan.off_segs);
        P.seg_unit_start = reinterpret_cast<const long long*>(ws + plan.off_ustart);
    }
    P.warp_smem = plan.warp_smem; P.off_stage = plan.off_stage;
    P.off_stg = plan.off_stg; P.off_bar = plan.off_bar; P.off_fifo = plan.off_fifo; P.cap = plan.cap;
    P.ring = reinterpret_cast<uint2*>(ws + plan.off_ring);
    P.ring_cap = plan.ring_cap;
    P.strip_counter = reinterpret_cast<unsigned long long*>(ws + plan.off_ctrl);
    P.park_cursor = reinterpret_cast<unsigned long long*>(ws + plan.off_ctrl + 128);
    P.park_off = reinterpret_cast<unsigned long long*>(ws + plan.off_parkoff);
    P.park = reinterpret_cast<uint2*>(ws + plan.off_park);
    P.park_cap = plan.park_cap;
    P.status = reinterpret_cast<unsigned long long*>(ws + plan.off_status);
    P.gdone = reinterpret_cast<unsigned int*>(ws + plan.off_gdone);
    P.gstatus = reinterpret_cast<unsigned long long*>(ws + plan.off_gstatus);
    P.ovf = reinterpret_cast<uint2*>(ws + plan.off_ovf);
    P.ovf_cap = plan.ovf_cap;
    P.out_start = out_start; P.out_end = out_end; P.out_order = out_order; P.out_cap = out_cap;
    P.seg_out_end = seg_out_end; P.result = result;

    int per_sm = wide2_ctas_per_sm(kern, plan.smem);
    if (per_sm < 1) {
        set_error("strip kernel does not fit on an SM (smem %zu B)", plan.smem);
        return MEMO_ERR_UNSUPPORTED;
    }
    if (opts && opts->ctas_per_sm > 0 && opts->ctas_per_sm < per_sm) per_sm = opts->ctas_per_sm;
    if (per_sm > W2_CTAS) per_sm = W2_CTAS;                       // (the per-warp areas are sized for that many)
    long long grid = (long long)device_sm_count() * per_sm;
    const long long need = (plan.n_units + W2_WARPS - 1) / W2_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    // strip counter, parking cursor, status words of all strips, group counters and status words,
    // parking table
    MEMO_CUDA_TRY(cudaMemsetAsync(ws + plan.off_ctrl, 0, plan.off_ovf - plan.off_ctrl, stream));
    profile_begin(stream);
    kern<<<(unsigned)grid, W2_WARPS * 32, plan.smem, stream>>>(P);
    MEMO_LAUNCH_CHECK(1);
    profile_end(stream);
    {
        long long fgrid = (plan.n_units + 32 * 8 - 1) / (32 * 8);
        const long long fmax = (long long)device_sm_count() * 4;
        if (fgrid > fmax) fgrid = fmax;
        if (fgrid < 1) fgrid = 1;
        wide2_finish_kernel<<<(unsigned)fgrid, 256, 0, stream>>>(P);
        MEMO_LAUNCH_CHECK(1);
    }
    return MEMO_OK;
}

}  // namespace memo
