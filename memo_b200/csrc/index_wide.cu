// DAP -> MEMO index rows on sm_100a: single-pass build for WIDE rows (any
// n_cols <= 512; the 94-genome configurations), one warp per strip of rows.
//
// Replaces the hot loop of the reference's src/dap_to_bed.py (--mem --overlap
// [--order]): get_new_record :85-91, dap_to_mem :116-134, print_interval /
// overlaps :93-109.  Mathematics as in index_build.cu (DESIGN.md "index build"):
// with E[r][c] = p(r) + v[r][c] and A[r] = E[r] sorted descending (--order) or
// E[r] itself, row r emits (p, A[r-1][j], j+1) for every j with A[r][j] >
// A[r-1][j] and A[r-1][j] >= p, provided no E decreases down a column (matching
// statistics).
//
// The reference sorts every row (:89-90).  Here a row is never sorted: a warp
// walks a strip of R consecutive rows and keeps A, the sorted MEM ends of the
// previous row, in registers (lane l holds positions l*KPL .. l*KPL+KPL-1).  E
// only moves where a new MEM starts, so for most rows nothing changes: the row
// loop only loads a row, compares it with the one before and votes (14 warp
// instructions per row at KPL = 3).  A row that moved something is handled
// outside that loop: a cell whose MEM end moves from x to y > x is a
// delete/insert in A -- the positions holding x <= A <= y shift down by one and
// y lands on the first of them, one shuffle, a min and a select per slot
// instead of a sort.  The index rows of the row are the positions where A
// changed.  A is sorted from scratch once per strip.
//
// Strips are handed out by an atomic counter (one per R rows).  A strip streams
// through the warp's private shared-memory stage in chunks of T rows (bulk async
// copies + mbarriers; the strip's next chunk is prefetched into L2 with
// evict_last priority, the copy that consumes it marks the lines evict_first),
// so a warp never waits for another warp.  Index rows go straight to the
// scratch area as 8-byte rows {end, row << 16 | order}: a warp reserves P.chunk
// rows at a time (one atomicAdd) and appends to them, so a strip's output is one
// block of consecutive scratch rows, or a short chain of blocks when it crosses
// into the warp's next chunk.  tile_scan_kernel / strip_gather_kernel
// (index_build.cu) copy the blocks into the ordered output.
#include "index_fast.cuh"
#include "warp_sort.cuh"

namespace memo {
namespace {

// (experiment knob: -DMEMO_WIDE_MIN_CTAS=4 caps the kernel at 64 registers for 32 resident warps;
//  measured slower, DESIGN.md 4.3b.  Without it ptxas settles on 80 registers = 4 CTAs of 6 warps;
//  note that __launch_bounds__(256, 1) is NOT the same as __launch_bounds__(256): it lets ptxas
//  take 105 registers and halves the occupancy)
#ifdef MEMO_WIDE_MIN_CTAS
#ifndef MEMO_WIDE_MAX_THREADS
#define MEMO_WIDE_MAX_THREADS 256
#endif
#define MEMO_WIDE_BOUNDS __launch_bounds__(MEMO_WIDE_MAX_THREADS, MEMO_WIDE_MIN_CTAS)
#else
#define MEMO_WIDE_BOUNDS __launch_bounds__(256)
#endif
template <int KPL, bool ORDER>
__global__ void MEMO_WIDE_BOUNDS wide_kernel(const FastParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int C = P.C, S = P.stages, T = P.T, R = P.R;
    const int ld_param = P.ld;

    unsigned char* const wbase = smem_raw + (size_t)warp * P.warp_smem;
    uint64_t* const bars = (uint64_t*)(wbase + P.off_bars);
    TileDesc* const descs = (TileDesc*)(wbase + P.off_descs);

    // ---------------- producer state (lane 0): the chunk sequence of the warp's strips
    long long c_lo = 0, c_hi = 0;
    memo_segment_t seg;
    seg.row_begin = seg.n_rows = 0;
    seg.pos0 = seg.rec_len = seg.rec_id = seg.flags = 0;
    unsigned long long look = 0;          // next strip (fetched one strip ahead: hides the atomic)
    unsigned long long p_addr = 0;        // byte offset of the strip's next row in the DAP
    uint32_t p_left = 0, p_pos = 0;       // rows of the strip still to fetch; position of the next one
    int p_strip = 0;
    bool p_first = false, p_done = false, p_lastseg = false;
    const uint32_t ldb = (uint32_t)ld_param * 4u, Cb = (uint32_t)C * 4u;
    const unsigned long long lim = (unsigned long long)P.total_bytes & ~15ull;
    const unsigned char* const src = reinterpret_cast<const unsigned char*>(P.dap);
    if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();

    // (per chunk: 32-bit arithmetic on top of the strip's running byte offset)
    auto issue = [&](int s) {
        if (p_done) return;
        uint64_t* bar = &bars[s];
        if (p_left == 0) {
            const long long strip = (long long)look;
            if (strip >= P.n_tiles) {
                TileDesc d;
                d.n = 0; d.off = 0; d.pos_h = 0; d.rec_len = 0; d.flags = WD_END; d.r_lo = 0; d.r_hi = 0; d.pad = 0;
                descs[s] = d;
                mbar_arrive(bar);
                p_done = true;
                return;
            }
            look = atomicAdd(P.strip_counter, 1ull);
            if (strip < c_lo || strip >= c_hi) {
                int s_lo = 0, s_hi = P.n_seg - 1;
                while (s_lo < s_hi) {
                    const int mid = (s_lo + s_hi + 1) >> 1;
                    if (P.seg_tile_start[mid] <= strip) s_lo = mid; else s_hi = mid - 1;
                }
                c_lo = P.seg_tile_start[s_lo];
                c_hi = P.seg_tile_start[s_lo + 1];
                seg = P.segs[s_lo];
            }
            const long long t = strip - c_lo;
            const int primed = (seg.flags & MEMO_SEG_PRIMED) ? 1 : 0;
            const long long m = seg.n_rows - primed;                     // compare rows of the run
            long long n_cmp = m - t * R;
            if (n_cmp > R) n_cmp = R;
            if (n_cmp < 0) n_cmp = 0;
            const long long row = seg.row_begin + primed + t * R - 1;    // the strip's predecessor row
            p_addr = (unsigned long long)row * ldb;
            p_left = (uint32_t)n_cmp + 1u;
            p_pos = (uint32_t)seg.pos0 + (uint32_t)(row - seg.row_begin);
            p_first = true;
            p_lastseg = (strip + 1 == c_hi) && (seg.flags & MEMO_SEG_CHR_END);
            p_strip = (int)strip;
        }
        const uint32_t n = p_left < (uint32_t)T ? p_left : (uint32_t)T;
        const uint32_t head = (uint32_t)p_addr & 15u;               // the copy starts at a 16-byte boundary
        const uint32_t bytes = (n - 1u) * ldb + Cb;                 // first row's first byte .. last row's last column
        uint32_t len = (head + bytes + 15u) & ~15u;
        const unsigned long long a0 = p_addr - head;
        const bool last = n == p_left;
        TileDesc d;
        d.n = (int)n;
        d.off = (int)(head >> 2);
        d.pos_h = p_pos;
        d.rec_len = (uint32_t)seg.rec_len;
        d.flags = (p_first ? WD_FIRST : 0) | (last ? WD_LAST : 0) | ((last && p_lastseg) ? WD_CHR : 0);
        d.r_lo = p_strip;
        d.r_hi = 0;
        d.pad = ld_param;                // the row stride: read back per chunk, it then lives in a register
                                         // instead of being reloaded from the parameter bank per row
        descs[s] = d;
        unsigned char* data = wbase + (size_t)s * P.stage_bytes;
        if (a0 + len > lim) {
            // the last < 16 bytes of the buffer cannot be part of a 16-byte granular bulk copy
            const unsigned long long end = p_addr + bytes;
            const unsigned long long a1 = lim > a0 ? lim : a0;
            for (unsigned long long b = a1; b < end; b += 4)
                *reinterpret_cast<uint32_t*>(data + (b - a0)) = *reinterpret_cast<const uint32_t*>(src + b);
            len = (uint32_t)(a1 - a0);
        }
        if (len) {
            mbar_arrive_expect_tx(bar, len);
            bulk_g2s_hint(data, src + a0, len, bar, pol_first);
        } else {
            mbar_arrive(bar);
        }
        p_addr += (unsigned long long)n * ldb;
        p_pos += n;
        p_left -= n;
        p_first = false;
        // the strip's next chunk on its way into L2 while this one is worked on: its copy, issued when
        // this stage is free again, then finds the rows in L2 (one stage per warp: the wait for a
        // chunk is the latency of its copy)
        if (P.prefetch && p_left > 0) {
            const uint32_t nn = p_left < (uint32_t)T ? p_left : (uint32_t)T;
            const uint32_t h2 = (uint32_t)p_addr & 15u;
            const uint32_t l2 = (h2 + nn * ldb + 15u) & ~15u;
            if (p_addr - h2 + l2 <= lim) bulk_prefetch_l2_hint(src + (p_addr - h2), l2, pol_last);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
        for (int s = 0; s < S; ++s) issue(s);

    // ---------------- consumer state
    // raw columns: slot k of lane l is DAP column l + 32 k.  Slots past the last
    // column read whatever follows the row (the stage is padded) and are masked.
    bool cvalid[KPL];
    uint32_t vm[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        cvalid[k] = lane + 32 * k < C;
        vm[k] = cvalid[k] ? 0xFFFFFFFFu : 0u;
        asm volatile("" : "+r"(vm[k]));        // keep it a mask: one LOP3 per slot in the row loop
    }
    const int ibase = lane * KPL;      // ORDER: first sorted position of the lane
    uint32_t prv[KPL];                 // previous row, raw values
    uint32_t A[KPL];                   // ORDER: sorted MEM ends of the previous row, position ibase + k
#pragma unroll
    for (int k = 0; k < KPL; ++k) prv[k] = A[k] = 0;
    StripOut so;                       // the strip's output blocks
    uint32_t strip_pos0 = 0;           // position of the strip's predecessor row: scratch rows count from it
    uint32_t irr_acc = 0;

    // index rows of one row: em[k] / endv[k] per slot, `row` = BED start - strip_pos0 (or
    // SCR_ROW_CHR: chr-end rows).  Output
    // order: ORDER -> position ibase + k; else column lane + 32 k.
    auto emit = [&](const bool (&em)[KPL], const uint32_t (&endv)[KPL], uint32_t row) {
        unsigned b[KPL];
        uint32_t total = 0, rank = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            b[k] = __ballot_sync(FULL, em[k]);
            total += __popc(b[k]);
        }
        // (ranks before the branch on the total: their popcounts overlap the votes' latency)
        if (ORDER) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) rank += __popc(b[k] & ltmask);
        }
        if (total == 0) return;
        uint32_t* const dst0 = so.reserve(P, total, lane);             // warp uniform, always writable
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const uint32_t rk = ORDER ? rank : rank + __popc(b[k] & ltmask);
            if (em[k])
                scr_store(dst0 + rk * SCR_WORDS, endv[k], row, (uint32_t)(ORDER ? ibase + k : lane + 32 * k) + 1u);
            if (ORDER) rank += em[k] ? 1u : 0u; else rank += __popc(b[k]);
        }
    };

    // a row in which some cell moved: `prev` / `cur` = raw values of the row before / the row,
    // dk = cur + 1 - prev (last slot masked), `pos` = the row's position
    auto changed = [&](const uint32_t (&prev)[KPL], const uint32_t (&cur)[KPL], const uint32_t (&dk)[KPL],
                       uint32_t pos) {
        bool em[KPL];
        uint32_t endv[KPL];
        if (ORDER) {
            // cells whose MEM end moved up (a decrease makes the input irregular:
            // flagged through irr_acc, skipped here)
            bool ch[KPL];
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                ch[k] = (int)dk[k] > 0;
                cnt += ch[k] ? 1u : 0u;
            }
            // x -> y of the common case (one changed cell; every other lane contributes 0 to the two
            // maxima), reduced next to the cell count rather than after the branch on it: three
            // warp reductions in flight at once (A/B on one box: 18.93 -> 18.80 ms at chr1 x 94)
            uint32_t myx = 0, myy = 0;
#pragma unroll
            for (int k = 0; k < KPL; ++k)
                if (ch[k]) { myx = prev[k]; myy = cur[k]; }
            const uint32_t nchg = __reduce_add_sync(FULL, cnt);
            const uint32_t x1 = __reduce_max_sync(FULL, myx);
            const uint32_t y1 = __reduce_max_sync(FULL, myy);
            const uint32_t up1 = __shfl_up_sync(FULL, A[KPL - 1], 1);
            if (nchg == 1) {
                // one cell x -> y: the positions holding x <= A <= y shift down by one, y lands
                // on the first of them, and exactly those positions can emit.  (Positions past
                // the last column hold 0 and never emit: a compare row has pos >= 1.)
                const uint32_t x = x1 + (pos - 1u);
                const uint32_t y = y1 + pos;
                uint32_t up = up1;
                if (lane == 0) up = 0xFFFFFFFFu;
                // in place, last slot first: slot kk needs the old value of slot kk - 1
#pragma unroll
                for (int kk = KPL - 1; kk >= 0; --kk) {
                    const uint32_t before = kk == 0 ? up : A[kk - 1];
                    const uint32_t old = A[kk];
                    // (x <= old <= y ? min(before, y) : old, with one compare: above y the minimum
                    //  is y < old, inside the range it is >= old because before >= old)
                    const uint32_t nw = old >= x ? max(old, min(before, y)) : old;
                    em[kk] = nw != old && old >= pos;
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
                for (int k = 0; k < KPL; ++k) todo |= (ch[k] ? 1u : 0u) << k;
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
                            A[kk] = A[kk] >= x ? max(A[kk], min(before, y)) : A[kk];
                        }
                    } while (m);
                }
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    em[k] = A[k] != Aold[k] && Aold[k] >= pos;
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
        emit(em, endv, pos - strip_pos0);
    };

    int s = 0;
    uint32_t parity = 0;
    for (;;) {
        mbar_wait(&bars[s], parity);
        TileDesc d = descs[s];
        // (warp reductions: the compiler then knows that flags and row count are warp uniform and
        //  emits no divergence guards around the votes of the row loop)
        d.flags = (int)__reduce_or_sync(FULL, (unsigned)d.flags);
        d.n = (int)__reduce_max_sync(FULL, (unsigned)d.n);
        if (d.flags & WD_END) break;
        const int ld = d.pad;
        // the lane's column of the stage: row r, slot k at lp[r * ld + 32 * k]
        const uint32_t* lp = reinterpret_cast<const uint32_t*>(wbase + (size_t)s * P.stage_bytes) + d.off + lane;
        uint32_t pos_end = d.pos_h + (uint32_t)d.n;              // position after the chunk's last row
        int left = d.n;                                          // rows from lp on
        if (d.flags & WD_FIRST) {
            // strip start: row 0 only primes the state
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = lp[32 * k];
            if (ORDER) {
#pragma unroll
                for (int k = 0; k < KPL; ++k) A[k] = cvalid[k] ? prv[k] + d.pos_h : 0u;
                group_sort_desc<32, KPL>(A, lane);
            }
            so.begin(d.r_lo);
            strip_pos0 = d.pos_h;
            lp += ld;
            --left;
        }
        // The chunk's rows.  Most rows move nothing: the inner loop only loads a row, compares it with
        // the one before and votes; it touches neither the sorted row nor the output state, alternates
        // between two register sets (no copies) and hands on nothing but the place of a row that moved
        // something (lp, left): that row and the one before it are then read again from the stage
        // (the row before the chunk's first row lives in registers only: `carry`).
        const int left_first = left;
        uint32_t carry[KPL];
#pragma unroll
        for (int k = 0; k < KPL; ++k) carry[k] = prv[k];
        for (;;) {
            {
                uint32_t ra[KPL], rb[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) ra[k] = prv[k];
                // a row against the one before: true if some cell moved (warp vote)
                auto moved = [&](const uint32_t* rowp, const uint32_t (&before)[KPL], uint32_t (&row)[KPL]) {
#ifndef MEMO_SCAN_OR
                    // (row - before == -1 in every cell of an unchanged row: differences ANDed, one
                    //  compare; the subtraction can go to the FMA pipe, the integer ALU runs at half rate)
                    uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) row[k] = rowp[32 * k];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const uint32_t dd = row[k] - before[k];
                        acc &= (k == KPL - 1) ? (dd | ~vm[k]) : dd;
                    }
                    return __any_sync(FULL, acc != 0xFFFFFFFFu) != 0;
#else               // (experiment knob: row + 1 - before ORed, three-input adds on the integer ALU)
                    uint32_t acc = 0;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) row[k] = rowp[32 * k];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const uint32_t dd = row[k] + 1u - before[k];
                        acc |= (k == KPL - 1) ? (dd & vm[k]) : dd;
                    }
                    return __any_sync(FULL, acc != 0u) != 0;
#endif
                };
                // (leaves the loop with left > 0 and lp at a row that moved something, or left == 0)
                while (left >= 2) {
                    if (moved(lp, ra, rb)) break;
                    if (moved(lp + ld, rb, ra)) {
                        lp += ld;
                        --left;
                        break;
                    }
                    lp += 2 * ld;
                    left -= 2;
                }
                // (after a hit on the second row of a pair with one row left, `ra` is the hit row itself:
                //  the vote below then sees row - row = 0, not -1, and leaves the hit alone)
                if (left == 1 && !moved(lp, ra, rb)) {
#pragma unroll
                    for (int k = 0; k < KPL; ++k) ra[k] = rb[k];
                    lp += ld;
                    left = 0;
                }
#pragma unroll
                for (int k = 0; k < KPL; ++k) prv[k] = ra[k];          // (only read when the chunk is done)
            }
            if (left == 0) break;
            uint32_t pv[KPL], cu[KPL], dk[KPL], acc = 0;
            const bool first = left == left_first;                // warp uniform
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                cu[k] = lp[32 * k];
                pv[k] = first ? carry[k] : (lp - ld)[32 * k];
                dk[k] = cu[k] + 1u - pv[k];
                if (k == KPL - 1) dk[k] &= vm[k];
                acc |= dk[k];
            }
            irr_acc |= acc;
            changed(pv, cu, dk, pos_end - (uint32_t)left);
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = cu[k];
            lp += ld;
            --left;
        }
        if (d.flags & WD_LAST) {
            if (d.flags & WD_CHR) {                 // chr-end rows after the run's last row
                bool em[KPL];
                uint32_t endv[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const uint32_t e = ORDER ? A[k] : prv[k] + (pos_end - 1u);
                    const bool valid = ORDER ? (ibase + k < C) : cvalid[k];
                    em[k] = valid && e >= d.rec_len;
                    endv[k] = min(e, 2u * d.rec_len);
                }
                emit(em, endv, SCR_ROW_CHR);
            }
            so.end(P, lane, strip_pos0, d.rec_len);
        }
        __syncwarp();                    // stage s is free again
        if (lane == 0) issue(s);
        if (++s == S) {
            s = 0;
            parity ^= 1u;
        }
    }
    if (irr_acc >> 31) P.result[MEMO_RES_IRREGULAR] = 1;
}

}  // namespace

stream_kernel_t select_wide_kernel(int n_cols, bool order, int* kpl_out) {
    // KPL = ceil(n_cols / 32) exactly: only a lane's last slot can lie past the last column
    const int kpl = (n_cols + 31) / 32;
    if (kpl_out) *kpl_out = kpl;
#define MEMO_WIDE(KK) \
    if (kpl == KK) return order ? wide_kernel<KK, true> : wide_kernel<KK, false>;
#ifdef MEMO_WIDE_ONLY            // (SASS experiments: one instantiation)
    MEMO_WIDE(MEMO_WIDE_ONLY)
#else
    MEMO_WIDE(1) MEMO_WIDE(2) MEMO_WIDE(3) MEMO_WIDE(4) MEMO_WIDE(5) MEMO_WIDE(6) MEMO_WIDE(7) MEMO_WIDE(8)
    MEMO_WIDE(9) MEMO_WIDE(10) MEMO_WIDE(11) MEMO_WIDE(12) MEMO_WIDE(13) MEMO_WIDE(14) MEMO_WIDE(15)
    MEMO_WIDE(16)
#endif
#undef MEMO_WIDE
    return nullptr;
}

}  // namespace memo
