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
// only moves where a new MEM starts, so for most rows nothing changes (one
// compare per cell), and a cell whose MEM end moves from x to y > x is a
// delete/insert in A: with i_new = #{A > y} and i_old = #{A >= x} - 1 the
// positions i_new..i_old shift down by one and y lands at i_new -- two warp
// reductions and one shuffle instead of a sort.  The index rows of the row are
// the positions where A changed.  A is sorted from scratch once per strip.
//
// Strips are handed out by an atomic counter (one per R rows).  A strip streams
// through the warp's private shared-memory ring in chunks of T rows (bulk async
// copies + mbarriers, several chunks in flight), so a warp never waits for
// another warp.  Index rows go straight to the scratch area: a warp reserves
// P.chunk rows at a time (one atomicAdd) and appends to them, so a strip's output
// is one block of consecutive scratch rows, or a short chain of blocks when it
// crosses into the warp's next chunk.  tile_scan_kernel / strip_gather_kernel
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
#define MEMO_WIDE_BOUNDS __launch_bounds__(256, MEMO_WIDE_MIN_CTAS)
#else
#define MEMO_WIDE_BOUNDS __launch_bounds__(256)
#endif
template <int KPL, bool ORDER>
__global__ void MEMO_WIDE_BOUNDS wide_kernel(const FastParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int C = P.C, ld = P.ld, S = P.stages, T = P.T, R = P.R;

    unsigned char* const wbase = smem_raw + (size_t)warp * P.warp_smem;
    uint64_t* const bars = (uint64_t*)(wbase + P.off_bars);
    TileDesc* const descs = (TileDesc*)(wbase + P.off_descs);

    // ---------------- producer state (lane 0): the chunk sequence of the warp's strips
    long long c_lo = 0, c_hi = 0;
    memo_segment_t seg;
    seg.row_begin = seg.n_rows = 0;
    seg.pos0 = seg.rec_len = seg.rec_id = seg.flags = 0;
    unsigned long long look = 0;          // next strip (fetched one strip ahead: hides the atomic)
    long long p_strip = 0, p_row = 0, p_left = 0;
    uint32_t p_pos = 0;
    bool p_first = false, p_done = false, p_lastseg = false;
    if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);

    auto issue = [&](int s) {
        if (p_done) return;
        uint64_t* bar = &bars[s];
        TileDesc d;
        d.pad = 0; d.r_hi = 0;
        if (p_left == 0) {
            p_strip = (long long)look;
            if (p_strip >= P.n_tiles) {
                d.n = 0; d.off = 0; d.pos_h = 0; d.rec_len = 0; d.flags = WD_END; d.r_lo = 0;
                descs[s] = d;
                mbar_arrive(bar);
                p_done = true;
                return;
            }
            look = atomicAdd(P.strip_counter, 1ull);
            if (p_strip < c_lo || p_strip >= c_hi) {
                int s_lo = 0, s_hi = P.n_seg - 1;
                while (s_lo < s_hi) {
                    const int mid = (s_lo + s_hi + 1) >> 1;
                    if (P.seg_tile_start[mid] <= p_strip) s_lo = mid; else s_hi = mid - 1;
                }
                c_lo = P.seg_tile_start[s_lo];
                c_hi = P.seg_tile_start[s_lo + 1];
                seg = P.segs[s_lo];
            }
            const long long t = p_strip - c_lo;
            const int primed = (seg.flags & MEMO_SEG_PRIMED) ? 1 : 0;
            const long long m = seg.n_rows - primed;                     // compare rows of the run
            long long n_cmp = m - t * R;
            if (n_cmp > R) n_cmp = R;
            if (n_cmp < 0) n_cmp = 0;
            p_row = seg.row_begin + primed + t * R - 1;                  // the strip's predecessor row
            p_left = n_cmp + 1;
            p_pos = (uint32_t)seg.pos0 + (uint32_t)(p_row - seg.row_begin);
            p_first = true;
            p_lastseg = (p_strip + 1 == c_hi) && (seg.flags & MEMO_SEG_CHR_END);
        }
        const long long n = p_left < T ? p_left : T;
        const long long start = p_row * (long long)ld * 4;
        const long long end = (p_row + n - 1) * (long long)ld * 4 + (long long)C * 4;
        const long long a0 = start & ~15ll;
        long long a1 = (end + 15) & ~15ll;
        const long long lim = P.total_bytes & ~15ll;
        if (a1 > lim) a1 = lim;
        d.n = (int)n;
        d.off = (int)((start - a0) >> 2);
        d.pos_h = p_pos;
        d.rec_len = (uint32_t)seg.rec_len;
        d.flags = (p_first ? WD_FIRST : 0) | (n == p_left ? WD_LAST : 0) | ((n == p_left && p_lastseg) ? WD_CHR : 0);
        d.r_lo = (int)p_strip;
        descs[s] = d;
        unsigned char* data = wbase + (size_t)s * P.stage_bytes;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.dap);
        // the last < 16 bytes of the buffer cannot be part of a 16-byte granular bulk copy
        for (long long b = (a1 > a0 ? a1 : a0); b < end; b += 4)
            *reinterpret_cast<uint32_t*>(data + (b - a0)) = *reinterpret_cast<const uint32_t*>(src + b);
        if (a1 > a0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(a1 - a0));
            bulk_g2s(data, src + a0, (uint32_t)(a1 - a0), bar);
        } else {
            mbar_arrive(bar);
        }
        p_row += n;
        p_pos += (uint32_t)n;
        p_left -= n;
        p_first = false;
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
    uint32_t irr_acc = 0;

    // index rows of one row: em[k] / endv[k] per slot, `p` = BED start.  Output
    // order: ORDER -> position ibase + k; else column lane + 32 k.
    auto emit = [&](const bool (&em)[KPL], const uint32_t (&endv)[KPL], uint32_t p) {
        unsigned b[KPL];
        uint32_t total = 0, rank = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            b[k] = __ballot_sync(FULL, em[k]);
            total += __popc(b[k]);
        }
        if (total == 0) return;
        uint32_t* const dst0 = so.reserve(P, total, lane);             // warp uniform
        if (dst0 != nullptr) {
            if (ORDER) {
#pragma unroll
                for (int k = 0; k < KPL; ++k) rank += __popc(b[k] & ltmask);
            }
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const uint32_t rk = ORDER ? rank : rank + __popc(b[k] & ltmask);
                if (em[k]) {
                    uint32_t* dst = dst0 + rk * 3u;
                    dst[0] = p;
                    dst[1] = endv[k];
                    dst[2] = (uint32_t)(ORDER ? ibase + k : lane + 32 * k) + 1u;
                }
                if (ORDER) rank += em[k] ? 1u : 0u; else rank += __popc(b[k]);
            }
        }
    };

    int s = 0;
    uint32_t parity = 0;
    for (;;) {
        mbar_wait(&bars[s], parity);
        const TileDesc d = descs[s];
        if (d.flags & WD_END) break;
        // the lane's column of the stage: row r, slot k at lp[r * ld + 32 * k]
        const uint32_t* lp = reinterpret_cast<const uint32_t*>(wbase + (size_t)s * P.stage_bytes) + d.off + lane;
        uint32_t pos = d.pos_h;                                  // position of the row at lp
        int left = d.n;
        if (d.flags & WD_FIRST) {
            // strip start: row 0 only primes the state
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = lp[32 * k];
            if (ORDER) {
#pragma unroll
                for (int k = 0; k < KPL; ++k) A[k] = cvalid[k] ? prv[k] + pos : 0u;
                group_sort_desc<32, KPL>(A, lane);
            }
            so.begin(d.r_lo);
            lp += ld;
            ++pos;
            --left;
        }
        // one row: `cur` = its raw values (returned), `prev` = the row before
        auto row = [&](const uint32_t (&prev)[KPL], uint32_t (&cur)[KPL]) {
            uint32_t dk[KPL], acc = 0;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                cur[k] = lp[32 * k];
                dk[k] = cur[k] + 1u - prev[k];
                acc |= k == KPL - 1 ? (dk[k] & vm[k]) : dk[k];
            }
            if (__any_sync(FULL, acc != 0u)) {
                irr_acc |= acc;
                bool em[KPL];
                uint32_t endv[KPL];
                if (ORDER) {
                    // cells whose MEM end moved up (a decrease makes the input irregular:
                    // flagged through irr_acc, skipped here)
                    bool ch[KPL];
                    uint32_t cnt = 0;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        ch[k] = (int)(k == KPL - 1 ? (dk[k] & vm[k]) : dk[k]) > 0;
                        cnt += ch[k] ? 1u : 0u;
                    }
                    const unsigned ball = __ballot_sync(FULL, cnt != 0u);
                    const uint32_t nchg = __reduce_add_sync(FULL, cnt);
                    if (nchg == 1) {
                        // the common case, one cell x -> y: the positions holding x <= A <= y
                        // shift down by one, y lands on the first of them, and exactly those
                        // positions can emit
                        uint32_t myx = prev[0], myd = dk[0];
#pragma unroll
                        for (int k = 1; k < KPL; ++k)
                            if (ch[k]) { myx = prev[k]; myd = dk[k]; }
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
                            const bool inr = old >= x && old <= y;
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
                                    A[kk] = (A[kk] >= x && A[kk] <= y) ? min(before, y) : A[kk];
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
                        em[k] = (int)(dk[k] & vm[k]) > 0 && e >= pos;
                        endv[k] = e;
                    }
                }
                emit(em, endv, pos);
            }
            lp += ld;
            ++pos;
        };
        // rows alternate between two register sets (no copies); prv is the set that
        // holds the last row seen when a chunk ends
        uint32_t alt[KPL];
        for (; left >= 2; left -= 2) {
            row(prv, alt);
            row(alt, prv);
        }
        if (left) {
            row(prv, alt);
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = alt[k];
        }
        if (d.flags & WD_LAST) {
            if (d.flags & WD_CHR) {                 // chr-end rows after the run's last row
                bool em[KPL];
                uint32_t endv[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const uint32_t e = ORDER ? A[k] : prv[k] + (pos - 1u);
                    const bool valid = ORDER ? (ibase + k < C) : cvalid[k];
                    em[k] = valid && e >= d.rec_len;
                    endv[k] = min(e, 2u * d.rec_len);
                }
                emit(em, endv, d.rec_len);
            }
            so.end(P, lane);
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
    MEMO_WIDE(1) MEMO_WIDE(2) MEMO_WIDE(3) MEMO_WIDE(4) MEMO_WIDE(5) MEMO_WIDE(6) MEMO_WIDE(7) MEMO_WIDE(8)
    MEMO_WIDE(9) MEMO_WIDE(10) MEMO_WIDE(11) MEMO_WIDE(12) MEMO_WIDE(13) MEMO_WIDE(14) MEMO_WIDE(15)
    MEMO_WIDE(16)
#undef MEMO_WIDE
    return nullptr;
}

}  // namespace memo
