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
// P.chunk rows at a time (one atomicAdd) and appends to them; the blocks a strip
// wrote are recorded in the strip's slots of tile_cnt / tile_off, and
// tile_scan_kernel / tile_gather_kernel (index_build.cu) copy the blocks into
// the ordered output.
#include "index_fast.cuh"
#include "warp_sort.cuh"

namespace memo {
namespace {

constexpr int WD_FIRST = 1;    // chunk row 0 is the strip's predecessor row
constexpr int WD_LAST = 2;     // last chunk of the strip
constexpr int WD_CHR = 4;      // chr-end rows follow the strip
constexpr int WD_END = 8;      // no more strips

template <int KPL, bool ORDER>
__global__ void __launch_bounds__(256) wide_kernel(const FastParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int C = P.C, ld = P.ld, S = P.stages, T = P.T, R = P.R, MAXB = P.maxb;

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
    // raw columns: slot k of lane l is DAP column l + 32 k (clamped: lanes past the
    // last column repeat it and are masked out of every vote)
    int col[KPL];
    unsigned vmask[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        const int c = lane + 32 * k;
        col[k] = c < C ? c : C - 1;
        vmask[k] = __ballot_sync(FULL, c < C);
    }
    uint32_t prv[KPL];                 // previous row, raw values
    uint32_t A[KPL];                   // ORDER: sorted MEM ends of the previous row, position lane*KPL + k
#pragma unroll
    for (int k = 0; k < KPL; ++k) prv[k] = A[k] = 0;
    unsigned long long w_cur = 0, w_end = 0;       // the warp's reserved scratch rows
    unsigned long long blk_off = 0, my_off = 0;    // open block; lane b keeps closed block b of the strip
    uint32_t blk_cnt = 0, my_cnt = 0;
    int nb = 0;
    uint32_t irr_acc = 0;

    // append `total` index rows (warp uniform, <= n_cols) to the strip's output;
    // returns the scratch row of the first one
    auto reserve = [&](uint32_t total) -> unsigned long long {
        if (total > w_end - w_cur) {               // next chunk: the strip continues in a new block
            if (lane == nb) { my_off = blk_off; my_cnt = blk_cnt; }
            ++nb;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.cursor, (unsigned long long)P.chunk);
            base = __shfl_sync(FULL, base, 0);
            w_cur = base;
            w_end = base + P.chunk;
            blk_off = base;
            blk_cnt = 0;
        }
        const unsigned long long at = w_cur;
        w_cur += total;
        blk_cnt += total;
        return at;
    };

    // index rows of one row: em[k] / end[k] per slot, `p` = BED start
    auto emit = [&](const bool (&em)[KPL], const uint32_t (&endv)[KPL], uint32_t p) {
        unsigned b[KPL];
        uint32_t total = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            b[k] = __ballot_sync(FULL, em[k]);
            total += __popc(b[k]);
        }
        if (total == 0) return;
        const unsigned long long at = reserve(total);
        // rank in output order: ORDER -> position lane*KPL + k; else column lane + 32 k
        uint32_t rank = 0;
        if (ORDER) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) rank += __popc(b[k] & ltmask);
        }
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const uint32_t rk = ORDER ? rank : rank + __popc(b[k] & ltmask);
            if (em[k]) {
                const unsigned long long gi = at + rk;
                if (gi < (unsigned long long)P.scr_cap) {
                    P.scr_start[gi] = p;
                    P.scr_end[gi] = endv[k];
                    P.scr_order[gi] = (uint32_t)(ORDER ? lane * KPL + k : lane + 32 * k) + 1u;
                }
            }
            if (ORDER) rank += em[k] ? 1u : 0u; else rank += __popc(b[k]);
        }
    };

    int s = 0;
    uint32_t parity = 0;
    for (;;) {
        mbar_wait(&bars[s], parity);
        const TileDesc d = descs[s];
        if (d.flags & WD_END) break;
        const uint32_t* const sdata = reinterpret_cast<const uint32_t*>(wbase + (size_t)s * P.stage_bytes) + d.off;
        int r = 0;
        if (d.flags & WD_FIRST) {
            // strip start: row 0 only primes the state
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = sdata[col[k]];
            if (ORDER) {
#pragma unroll
                for (int k = 0; k < KPL; ++k) A[k] = (lane + 32 * k < C) ? prv[k] + d.pos_h : 0u;
                group_sort_desc<32, KPL>(A, lane);
            }
            nb = 0;
            blk_off = w_cur;
            blk_cnt = 0;
            r = 1;
        }
        for (; r < d.n; ++r) {
            const uint32_t pos = d.pos_h + (uint32_t)r;          // position of this row
            const uint32_t* rowp = sdata + r * ld;
            uint32_t cur[KPL], dk[KPL], acc = 0;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                cur[k] = rowp[col[k]];
                dk[k] = cur[k] + 1u - prv[k];
                acc |= dk[k];
            }
            if (__any_sync(FULL, acc != 0u)) {
                irr_acc |= acc;
                bool em[KPL];
                uint32_t endv[KPL];
                if (ORDER) {
                    uint32_t Aold[KPL];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) Aold[k] = A[k];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        // cells whose MEM end moved up (a decrease makes the input irregular)
                        unsigned m = __ballot_sync(FULL, dk[k] != 0u && !(dk[k] >> 31)) & vmask[k];
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t x = __shfl_sync(FULL, prv[k], src) + (pos - 1u);
                            const uint32_t y = __shfl_sync(FULL, cur[k], src) + pos;
                            uint32_t cgt = 0, cge = 0;
#pragma unroll
                            for (int kk = 0; kk < KPL; ++kk) {
                                cgt += A[kk] > y ? 1u : 0u;
                                cge += A[kk] >= x ? 1u : 0u;
                            }
                            const int i_new = (int)__reduce_add_sync(FULL, cgt);
                            const int i_old = (int)__reduce_add_sync(FULL, cge) - 1;
                            const uint32_t up = __shfl_up_sync(FULL, A[KPL - 1], 1);
                            uint32_t nA[KPL];
#pragma unroll
                            for (int kk = 0; kk < KPL; ++kk) {
                                const int i = lane * KPL + kk;
                                const uint32_t before = kk == 0 ? up : A[kk - 1];
                                nA[kk] = (i < i_new || i > i_old) ? A[kk] : (i == i_new ? y : before);
                            }
#pragma unroll
                            for (int kk = 0; kk < KPL; ++kk) A[kk] = nA[kk];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        em[k] = (lane * KPL + k < C) && A[k] > Aold[k] && Aold[k] >= pos;
                        endv[k] = Aold[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const uint32_t e = prv[k] + (pos - 1u);
                        em[k] = (lane + 32 * k < C) && dk[k] != 0u && !(dk[k] >> 31) && e >= pos;
                        endv[k] = e;
                    }
                }
                emit(em, endv, pos);
            }
#pragma unroll
            for (int k = 0; k < KPL; ++k) prv[k] = cur[k];
        }
        if (d.flags & WD_LAST) {
            if (d.flags & WD_CHR) {                 // chr-end rows after the run's last row
                const uint32_t last_pos = d.pos_h + (uint32_t)(d.n - 1);
                bool em[KPL];
                uint32_t endv[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const uint32_t e = ORDER ? A[k] : prv[k] + last_pos;
                    const bool valid = ORDER ? (lane * KPL + k < C) : (lane + 32 * k < C);
                    em[k] = valid && e >= d.rec_len;
                    endv[k] = min(e, 2u * d.rec_len);
                }
                emit(em, endv, d.rec_len);
            }
            // the strip's blocks -> its slots
            if (lane == nb) { my_off = blk_off; my_cnt = blk_cnt; }
            if (lane < MAXB) {
                const long long slot = (long long)d.r_lo * MAXB + lane;
                P.tile_cnt[slot] = lane <= nb ? my_cnt : 0u;
                P.tile_off[slot] = my_off;
            }
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
    static const int kpls[] = {1, 2, 3, 4, 6, 8, 16};
    for (int kpl : kpls) {
        if (32 * kpl < n_cols) continue;
        if (kpl_out) *kpl_out = kpl;
#define MEMO_WIDE(KK) \
    if (kpl == KK) return order ? wide_kernel<KK, true> : wide_kernel<KK, false>;
        MEMO_WIDE(1) MEMO_WIDE(2) MEMO_WIDE(3) MEMO_WIDE(4) MEMO_WIDE(6) MEMO_WIDE(8) MEMO_WIDE(16)
#undef MEMO_WIDE
    }
    return nullptr;
}

}  // namespace memo
