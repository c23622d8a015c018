// DAP -> MEMO index rows on sm_100a: single-pass build for NARROW rows
// (n_cols == ld == CT known at compile time, CT <= 16 -- the 5- and 10-genome
// configurations), one lane per DAP row.
//
// Replaces the hot loop of the reference's src/dap_to_bed.py (--mem --overlap
// [--order]): get_new_record :85-91, dap_to_mem :116-134, print_interval /
// overlaps :93-109.  Same mathematics as index_build.cu (DESIGN.md "index
// build"): with E[r][c] = p(r) + v[r][c] and A[r] = E[r] sorted descending
// (--order) or E[r] itself, row r emits (p, A[r-1][j], j+1) for every j with
// A[r][j] > A[r-1][j] and A[r-1][j] >= p, provided no E decreases down a column
// (matching statistics).  Only rows in which some E moved can emit.
//
// Every warp is an independent stream: it takes strips of G consecutive tiles
// from an atomic counter (tile = T consecutive rows + the predecessor row, fetched
// by one bulk async copy into the warp's private multi-stage shared-memory ring).
// Tile bases are multiples of RPL rows of the buffer, which makes them 16-byte
// aligned, so that
//   phase A  a lane scans RPL consecutive rows with 128-bit shared-memory loads
//            (lane stride RPL*CT*4 bytes = an odd number of 16-byte units: no
//            bank conflicts) and flags the rows where v[r][c] != v[r-1][c] - 1,
//   phase B  the flagged rows (a few per tile) are queued in shared memory as
//            (start, limit, MEM ends of the predecessor row, MEM ends of the row);
//            whenever 32 are waiting -- or the strip ends -- one lane per queued
//            row sorts both with a register sorting network (--order), compares
//            them position by position and stores the index rows straight into
//            the scratch area, appended to the strip's output (StripOut: the warp
//            reserves scratch rows in chunks; tile_scan_kernel /
//            strip_gather_kernel of index_build.cu copy the strips' blocks into
//            the ordered output).  The sorting network thus always runs on a full
//            warp, not on the handful of rows one tile flags.
// No warp ever waits for another.
#include "index_fast.cuh"

namespace memo {
namespace {

constexpr int NARROW_QCAP = 40;        // queued flagged rows per warp (>= 32 + what one more append leaves)

__host__ __device__ constexpr int gcd4(int c) { return (c % 4 == 0) ? 4 : ((c % 2 == 0) ? 2 : 1); }

// Batcher odd-even merge sort network on P = 2^ceil(log2 N) wires with every
// comparator touching a wire >= N removed (the missing wires would hold -inf and
// sort to the end of a descending order anyway).
template <int N>
__device__ __forceinline__ void sort_desc_network(uint32_t (&a)[N]) {
    constexpr int P = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : N <= 8 ? 8 : 16;
#pragma unroll
    for (int p = 1; p < P; p <<= 1) {
#pragma unroll
        for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = k % p; j + k < P; j += 2 * k) {
#pragma unroll
                for (int i = 0; i < k; ++i) {
                    const int x = i + j, y = i + j + k;
                    if (y < N && (x / (2 * p)) == (y / (2 * p))) {
                        const uint32_t hi = max(a[x], a[y]);
                        const uint32_t lo = min(a[x], a[y]);
                        a[x] = hi;
                        a[y] = lo;
                    }
                }
            }
        }
    }
}

template <int CT, bool ORDER>
__global__ void __launch_bounds__(256) narrow_kernel(const FastParams P) {
    constexpr int RPL = 4 / gcd4(CT);             // rows per lane per step
    constexpr int RI = 32 * RPL;                  // rows per warp step
    constexpr int NW = (RPL + 1) * CT;            // words a lane needs per step
    constexpr int NV = (NW + 3) / 4;              // ... as 128-bit loads
    static_assert(CT >= 1 && CT <= 16, "narrow rows only");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int S = P.stages, T = P.T;

    unsigned char* const wbase = smem_raw + (size_t)warp * P.warp_smem;
    uint64_t* const bars = (uint64_t*)(wbase + P.off_bars);
    TileDesc* const descs = (TileDesc*)(wbase + P.off_descs);
    uint16_t* const list = (uint16_t*)(wbase + P.off_list);

    // ---------------- producer state (lane 0): the tile sequence of the warp's strips
    const int G = P.R;                    // tiles per strip
    long long c_lo = 0, c_hi = 0;         // strips of the cached record run
    memo_segment_t seg;
    seg.row_begin = seg.n_rows = 0;
    seg.pos0 = seg.rec_len = seg.rec_id = seg.flags = 0;
    unsigned long long look = 0;          // next strip (fetched one strip ahead: hides the atomic)
    long long p_strip = 0, p_tile = 0, p_left = 0, p_ntiles = 0, p_fc = 0, p_lc = 0, p_B = 0;
    bool p_first = false, p_done = false;
    if (lane == 0) look = atomicAdd(P.strip_counter, 1ull);

    // fetch the next tile into stage s (lane 0 only)
    auto issue = [&](int s) {
        if (p_done) return;
        uint64_t* bar = &bars[s];
        TileDesc d;
        if (p_left == 0) {
            p_strip = (long long)look;
            if (p_strip >= P.n_tiles) {
                d.n = 0; d.off = 0; d.pos_h = 0; d.rec_len = 0; d.flags = WD_END; d.r_lo = 1; d.r_hi = 0; d.pad = 0;
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
            // tiles of the run: see make_fast_plan
            const int primed = (seg.flags & MEMO_SEG_PRIMED) ? 1 : 0;
            p_fc = seg.row_begin + primed;                        // first / last compare row (buffer rows)
            p_lc = seg.row_begin + seg.n_rows - 1;
            const long long g = ((p_fc - 1) / RPL) * RPL;         // aligned base of the run's first tile
            long long nt = p_lc >= p_fc ? (p_lc - g + T - 1) / T : 0;
            if (nt < 1) nt = 1;
            p_ntiles = nt;
            p_tile = (p_strip - c_lo) * G;
            p_left = nt - p_tile < G ? nt - p_tile : G;
            p_B = g + p_tile * T;                                 // buffer row of the next tile's row 0
            p_first = true;
        }
        const long long B = p_B;
        const long long lo = (p_fc > B + 1 ? p_fc : B + 1) - B;
        const long long hi = (p_lc < B + T ? p_lc : B + T) - B;
        const long long lc = p_lc;
        const long long a0 = B * (long long)(CT * 4);             // 16-byte aligned
        long long end = a0 + (long long)(T + 1) * (CT * 4);
        if (end > P.total_bytes) end = P.total_bytes;
        long long a1 = (end + 15) & ~15ll;
        const long long lim = P.total_bytes & ~15ll;
        if (a1 > lim) a1 = lim;
        const bool runlast = p_tile + 1 == p_ntiles;
        d.n = T;
        d.off = (int)p_strip;
        d.pos_h = (uint32_t)seg.pos0 + (uint32_t)(B - seg.row_begin);
        d.rec_len = (uint32_t)seg.rec_len;
        d.flags = (p_first ? WD_FIRST : 0) | (p_left == 1 ? WD_LAST : 0) | (runlast ? WD_RUNLAST : 0) |
                  ((runlast && (seg.flags & MEMO_SEG_CHR_END)) ? WD_CHR : 0);
        d.r_lo = (int)lo;
        d.r_hi = (int)hi;
        d.pad = (int)(lc - B);                                    // tile row of the run's last row
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
        ++p_tile;
        --p_left;
        p_B += T;
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

    uint32_t irr_acc = 0;
    int s = 0;
    uint32_t parity = 0;
    StripOut so;                                      // the strip's output blocks
    uint32_t strip_pos0 = 0;                          // position of row 0 of the strip's first tile
    // queue of flagged rows (ring): entry = {start, limit, e[CT], f[CT]}, odd stride
    constexpr int ESTR = (2 * CT + 2) | 1;
    constexpr int QCAP = NARROW_QCAP;
    uint32_t* const queue = reinterpret_cast<uint32_t*>(wbase + P.off_stg);
    int qhead = 0, qn = 0;

    // phase B on the first n (<= 32) queued rows: one lane per row
    auto flush = [&](int n) {
        uint32_t e[CT], p = 0, lim = 0;
        unsigned m = 0;
        int slot = qhead + lane;
        if (slot >= QCAP) slot -= QCAP;
        uint32_t* const q = queue + slot * ESTR;
        if (lane < n) {
            uint32_t f[CT];
            p = q[0];
            lim = q[1];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                e[c] = q[2 + c];
                f[c] = q[2 + CT + c];
            }
            if (ORDER) {
                sort_desc_network<CT>(e);
                sort_desc_network<CT>(f);
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) m |= (f[c] > e[c] && e[c] >= p) ? (1u << c) : 0u;
            // the (sorted) ends go back into the entry: the store loop below then runs
            // once per emitted row, not once per column
#pragma unroll
            for (int c = 0; c < CT; ++c) q[2 + c] = min(e[c], lim);
        }
        const uint32_t cnt = __popc(m);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (total) {
            uint32_t* dst = so.reserve(P, total, lane);
            unsigned mm = dst ? m : 0u;
            dst += (incl - cnt) * SCR_WORDS;
            // (all queued rows belong to the open strip: the queue is emptied when a strip ends)
            const uint32_t rowf = lim != 0xFFFFFFFFu ? SCR_ROW_CHR : p - strip_pos0;
            while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                scr_store(dst, q[2 + c], rowf, (uint32_t)(c + 1));
                dst += SCR_WORDS;
            }
        }
        __syncwarp();
        qhead += n;
        if (qhead >= QCAP) qhead -= QCAP;
        qn -= n;
    };

    for (;;) {
        mbar_wait(&bars[s], parity);
        const TileDesc d = descs[s];
        if (d.flags & WD_END) break;
        if (d.flags & WD_FIRST) {
            so.begin(d.off);
            strip_pos0 = d.pos_h;
        }
        const uint32_t* const sdata = reinterpret_cast<const uint32_t*>(wbase + (size_t)s * P.stage_bytes);

        // ---------------- phase A: ordered list of the live rows that moved a MEM end
        int n_ch = 0;
        for (int r0 = 0; r0 < T; r0 += RI) {
            const int row0 = r0 + lane * RPL;                 // predecessor of the lane's first row
            uint32_t acc[RPL];
            if (row0 < d.r_hi && row0 + RPL >= d.r_lo) {
                const uint4* src = reinterpret_cast<const uint4*>(sdata + row0 * CT);
                uint32_t w[NV * 4];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const uint4 x = src[v];
                    w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
                }
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    uint32_t a = 0;
#pragma unroll
                    for (int c = 0; c < CT; ++c) a |= w[(q + 1) * CT + c] + 1u - w[q * CT + c];
                    const int row = row0 + 1 + q;
                    acc[q] = (row >= d.r_lo && row <= d.r_hi) ? a : 0u;
                }
            } else {
#pragma unroll
                for (int q = 0; q < RPL; ++q) acc[q] = 0u;
            }
            unsigned bal[RPL];
            unsigned any = 0;
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                irr_acc |= acc[q];
                bal[q] = __ballot_sync(FULL, acc[q] != 0u);
                any |= bal[q];
            }
            if (any) {
                // list order = row order = (lane, q) lexicographic
                int o = n_ch, total = 0;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    o += __popc(bal[q] & ltmask);
                    total += __popc(bal[q]);
                }
#pragma unroll
                for (int q = 0; q < RPL; ++q)
                    if (acc[q] != 0u) list[o++] = (uint16_t)(row0 + 1 + q);
                n_ch += total;
            }
        }
        if (d.flags & WD_CHR) {                         // the chr-end rows after the run's last row
            if (lane == 0) list[n_ch] = (uint16_t)(0x8000 | (d.pad + 1));
            ++n_ch;
        }
        __syncwarp();

        // ---------------- queue the listed rows; phase B whenever a full warp of them waits
        for (int done = 0; done < n_ch;) {
            int m = n_ch - done;
            if (m > 32) m = 32;
            if (m > QCAP - qn) m = QCAP - qn;
            if (lane < m) {
                const unsigned ent = list[done + lane];
                const bool chr = (ent & 0x8000u) != 0u;
                const int row = (int)(ent & 0x7FFFu);
                const uint32_t* prevp = sdata + (row - 1) * CT;
                const uint32_t ppos = d.pos_h + (uint32_t)(row - 1);
                int slot = qhead + qn + lane;
                if (slot >= QCAP) slot -= QCAP;
                uint32_t* q = queue + slot * ESTR;
                q[0] = chr ? d.rec_len : ppos + 1u;
                q[1] = chr ? 2u * d.rec_len : 0xFFFFFFFFu;
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    q[2 + c] = prevp[c] + ppos;
                    q[2 + CT + c] = chr ? 0xFFFFFFFFu : prevp[CT + c] + ppos + 1u;
                }
            }
            qn += m;
            done += m;
            __syncwarp();
            if (qn >= 32) flush(32);
        }
        if (d.flags & WD_LAST) {
            while (qn > 0) flush(qn < 32 ? qn : 32);
        }
        if (d.flags & WD_LAST) so.end(P, lane, strip_pos0, d.rec_len);
        __syncwarp();                    // stage s and the list are free again
        if (lane == 0) issue(s);
        if (++s == S) {
            s = 0;
            parity ^= 1u;
        }
    }
    if (irr_acc >> 31) P.result[MEMO_RES_IRREGULAR] = 1;
}

}  // namespace

stream_kernel_t select_narrow_kernel(int n_cols, bool order, int* rows_per_lane) {
#define MEMO_NARROW(CC)                                                            \
    if (n_cols == CC) {                                                            \
        if (rows_per_lane) *rows_per_lane = 4 / gcd4(CC);                          \
        return order ? narrow_kernel<CC, true> : narrow_kernel<CC, false>;         \
    }
    MEMO_NARROW(1) MEMO_NARROW(2) MEMO_NARROW(3) MEMO_NARROW(4) MEMO_NARROW(5) MEMO_NARROW(6)
    MEMO_NARROW(7) MEMO_NARROW(8) MEMO_NARROW(9) MEMO_NARROW(10) MEMO_NARROW(11) MEMO_NARROW(12)
    MEMO_NARROW(13) MEMO_NARROW(14) MEMO_NARROW(15) MEMO_NARROW(16)
#undef MEMO_NARROW
    return nullptr;
}

}  // namespace memo
