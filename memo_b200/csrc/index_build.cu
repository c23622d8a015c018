// DAP -> MEMO index rows on sm_100a: the single-pass build for matching statistics.
//
// Replaces the hot loop of the reference's src/dap_to_bed.py (--mem --overlap
// [--order]): get_new_record :85-91, dap_to_mem :116-134, print_interval /
// overlaps :93-109.  DESIGN.md "index build" has the derivation; summary:
//
//   E[r][c] = p(r) + v[r][c]                      ("MEM end" of DAP cell (r, c))
//   A[r]    = E[r] sorted descending (--order) or E[r] itself (membership)
//   row r emits (p, A[r-1][j], j+1) for every j with A[r][j] > A[r-1][j] and
//   A[r-1][j] >= p -- provided no E ever decreases down a column, which holds
//   for matching statistics (v[r] >= v[r-1] - 1).  Then "the previous MEM of
//   column j" (the dict of :107) is always A[r-1][j], a row depends on nothing
//   but its predecessor, and rows with E[r] == E[r-1] (the vast majority: E
//   only moves where a new MEM starts) emit nothing.
//
// In --order mode no sort is run.  With G(v) = #{c : E[r-1][c] > v} and
// D(v) = #{c : E[r-1][c] <= v < E[r][c]} (columns whose MEM end crossed v), the
// sorted positions that change are, for every value v of the previous row with
// D(v) > 0, the first min(D(v), multiplicity(v)) positions holding v:
// j = G(v) + t.  A changed row costs counting proportional to the index rows it
// emits, not a sort.
//
// Three kernels:
//  1 stream_kernel  every warp is an independent stream over its own tiles
//      (tile = T consecutive rows of one record run + the predecessor row,
//      tiles dealt round-robin to warps).  A tile is fetched into the warp's
//      private shared-memory stage by one bulk async copy (TMA: cp.async.bulk +
//      mbarrier, multi-stage), then
//        phase A  flat 128-bit scan for cells with v[r][c] != v[r-1][c] - 1
//                 -> bitmap of changed rows (and the "irregular" verdict when a
//                 MEM end decreases),
//        phase B  groups of G lanes turn the changed rows into index rows,
//                 staged in shared memory,
//      and the tile's rows are appended to a scratch area at a block obtained by
//      one atomicAdd (unordered, exactly sized).  No warp ever waits for another.
//  2 tile_scan_kernel   block sums of the per-tile row counts; the last block to
//      finish scans the block sums.
//  3 tile_gather_kernel exclusive scan inside each block of tiles and copy of
//      every tile's rows from its scratch block to its place in the ordered
//      output (the extra traffic is 24 B per index row, a few % of the DAP).
// If the input is irregular the result must be discarded and the general build
// (index_general.cu) run instead; memo_index_build reports that in
// result[MEMO_RES_IRREGULAR].
#include "index_fast.cuh"

namespace memo {
namespace {

// ---------------------------------------------------------------- kernel 1
// KPL = DAP columns per lane in phase B (1: a row is handled by a group of
// P.gw lanes and several rows share a warp; > 1: one warp per row).
// CT = compile-time number of DAP columns with ld == CT and one lane per row in
// phase A (0: generic): every loop over columns unrolls to immediate offsets.
template <int KPL, bool ORDER, int CT>
__global__ void __launch_bounds__(256) stream_kernel(const FastParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ltmask = (1u << lane) - 1u;
    const int C = CT ? CT : P.C, ld = CT ? CT : P.ld, K = P.K, S = P.stages;
    // phase B geometry: groups of GW lanes, RP rows per pass
    const int GW = (KPL == 1) ? (CT ? (CT <= 16 ? CT : 32) : P.gw) : 32;
    const int RP = 32 / GW;
    const int g = lane / GW;
    const int lg = lane - g * GW;
    const bool glive = g < RP;
    const unsigned gmask = !glive ? 0u : (GW == 32 ? FULL : (((1u << (GW & 31)) - 1u) << (g * GW)));
    // phase A geometry: a row is scanned by SL lanes (SL a power of two), W columns each
    const int SL = CT ? 1 : P.sl, W = CT ? CT : P.wcols;
    const int RL = 32 / SL;
    const int ar = lane / SL, sg = lane & (SL - 1);
    const int a_c0 = sg * W;
    const int a_cw = min(W, C - a_c0);

    // the warp's private shared memory
    unsigned char* const wbase = smem_raw + (size_t)warp * P.warp_smem;
    uint64_t* const bars = (uint64_t*)(wbase + P.off_bars);
    TileDesc* const descs = (TileDesc*)(wbase + P.off_descs);
    uint32_t* const stg = (uint32_t*)(wbase + P.off_stg) + (size_t)(glive ? g : 0) * K * 3;
    uint16_t* const list = (uint16_t*)(wbase + P.off_list);

    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long w_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;

    // fetch tile `tile` into stage s (lane 0 only); the record run of the previous
    // tile is kept in registers, consecutive tiles of a warp mostly share it
    long long c_lo = 0, c_hi = 0;
    memo_segment_t seg;
    seg.row_begin = seg.n_rows = 0;
    seg.pos0 = seg.rec_len = seg.rec_id = seg.flags = 0;
    auto issue = [&](int s, long long tile) {
        uint64_t* bar = &bars[s];
        if (tile >= P.n_tiles) return;
        if (tile < c_lo || tile >= c_hi) {
            int s_lo = 0, s_hi = P.n_seg - 1;
            while (s_lo < s_hi) {
                const int mid = (s_lo + s_hi + 1) >> 1;
                if (P.seg_tile_start[mid] <= tile) s_lo = mid; else s_hi = mid - 1;
            }
            c_lo = P.seg_tile_start[s_lo];
            c_hi = P.seg_tile_start[s_lo + 1];
            seg = P.segs[s_lo];
        }
        const long long t = tile - c_lo;
        const int primed = (seg.flags & MEMO_SEG_PRIMED) ? 1 : 0;
        const long long m = seg.n_rows - primed;                    // compare rows of the run
        const long long h = seg.row_begin - (1 - primed) + t * P.T; // buffer row of the tile's row 0
        long long n = m - t * P.T;
        if (n > P.T) n = P.T;
        if (n < 0) n = 0;
        const long long start = h * (long long)ld * 4;
        const long long end = (h + n) * (long long)ld * 4 + (long long)C * 4;
        const long long a0 = start & ~15ll;
        long long a1 = (end + 15) & ~15ll;
        const long long lim = P.total_bytes & ~15ll;
        if (a1 > lim) a1 = lim;
        TileDesc d;
        d.n = (int)n;
        d.off = (int)((start - a0) >> 2);
        d.pos_h = (uint32_t)(seg.pos0 - (1 - primed)) + (uint32_t)(t * P.T);
        d.rec_len = (uint32_t)seg.rec_len;
        d.flags = ((tile + 1 == c_hi) ? 1 : 0) | ((seg.flags & MEMO_SEG_CHR_END) ? 2 : 0);
        d.r_lo = 1; d.r_hi = (int)n; d.pad = 0;
        descs[s] = d;
        unsigned char* data = wbase + (size_t)s * P.stage_bytes + 16;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.dap);
        // the last < 16 bytes of the buffer cannot be part of a 16-byte granular bulk copy
        for (long long b = a1; b < end; b += 4)
            *reinterpret_cast<uint32_t*>(data + (b - a0)) = *reinterpret_cast<const uint32_t*>(src + b);
        if (a1 > a0) {
            mbar_arrive_expect_tx(bar, (uint32_t)(a1 - a0));
            bulk_g2s(data, src + a0, (uint32_t)(a1 - a0), bar);
        } else {
            mbar_arrive(bar);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
        for (int s = 0; s < S; ++s) issue(s, w_global + (long long)s * n_warps);

    uint32_t irr_acc = 0;
    unsigned long long replays = 0;
    int s = 0;
    uint32_t parity = 0;

    for (long long tile = w_global; tile < P.n_tiles; tile += n_warps) {
        mbar_wait(&bars[s], parity);
        const TileDesc d = descs[s];
        const int n = d.n, off = d.off;
        const uint32_t* const sdata = reinterpret_cast<const uint32_t*>(wbase + (size_t)s * P.stage_bytes + 16);

        // ---------------- phase A: ordered list of the rows that moved a MEM end
        // (v[r][c] + 1 - v[r-1][c] != 0 somewhere in the row).  32 / SL rows per step,
        // SL lanes per row with W columns each.
        int n_ch = 0;
        for (int r0 = 1; r0 <= n; r0 += RL) {
            const int row = r0 + ar;
            uint32_t acc = 0;
            if (row <= n && a_cw > 0) {
                const uint32_t* cur = sdata + off + row * ld + a_c0;
                const uint32_t* prv = cur - ld;
                if (CT) {
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc |= cur[c] + 1u - prv[c];
                } else {
#pragma unroll 4
                    for (int c = 0; c < a_cw; ++c) acc |= cur[c] + 1u - prv[c];
                }
            }
            for (int o = 1; o < SL; o <<= 1) acc |= __shfl_xor_sync(FULL, acc, o);
            irr_acc |= acc;
            const bool mine = acc != 0u && sg == 0;
            const unsigned bal = __ballot_sync(FULL, mine);
            if (bal) {
                if (mine) list[n_ch + __popc(bal & ltmask)] = (uint16_t)row;
                n_ch += __popc(bal);
            }
        }
        if ((d.flags & 3) == 3) {                       // pseudo row n + 1: the chr-end rows
            if (lane == 0) list[n_ch] = (uint16_t)(n + 1);
            ++n_ch;
        }
        __syncwarp();
        const int chunk = (n_ch + RP - 1) / RP;

        // ---------------- phase B: index rows of the changed rows
        uint32_t gcount = 0;
        unsigned long long gbase = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const bool direct = pass == 1;
            gcount = 0;
            for (int ci = 0; ci < chunk; ++ci) {
                const int idx = g * chunk + ci;
                const bool act = glive && idx < n_ch;
                const int row = act ? (int)list[idx] : 1;
                const bool chr = act && row == n + 1;
                const uint32_t* prevp = sdata + off + (row - 1) * ld;
                const uint32_t ppos = d.pos_h + (uint32_t)(row - 1);
                const uint32_t p = chr ? d.rec_len : ppos + 1u;
                const uint32_t lim = chr ? 2u * d.rec_len : 0xFFFFFFFFu;
                uint32_t e[KPL], f[KPL];
                bool valid[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const int c = k * GW + lg;
                    valid[k] = c < C && glive;
                    e[k] = valid[k] ? prevp[c] + ppos : 0u;
                    f[k] = valid[k] ? (chr ? 0xFFFFFFFFu : prevp[ld + c] + ppos + 1u) : 0u;
                }
                bool em[KPL];           // slot emits an index row
                uint32_t jpos[KPL];     // its 0-based order / genome column
                uint32_t rank[KPL];     // its rank among the row's index rows
                uint32_t total = 0;

                if (!ORDER) {
                    unsigned bal[KPL];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        em[k] = act && valid[k] && f[k] > e[k] && e[k] >= p;
                        jpos[k] = (uint32_t)(k * GW + lg);
                        bal[k] = __ballot_sync(FULL, em[k]) & gmask;
                    }
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        rank[k] = total + __popc(bal[k] & ltmask);
                        total += __popc(bal[k]);
                    }
                } else if (KPL == 1 && (CT ? CT <= 12 : P.all_pairs)) {
                    // narrow rows.  D(v): one step per changed column of the row (lockstep
                    // over the warp's groups); tie index t from a match on (group, v);
                    // G(v) for every lane at once by all pairs within the group.
                    uint32_t dd = chr ? 0x7FFFFFFFu : 0u;
                    unsigned m = __ballot_sync(FULL, act && !chr && f[0] != e[0]) & gmask;
                    while (__any_sync(FULL, m != 0u)) {
                        const bool on = m != 0u;
                        const int src = on ? __ffs(m) - 1 : lane;
                        m &= m - 1;
                        const uint32_t x = __shfl_sync(FULL, e[0], src);
                        const uint32_t y = __shfl_sync(FULL, f[0], src);
                        dd += (on && e[0] >= x && e[0] < y) ? 1u : 0u;
                    }
                    const bool cand = act && valid[0] && dd > 0u && e[0] >= p;
                    uint32_t Gt = 0, tt = 0;
                    if (__any_sync(FULL, cand)) {
                        const int src0 = glive ? g * GW : 0;
#pragma unroll
                        for (int kk = 0; kk < (CT ? CT : 1); ++kk) {
                            if (CT == 0) break;
                            const uint32_t ek = __shfl_sync(FULL, e[0], src0 + kk);
                            Gt += (ek > e[0]) ? 1u : 0u;
                            tt += (ek == e[0] && kk < lg) ? 1u : 0u;
                        }
                        if (CT == 0) {
                            for (int kk = 0; kk < C; ++kk) {
                                const uint32_t ek = __shfl_sync(FULL, e[0], src0 + kk);
                                Gt += (ek > e[0]) ? 1u : 0u;
                                tt += (ek == e[0] && kk < lg) ? 1u : 0u;
                            }
                        }
                    }
                    em[0] = cand && tt < dd;
                    jpos[0] = Gt + tt;
                    const int sh = glive ? g * GW : 0;
                    const unsigned all = __reduce_or_sync(FULL, em[0] ? (1u << ((sh + (int)jpos[0]) & 31)) : 0u);
                    const unsigned mine = (all & gmask) >> sh;
                    rank[0] = __popc(mine & ((1u << (jpos[0] & 31)) - 1u));
                    total = __popc(mine);
                } else {
                    // crossings per slot: D(v) for v = this slot's previous MEM end.  The
                    // loops run in lockstep over the warp's groups; `on` = my group still
                    // has work in this step.
                    uint32_t dcross[KPL];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) dcross[k] = chr ? 0x7FFFFFFFu : 0u;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        unsigned m = __ballot_sync(FULL, act && !chr && f[k] != e[k]) & gmask;
                        while (__any_sync(FULL, m != 0u)) {
                            const bool on = m != 0u;
                            const int src = on ? __ffs(m) - 1 : lane;
                            m &= m - 1;
                            const uint32_t x = __shfl_sync(FULL, e[k], src);
                            const uint32_t y = __shfl_sync(FULL, f[k], src);
#pragma unroll
                            for (int kk = 0; kk < KPL; ++kk)
                                dcross[kk] += (on && e[kk] >= x && e[kk] < y) ? 1u : 0u;
                        }
                    }
                    // one counting step per candidate value: G(v), tie index t
                    uint32_t mw[KPL];   // bitmap of the sorted positions that emit (group uniform)
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        mw[k] = 0;
                        em[k] = false;
                        jpos[k] = 0;
                    }
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        unsigned m = __ballot_sync(FULL, act && valid[k] && dcross[k] > 0u && e[k] >= p) & gmask;
                        while (__any_sync(FULL, m != 0u)) {
                            const bool on = m != 0u;
                            const int src = on ? __ffs(m) - 1 : lane;
                            m &= m - 1;
                            const uint32_t v = __shfl_sync(FULL, e[k], src);
                            const uint32_t dsrc = __shfl_sync(FULL, dcross[k], src);
                            uint32_t Gc = 0, tc = 0;
#pragma unroll
                            for (int kk = 0; kk < KPL; ++kk) {
                                Gc += __popc(__ballot_sync(FULL, e[kk] > v) & gmask);
                                if (kk <= k) {
                                    const unsigned eq = __ballot_sync(FULL, e[kk] == v) & gmask;
                                    tc += (kk < k) ? __popc(eq) : __popc(eq & ((1u << src) - 1u));
                                }
                            }
                            if (on && tc < dsrc) {
                                const uint32_t j = Gc + tc;
                                if (lane == src) {
                                    em[k] = true;
                                    jpos[k] = j;
                                }
#pragma unroll
                                for (int w = 0; w < KPL; ++w)
                                    if ((int)(j >> 5) == w) mw[w] |= 1u << (j & 31);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        uint32_t r = 0;
#pragma unroll
                        for (int w = 0; w < KPL; ++w) {
                            if (w < (int)(jpos[k] >> 5)) r += __popc(mw[w]);
                            else if (w == (int)(jpos[k] >> 5)) r += __popc(mw[w] & ((1u << (jpos[k] & 31)) - 1u));
                        }
                        rank[k] = r;
                    }
#pragma unroll
                    for (int w = 0; w < KPL; ++w) total += __popc(mw[w]);
                }

#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    if (em[k]) {
                        const uint32_t i = gcount + rank[k];
                        const uint32_t endv = min(e[k], lim);
                        if (direct) {
                            const unsigned long long gi = gbase + i;
                            if (gi < (unsigned long long)P.out_cap) {
                                P.scr_start[gi] = p;
                                P.scr_end[gi] = endv;
                                P.scr_order[gi] = jpos[k] + 1u;
                            }
                        } else if (i < (uint32_t)K) {
                            stg[3 * i + 0] = p;
                            stg[3 * i + 1] = endv;
                            stg[3 * i + 2] = jpos[k] + 1u;
                        }
                    }
                }
                gcount += total;
            }
            if (direct) break;

            // ---------------- the tile's block in the scratch area
            // exclusive prefix of the group counts (group leaders carry the count)
            const uint32_t mycnt = (glive && lg == 0) ? gcount : 0u;
            uint32_t incl = mycnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            const uint32_t gprefix = __shfl_sync(FULL, incl - mycnt, glive ? g * GW : 0);
            const bool ovf = __any_sync(FULL, glive && gcount > (uint32_t)K);
            unsigned long long base = 0;
            if (lane == 0) {
                if (total) base = atomicAdd(P.cursor, (unsigned long long)total);
                P.tile_cnt[tile] = total;
                P.tile_off[tile] = base;
            }
            base = __shfl_sync(FULL, base, 0);
            gbase = base + gprefix;
            __syncwarp();
            if (!ovf) {
                if (glive) {
                    for (uint32_t i = lg; i < gcount; i += GW) {
                        const unsigned long long gi = gbase + i;
                        if (gi < (unsigned long long)P.out_cap) {
                            P.scr_start[gi] = stg[3 * i + 0];
                            P.scr_end[gi] = stg[3 * i + 1];
                            P.scr_order[gi] = stg[3 * i + 2];
                        }
                    }
                }
                break;
            }
            ++replays;      // staging overflowed: recompute the tile with direct stores
        }
        __syncwarp();                    // stage s, list and staging are free again
        if (lane == 0) issue(s, tile + (long long)S * n_warps);
        if (++s == S) {
            s = 0;
            parity ^= 1u;
        }
    }
    if (irr_acc >> 31) P.result[MEMO_RES_IRREGULAR] = 1;
    if (lane == 0 && replays) atomicAdd((unsigned long long*)(P.result + MEMO_RES_REPLAYS), replays);
}

// ---------------------------------------------------------------- kernel 2
// partial[b] = index rows of tile block b; the last block to arrive turns
// partial[] into exclusive block offsets and writes the grand total.
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(const uint32_t* __restrict__ tile_cnt, long long n_tiles,
                 unsigned long long* partial, unsigned int* done, int64_t* result) {
    __shared__ unsigned long long red[SCAN_THREADS / 32];
    __shared__ bool is_last;
    const long long base = (long long)blockIdx.x * SCAN_BLOCK;
    unsigned long long sum = 0;
    for (int i = threadIdx.x; i < SCAN_BLOCK; i += SCAN_THREADS)
        if (base + i < n_tiles) sum += tile_cnt[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) t += red[i];
        *reinterpret_cast<volatile unsigned long long*>(&partial[blockIdx.x]) = t;
        __threadfence();
        is_last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // exclusive scan of partial[0 .. gridDim.x) by this block
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned b0 = 0; b0 < gridDim.x; b0 += SCAN_THREADS) {
        const unsigned i = b0 + threadIdx.x;
        const unsigned long long v =
            i < gridDim.x ? *reinterpret_cast<volatile unsigned long long*>(&partial[i]) : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) red[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += red[w];
        const unsigned long long c = carry;
        if (i < gridDim.x) partial[i] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) result[MEMO_RES_N_OUT] = (int64_t)carry;
}

// ---------------------------------------------------------------- kernel 3
__global__ void __launch_bounds__(SCAN_THREADS)
tile_gather_kernel(const uint32_t* __restrict__ tile_cnt, const unsigned long long* __restrict__ tile_off,
                   long long n_tiles, const unsigned long long* __restrict__ block_base,
                   const long long* __restrict__ seg_tile_start, int n_seg,
                   const uint32_t* __restrict__ scr_start, const uint32_t* __restrict__ scr_end,
                   const uint32_t* __restrict__ scr_order, int32_t* __restrict__ out_start,
                   uint32_t* __restrict__ out_end, int32_t* __restrict__ out_order, long long out_cap,
                   long long scr_cap, int64_t* __restrict__ seg_out_end) {
    __shared__ uint32_t excl[SCAN_BLOCK];          // exclusive row offset of each tile in the block
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const long long blk_lo = (long long)blockIdx.x * SCAN_BLOCK;
    const long long blk_hi = min(blk_lo + (long long)SCAN_BLOCK, n_tiles);
    const unsigned long long base = block_base[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // exclusive scan of the block's tile counts (thread t owns SCAN_ITEMS consecutive tiles)
    uint32_t c[SCAN_ITEMS];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const long long t = blk_lo + (long long)threadIdx.x * SCAN_ITEMS + i;
        c[i] = t < n_tiles ? tile_cnt[t] : 0u;
        tsum += c[i];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t run = incl - tsum;
    for (int w = 0; w < warp; ++w) run += wsum[w];
    uint32_t block_total = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) block_total += wsum[w];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        excl[threadIdx.x * SCAN_ITEMS + i] = run;
        run += c[i];
    }
    __syncthreads();

    // rows emitted up to the end of each record run that ends in this block:
    // run s ends where tile seg_tile_start[s + 1] begins
    for (int s = threadIdx.x; s < n_seg; s += SCAN_THREADS) {
        const long long t = seg_tile_start[s + 1];
        if (t > blk_lo && t <= blk_hi) {
            const uint32_t e = (t == blk_hi) ? block_total : excl[t - blk_lo];
            seg_out_end[s] = (int64_t)(base + e);
        }
    }
    if (out_cap == 0) return;

    // copy: each warp takes batches of 32 tiles; lane <-> tile for the metadata,
    // then the batch's rows are copied with consecutive lanes on consecutive rows
    const int n_batches = (int)((blk_hi - blk_lo + 31) / 32);
    for (int bt = warp; bt < n_batches; bt += SCAN_THREADS / 32) {
        const int ti = bt * 32 + lane;
        const bool have = blk_lo + ti < blk_hi;
        const uint32_t my_excl = have ? excl[ti] : 0xFFFFFFFFu;
        const uint32_t my_cnt = have ? tile_cnt[blk_lo + ti] : 0u;
        const unsigned long long my_off = have ? tile_off[blk_lo + ti] : 0ull;
        const uint32_t first = __shfl_sync(FULL, my_excl, 0);
        uint32_t tot = my_cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
        const uint32_t rel = have ? my_excl - first : 0xFFFFFFFFu;    // row offset of my tile in the batch
        for (uint32_t i = lane; i < ((tot + 31u) & ~31u); i += 32) {
            // owner tile of batch row i: the last lane whose rel <= i (rel is non-decreasing)
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t r = __shfl_sync(FULL, rel, lo + step);
                if (r <= i) lo += step;
            }
            const uint32_t orel = __shfl_sync(FULL, rel, lo);
            const unsigned long long ooff = __shfl_sync(FULL, my_off, lo);
            if (i < tot) {
                const unsigned long long src = ooff + (i - orel);
                const unsigned long long dst = base + first + i;
                if (src < (unsigned long long)scr_cap && dst < (unsigned long long)out_cap) {
                    out_start[dst] = (int32_t)scr_start[src];
                    out_end[dst] = scr_end[src];
                    out_order[dst] = (int32_t)scr_order[src];
                }
            }
        }
    }
}

// ---------------------------------------------------------------- host side
// phase B shape: KPL columns per lane; for KPL == 1 a row group is gw lanes wide
struct Geometry {
    int KPL, gw, all_pairs, sl, wcols, rb;
};

bool pick_geometry(int C, Geometry* geo) {
    static const int kpls[] = {1, 2, 3, 4, 6, 8, 16};
    for (int kpl : kpls)
        if (32 * kpl >= C) {
            geo->KPL = kpl;
            geo->gw = (kpl == 1 && C <= 16) ? C : 32;
            geo->all_pairs = (kpl == 1 && C <= 12) ? 1 : 0;
            geo->sl = geo->wcols = geo->rb = 0;          // phase A shape: set_scan_shape()
            return true;
        }
    return false;
}

// phase A shape for tiles of T rows: sl lanes share a row block (sl a power of two),
// 32 / sl row blocks of rb <= RBMAX rows
void set_scan_shape(int C, int T, Geometry* geo) {
    int sl = 1;
    while (sl < 32 && (32 / (2 * sl)) >= T) sl *= 2;     // no more rows per step than rows
    geo->sl = sl;
    geo->wcols = (C + sl - 1) / sl;
    geo->rb = 0;
}

stream_kernel_t select_kernel(const Geometry& g, bool order, int C, int ld) {
#define MEMO_CASE(KK) \
    if (g.KPL == KK) return order ? stream_kernel<KK, true, 0> : stream_kernel<KK, false, 0>;
    MEMO_CASE(1) MEMO_CASE(2) MEMO_CASE(3) MEMO_CASE(4) MEMO_CASE(6) MEMO_CASE(8) MEMO_CASE(16)
#undef MEMO_CASE
    return nullptr;
}

struct FastPlan {
    Geometry geo;
    int narrow, rpl;            // lane-per-row kernel (index_narrow.cu) and its rows per lane
    int T, K, stages, warps, ctas_per_sm;
    uint32_t chunk;             // scratch rows per warp reservation
    long long scr_cap;          // entries per scratch array
    uint32_t stage_bytes, warp_smem, off_bars, off_descs, off_stg, off_list;
    long long n_tiles, n_blocks;
    size_t smem;
    size_t off_segs, off_tstart, off_cnt, off_off, off_partial, off_ctrl, off_scratch, total;
};

int make_fast_plan(int64_t rows, int32_t C, int32_t ld, int64_t out_cap, const memo_segment_t* segs,
                   int32_t n_seg, const memo_index_opts_t* opts, FastPlan* plan, long long* tstart_host) {
    MEMO_REQUIRE(C >= 1, "n_cols must be >= 1");
    MEMO_REQUIRE(ld >= C, "ld must be >= n_cols");
    MEMO_REQUIRE(out_cap >= 0, "out_cap must be >= 0");
    MEMO_REQUIRE(n_seg >= 0 && (n_seg == 0 || segs != nullptr), "bad segment table");
    if (!pick_geometry(C, &plan->geo)) {
        set_error("n_cols = %d not supported (max 512)", C);
        return MEMO_ERR_UNSUPPORTED;
    }
    const int groups = 32 / plan->geo.gw;
    plan->warps = (opts && opts->warps_per_cta > 0) ? opts->warps_per_cta : 8;
    MEMO_REQUIRE(plan->warps >= 1 && plan->warps <= 8, "warps_per_cta must be 1..8");
    plan->stages = (opts && opts->stages > 0) ? opts->stages : 2;
    MEMO_REQUIRE(plan->stages <= MAX_STAGES, "stages must be <= %d", MAX_STAGES);
    plan->ctas_per_sm = (opts && opts->ctas_per_sm > 0) ? opts->ctas_per_sm : 0;
    const long long row_bytes = (long long)ld * 4;
    plan->rpl = 1;
    plan->narrow = (ld == C && !(opts && opts->kernel_variant == 1) &&
                    select_narrow_kernel(C, true, &plan->rpl) != nullptr) ? 1 : 0;
    if (!plan->narrow) plan->rpl = 1;
    long long T;
    if (plan->narrow) {
        // whole warp steps of 32 * rpl rows, ~4.5-9 KB of DAP per tile
        const long long step = 32ll * plan->rpl;
        long long it = (opts && opts->rows_per_tile > 0) ? (opts->rows_per_tile + step - 1) / step
                                                          : (5632 + step * row_bytes / 2) / (step * row_bytes);
        if (it < 1) it = 1;
        // keep the CTA's stages within the SM's shared memory
        const long long budget = (220 * 1024 / plan->warps - 1024) / plan->stages;
        while (it > 1 && (it * step > MAX_TILE_ROWS || (it * step + 2) * row_bytes + 144 > budget)) --it;
        T = it * step;
        plan->T = (int)T;
        plan->K = 0;
        plan->stage_bytes = (uint32_t)align_up((size_t)((T + 2) * row_bytes + 16), 128);
    } else {
        // a stage holds T + 2 rows; keep a warp's stages within its share of the SM
        const long long stage_budget = (200 * 1024 / plan->warps - 1536) / plan->stages;
        if (opts && opts->rows_per_tile > 0) {
            T = opts->rows_per_tile;
        } else {
            T = 5632 / row_bytes - 2;                // ~5.5 KB of DAP per tile: 16 warps per SM
        }
        if (T > MAX_TILE_ROWS) T = MAX_TILE_ROWS;
        if ((T + 2) * row_bytes + 64 > stage_budget) T = (stage_budget - 64) / row_bytes - 2;
        if (T < 1) T = 1;
        plan->T = (int)T;
        set_scan_shape(C, plan->T, &plan->geo);
        plan->stage_bytes = (uint32_t)align_up((size_t)((T + 2) * row_bytes + 64), 128);
        if (opts && opts->emit_buf_records > 0) {
            plan->K = opts->emit_buf_records;
        } else {
            long long k = (T * C) / (16ll * groups);  // ~10x the HPRC-shaped density
            if (k < 8) k = 8;
            if (k > 128) k = 128;
            plan->K = (int)k;
        }
    }
    size_t o = (size_t)plan->stages * plan->stage_bytes;
    plan->off_bars = (uint32_t)o;    o += 8 * MAX_STAGES;
    plan->off_descs = (uint32_t)o;   o += sizeof(TileDesc) * MAX_STAGES;
    plan->off_stg = (uint32_t)o;     o += plan->narrow ? 4 * 32 * (size_t)(C | 1) : 12 * (size_t)groups * plan->K;
    plan->off_list = (uint32_t)o;    o += 2 * (size_t)(T + 4);
    plan->warp_smem = (uint32_t)align_up(o, 128);
    plan->smem = (size_t)plan->warp_smem * plan->warps;
    MEMO_REQUIRE(plan->smem <= 227 * 1024, "tile configuration needs %zu B of shared memory", plan->smem);

    long long t = 0;
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
        if (tstart_host) tstart_host[i] = t;
        const long long primed = (s.flags & MEMO_SEG_PRIMED) ? 1 : 0;
        long long nt;
        if (plan->narrow) {
            // tiles start at multiples of rpl buffer rows (16-byte aligned bases)
            const long long fc = s.row_begin + primed, lc = s.row_begin + s.n_rows - 1;
            const long long g = ((fc - 1) / plan->rpl) * plan->rpl;
            nt = lc >= fc ? (lc - g + T - 1) / T : 0;
        } else {
            nt = (s.n_rows - primed + T - 1) / T;
        }
        t += nt > 0 ? nt : 1;
    }
    if (tstart_host) tstart_host[n_seg] = t;
    plan->n_tiles = t;
    plan->n_blocks = (t + SCAN_BLOCK - 1) / SCAN_BLOCK;
    size_t off = 0;
    plan->off_segs = off;    off = align_up(off + sizeof(memo_segment_t) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
    plan->off_tstart = off;  off = align_up(off + sizeof(long long) * (size_t)(n_seg + 1), 256);
    plan->off_cnt = off;     off = align_up(off + 4 * (size_t)(t > 0 ? t : 1), 256);
    plan->off_off = off;     off = align_up(off + 8 * (size_t)(t > 0 ? t : 1), 256);
    plan->off_partial = off; off = align_up(off + 8 * (size_t)(plan->n_blocks > 0 ? plan->n_blocks : 1), 256);
    plan->off_ctrl = off;    off = align_up(off + 256, 256);
    // scratch: warps reserve it in chunks (warp_alloc), which wastes < 1/4 of every
    // chunk plus each warp's last one: (4/3) (out_cap + warps * chunk) rows suffice
    const long long max_warps = (long long)device_sm_count() * 32;
    long long chunk = 4096;
    while (chunk > 256 && chunk * max_warps * 8 > out_cap) chunk >>= 1;
    while (chunk < 4ll * C) chunk <<= 1;
    plan->chunk = (uint32_t)chunk;
    plan->scr_cap = out_cap > 0 ? ((out_cap + max_warps * chunk) * 4 / 3 + 64) & ~3ll : 0;
    plan->off_scratch = off; off = align_up(off + 12 * (size_t)plan->scr_cap + 48, 256);
    plan->total = off;
    return MEMO_OK;
}

}  // namespace
}  // namespace memo

extern "C" {

size_t memo_index_workspace_bytes(int64_t rows, int32_t n_cols, int32_t ld, int64_t out_cap,
                                  const memo_segment_t* segs, int32_t n_seg,
                                  const memo_index_opts_t* opts) {
    memo::FastPlan plan;
    if (memo::make_fast_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, nullptr) != MEMO_OK)
        return 0;
    const size_t general = memo::general_workspace_bytes(rows, n_cols, segs, n_seg, opts);
    if (general == 0) return 0;
    return plan.total > general ? plan.total : general;
}

int memo_index_build(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld,
                     const memo_segment_t* segs, int32_t n_seg, const memo_index_opts_t* opts,
                     int32_t* out_start, uint32_t* out_end, int32_t* out_order, int64_t out_cap,
                     int64_t* seg_out_end, int64_t* result, void* workspace,
                     size_t workspace_bytes, void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(rows >= 0 && ld >= n_cols, "bad dap shape (rows=%lld, n_cols=%d, ld=%d)",
                 (long long)rows, n_cols, ld);
    MEMO_REQUIRE(result != nullptr, "result must not be NULL");
    MEMO_REQUIRE(out_cap == 0 || (out_start && out_end && out_order), "out_* NULL with out_cap > 0");
    MEMO_REQUIRE((reinterpret_cast<uintptr_t>(dap) & 15) == 0, "dap must be 16-byte aligned");
    MEMO_REQUIRE(n_seg == 0 || seg_out_end != nullptr, "seg_out_end must not be NULL");
    FastPlan plan;
    long long* tstart = new long long[(size_t)n_seg + 1];
    int rc = make_fast_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, tstart);
    if (rc != MEMO_OK) { delete[] tstart; return rc; }
    if (workspace_bytes < plan.total || workspace == nullptr) {
        delete[] tstart;
        set_error("workspace too small: %zu < %zu", workspace_bytes, plan.total);
        return MEMO_ERR_WORKSPACE;
    }
    cudaError_t e0 = cudaMemsetAsync(result, 0, sizeof(int64_t) * MEMO_RES_SLOTS, stream);
    if (e0 != cudaSuccess) { delete[] tstart; set_error("memset result: %s", cudaGetErrorString(e0)); return MEMO_ERR_CUDA; }
    if (n_seg == 0 || plan.n_tiles == 0) { delete[] tstart; return MEMO_OK; }

    char* ws = static_cast<char*>(workspace);
    // pageable host -> device copies are staged by the runtime before returning,
    // so the host buffers may be released right after these calls
    cudaError_t e1 = cudaMemcpyAsync(ws + plan.off_segs, segs, sizeof(memo_segment_t) * (size_t)n_seg,
                                     cudaMemcpyHostToDevice, stream);
    cudaError_t e2 = cudaMemcpyAsync(ws + plan.off_tstart, tstart, sizeof(long long) * (size_t)(n_seg + 1),
                                     cudaMemcpyHostToDevice, stream);
    delete[] tstart;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("segment table upload failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return MEMO_ERR_CUDA;
    }
    const bool order = opts ? (opts->order_mode != 0) : true;

    FastParams P{};
    P.dap = dap;
    P.total_bytes = (long long)rows * ld * 4;
    P.C = n_cols; P.ld = ld;
    P.segs = reinterpret_cast<const memo_segment_t*>(ws + plan.off_segs);
    P.seg_tile_start = reinterpret_cast<const long long*>(ws + plan.off_tstart);
    P.n_seg = n_seg; P.n_tiles = plan.n_tiles; P.T = plan.T; P.K = plan.K;
    P.stages = plan.stages; P.stage_bytes = plan.stage_bytes; P.warp_smem = plan.warp_smem;
    P.off_bars = plan.off_bars; P.off_descs = plan.off_descs;
    P.gw = plan.geo.gw; P.all_pairs = plan.geo.all_pairs; P.sl = plan.geo.sl; P.wcols = plan.geo.wcols; P.rb = plan.geo.rb;
    P.off_stg = plan.off_stg; P.off_list = plan.off_list;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(ws + plan.off_scratch);
    const size_t cap4 = (size_t)plan.scr_cap;
    P.scr_start = scratch; P.scr_end = scratch + cap4; P.scr_order = scratch + 2 * cap4;
    P.out_cap = out_cap; P.scr_cap = plan.scr_cap; P.chunk = plan.chunk;
    P.tile_cnt = reinterpret_cast<uint32_t*>(ws + plan.off_cnt);
    P.tile_off = reinterpret_cast<unsigned long long*>(ws + plan.off_off);
    P.cursor = reinterpret_cast<unsigned long long*>(ws + plan.off_ctrl);
    unsigned int* done = reinterpret_cast<unsigned int*>(ws + plan.off_ctrl + 64);
    unsigned long long* partial = reinterpret_cast<unsigned long long*>(ws + plan.off_partial);
    P.result = result;

    stream_kernel_t kern = plan.narrow ? select_narrow_kernel(n_cols, order, nullptr)
                                       : select_kernel(plan.geo, order, n_cols, ld);
    if (!kern) {
        set_error("no kernel for KPL=%d", plan.geo.KPL);
        return MEMO_ERR_UNSUPPORTED;
    }
    const int threads = plan.warps * 32;
    MEMO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    int per_sm = 0;
    MEMO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, plan.smem));
    if (per_sm < 1) {
        set_error("stream kernel does not fit on an SM (smem %zu B)", plan.smem);
        return MEMO_ERR_UNSUPPORTED;
    }
    if (plan.ctas_per_sm > 0 && plan.ctas_per_sm < per_sm) per_sm = plan.ctas_per_sm;
    long long grid = (long long)device_sm_count() * per_sm;
    const long long need = (plan.n_tiles + plan.warps - 1) / plan.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    MEMO_CUDA_TRY(cudaMemsetAsync(ws + plan.off_ctrl, 0, 256, stream));
    kern<<<(unsigned)grid, threads, plan.smem, stream>>>(P);
    MEMO_CUDA_TRY(cudaGetLastError());
    tile_scan_kernel<<<(unsigned)plan.n_blocks, SCAN_THREADS, 0, stream>>>(P.tile_cnt, plan.n_tiles, partial,
                                                                          done, result);
    MEMO_CUDA_TRY(cudaGetLastError());
    tile_gather_kernel<<<(unsigned)plan.n_blocks, SCAN_THREADS, 0, stream>>>(
        P.tile_cnt, P.tile_off, plan.n_tiles, partial, P.seg_tile_start, n_seg, P.scr_start, P.scr_end,
        P.scr_order, out_start, out_end, out_order, out_cap, plan.scr_cap, seg_out_end);
    MEMO_CUDA_TRY(cudaGetLastError());
    return MEMO_OK;
}

}  // extern "C"
