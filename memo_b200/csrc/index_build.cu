// DAP -> MEMO index rows on sm_100a: the single-pass build for matching statistics
// (host plan, scan and gather kernels; the streaming kernels live in
// index_narrow.cu and index_wide.cu).
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
// Four kernels per build (0: prep_kernel, the record runs as kernel parameters + cleared control words):
//  1 a streaming kernel reads the DAP exactly once (bulk async copies into
//    per-warp shared-memory rings; warps never wait for each other) and appends
//    the index rows of its work units to a scratch area, unordered:
//      narrow_kernel  n_cols <= 16: unit = strip of G tiles of T rows, one lane per row
//      wide_kernel    otherwise:    unit = strip of R rows, one warp per strip,
//                     sorted row kept in registers and updated incrementally
//    A unit's index rows are one block of consecutive scratch rows (wide: or a
//    short chain of blocks); tile_cnt / tile_off describe it.
//  2 tile_scan_kernel   block sums of the unit counts; the last block to finish
//    scans the block sums.
//  3 strip_gather_kernel   exclusive scan inside each block
//    of units and copy of every unit's rows (8-byte scratch rows: MEM end, row
//    number within the unit << 16 | order) from the scratch area to their place
//    in the ordered output (the extra traffic is 16 B per index row, a few % of
//    the DAP).
// If the input is irregular the result must be discarded and the general build
// (index_general.cu) run instead; memo_index_build reports that in
// result[MEMO_RES_IRREGULAR].
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "index_fast.cuh"

namespace memo {
namespace {

// event pairs around the streaming kernel (memo_profile_enable / _collect)
struct Profile {
    bool on = false;
    std::vector<cudaEvent_t> events;
};
thread_local Profile g_profile;
int pick_build(int32_t C, int32_t ld, const memo_index_opts_t* opts);

// ---------------------------------------------------------------- prep
// The record runs and their first units go to the device as KERNEL PARAMETERS, and the same
// launch clears the result slots and the control words: a build then contains no copy-engine
// operation at all.  (cudaMemcpyAsync from pageable memory synchronises the stream before it
// copies -- with a 256 MB chunk copy queued ahead of the build, the streaming host path stalled
// for that copy on every chunk -- and small copies / memsets may queue behind a large host ->
// device copy on the same engine.)
constexpr int PREP_SEGS = 64;
struct PrepArgs {
    memo_segment_t segs[PREP_SEGS];
    long long tstart[PREP_SEGS + 1];
    memo_segment_t* d_segs;
    long long* d_tstart;
    int64_t* result;            // MEMO_RES_SLOTS slots, cleared by the first batch
    unsigned int* ctrl;         // 64 control words, cleared by the first batch
    int first, n;               // runs [first, first + n) and tstart[first .. first + n]
};
__global__ void prep_kernel(const PrepArgs a) {
    const int t = threadIdx.x;
    if (t < a.n) a.d_segs[a.first + t] = a.segs[t];
    if (t <= a.n) a.d_tstart[a.first + t] = a.tstart[t];
    if (a.first == 0) {
        if (t < MEMO_RES_SLOTS) a.result[t] = 0;
        if (t < 64) a.ctrl[t] = 0u;
    }
}

// ---------------------------------------------------------------- scan
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

// ---------------------------------------------------------------- gather
// Common first part of the gather kernels: exclusive scan of the block's unit
// counts into excl[] (shared memory), and the per-run row totals.  Returns the
// block's total.
__device__ __forceinline__ uint32_t gather_block_scan(
    const uint32_t* __restrict__ tile_cnt, long long n_tiles, long long blk_lo, long long blk_hi,
    unsigned long long base, const long long* __restrict__ seg_tile_start, int n_seg,
    int64_t* __restrict__ seg_out_end, bool write_segs, uint32_t* excl, uint32_t* wsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread t owns SCAN_ITEMS consecutive units
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
    // run s ends where unit seg_tile_start[s + 1] begins
    if (write_segs) {
        for (int s = threadIdx.x; s < n_seg; s += SCAN_THREADS) {
            const long long t = seg_tile_start[s + 1];
            if (t > blk_lo && t <= blk_hi) {
                const uint32_t e = (t == blk_hi) ? block_total : excl[t - blk_lo];
                seg_out_end[s] = (int64_t)(base + e);
            }
        }
    }
    return block_total;
}

// A unit (strip) wrote one block of scratch rows, or a chain of them.  A CTA stages the
// metadata of its block of 1024 units in shared memory; gridDim.y CTAs share a block, each
// copying a contiguous slice of the block's output rows in chunks of consecutive rows per
// warp (consecutive lanes, consecutive output rows).  One search for the unit of a chunk's
// first row (warp uniform), after that every lane walks the units forward as its rows
// advance: a few instructions per row (a binary search per row made this kernel
// instruction bound: 5 warp instructions per row).  The source row comes from the unit's
// block chain.
__global__ void __launch_bounds__(SCAN_THREADS)
strip_gather_kernel(const uint32_t* __restrict__ tile_cnt, const unsigned long long* __restrict__ tile_off,
                  const uint32_t* __restrict__ unit_pos0, const uint32_t* __restrict__ unit_aux,
                  const uint32_t* __restrict__ first_cnt, const int32_t* __restrict__ unit_next,
                  const BlockRec* __restrict__ pool, uint32_t pool_cap, long long n_tiles,
                  const unsigned long long* __restrict__ block_base,
                  const long long* __restrict__ seg_tile_start, int n_seg,
                  const uint32_t* __restrict__ scr, int32_t* __restrict__ out_start,
                  uint32_t* __restrict__ out_end, int32_t* __restrict__ out_order, long long out_cap,
                  long long scr_cap, int64_t* __restrict__ seg_out_end, int n_blocks, int split) {
    __shared__ uint32_t excl[SCAN_BLOCK + 1];
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t s_fcnt[SCAN_BLOCK];
    __shared__ int32_t s_next[SCAN_BLOCK];
    __shared__ unsigned long long s_off[SCAN_BLOCK];
    // The grid is one wave of resident CTAs; a CTA walks work items (block of units, slice of the
    // block's rows), neighbouring CTAs share a block.  (One CTA per item left the last wave of a
    // grid of n_blocks x split CTAs partly empty: 1060 CTAs on 740 slots at chr1 x 94.)
    for (long long item = blockIdx.x; item < (long long)n_blocks * split; item += gridDim.x) {
    const int blk = (int)(item / split), slice = (int)(item % split);
    __syncthreads();                                            // (shared arrays of the previous item)
    const long long blk_lo = (long long)blk * SCAN_BLOCK;
    const long long blk_hi = min(blk_lo + (long long)SCAN_BLOCK, n_tiles);
    const unsigned long long base = block_base[blk];
    const uint32_t block_total = gather_block_scan(tile_cnt, n_tiles, blk_lo, blk_hi, base, seg_tile_start,
                                                   n_seg, seg_out_end, slice == 0, excl, wsum);
    if (out_cap == 0 || block_total == 0) continue;
    const int n_units = (int)(blk_hi - blk_lo);
    for (int u = threadIdx.x; u < n_units; u += SCAN_THREADS) {
        s_off[u] = tile_off[blk_lo + u];
        s_fcnt[u] = first_cnt[blk_lo + u];
        s_next[u] = unit_next[blk_lo + u];
    }
    if (threadIdx.x == 0) excl[n_units] = block_total;          // upper end of the last unit
    __syncthreads();
    // this item's slice of the block's rows, in chunks of CH consecutive rows per warp
    const uint32_t lo = (uint32_t)((unsigned long long)block_total * slice / split);
    const uint32_t hi = (uint32_t)((unsigned long long)block_total * (slice + 1) / split);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int U = 4;                            // rows per lane in flight
    constexpr uint32_t CH = 32 * U * 4;
    for (uint32_t c0 = lo + warp * CH; c0 < hi; c0 += (SCAN_THREADS / 32) * CH) {
        const uint32_t c1 = min(c0 + CH, hi);
        int u = 0;
        {
            int a = 0, b = n_units - 1;             // last unit with excl[u] <= c0 (skips empty ones)
            while (a < b) {
                const int m = (a + b + 1) >> 1;
                if (excl[m] <= c0) a = m; else b = m - 1;
            }
            u = a;
        }
        for (uint32_t b0 = c0; b0 < c1; b0 += 32 * U) {
            unsigned long long src[U];
            uint32_t base_pos[U], chr_start[U];
            uint2 v[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t d = b0 + j * 32 + lane;
                src[j] = ~0ull;
                if (d < c1) {
                    while (excl[u + 1] <= d) ++u;   // d < block_total = excl[n_units]: stops in range
                    uint32_t r = d - excl[u];
                    unsigned long long off = s_off[u];
                    uint32_t cnt = s_fcnt[u];
                    int32_t nx = s_next[u];
                    bool ok = true;
                    while (r >= cnt) {              // continue in the unit's next block
                        if (nx < 0 || (uint32_t)nx >= pool_cap) { ok = false; break; }
                        r -= cnt;
                        const BlockRec rec = pool[nx];
                        off = rec.off;
                        cnt = rec.cnt;
                        nx = rec.next;
                    }
                    if (ok && off + r < (unsigned long long)scr_cap) {
                        src[j] = off + r;
                        base_pos[j] = unit_pos0[blk_lo + u];      // (the lanes of a warp share a few units)
                        chr_start[j] = unit_aux[blk_lo + u];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (src[j] != ~0ull) v[j] = reinterpret_cast<const uint2*>(scr)[src[j]];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const unsigned long long d = base + b0 + j * 32 + lane;
                if (src[j] != ~0ull && d < (unsigned long long)out_cap) {
                    const uint32_t rowf = v[j].y >> 16;
                    out_start[d] = (int32_t)(rowf == SCR_ROW_CHR ? chr_start[j] : base_pos[j] + rowf);
                    out_end[d] = v[j].x;
                    out_order[d] = (int32_t)(v[j].y & 0xFFFFu);
                }
            }
        }
    }
    }
}

// ---------------------------------------------------------------- host side
struct FastPlan {
    int narrow, rpl;            // lane-per-row kernel (index_narrow.cu) and its rows per lane
    int kpl;                    // wide kernel: sorted positions per lane
    int T, R, stages, warps, ctas_per_sm;
    uint32_t pool_cap;          // wide: BlockRec records
    uint32_t chunk;             // scratch rows per warp reservation
    long long scr_cap;          // entries per scratch array
    uint32_t stage_bytes, warp_smem, off_bars, off_descs, off_stg, off_list;
    long long n_units, n_blocks;
    size_t smem;
    size_t off_segs, off_tstart, off_cnt, off_off, off_partial, off_ctrl, off_fcnt, off_next, off_pool,
        off_pos0, off_aux, off_scratch, total;
};

// tuning knobs read once from the environment (experiments; the defaults are the measured best)
int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int make_fast_plan(int64_t rows, int32_t C, int32_t ld, int64_t out_cap, const memo_segment_t* segs,
                   int32_t n_seg, const memo_index_opts_t* opts, FastPlan* plan, long long* tstart_host) {
    MEMO_REQUIRE(C >= 1, "n_cols must be >= 1");
    MEMO_REQUIRE(ld >= C, "ld must be >= n_cols");
    MEMO_REQUIRE(out_cap >= 0, "out_cap must be >= 0");
    MEMO_REQUIRE(n_seg >= 0 && (n_seg == 0 || segs != nullptr), "bad segment table");
    plan->kpl = 0;
    if (select_wide_kernel(C, true, &plan->kpl) == nullptr) {
        set_error("n_cols = %d not supported (max 512)", C);
        return MEMO_ERR_UNSUPPORTED;
    }
    plan->rpl = 1;
    plan->narrow = (ld == C && !(opts && opts->kernel_variant == 1) &&
                    select_narrow_kernel(C, true, &plan->rpl) != nullptr) ? 1 : 0;
    if (!plan->narrow) plan->rpl = 1;
    // measured on B200: 2 x 8 warps per SM for tiles, 4 x 6 warps for strips
    plan->warps = (opts && opts->warps_per_cta > 0) ? opts->warps_per_cta : (plan->narrow ? 8 : 6);
    MEMO_REQUIRE(plan->warps >= 1 && plan->warps <= 8, "warps_per_cta must be 1..8");
    // measured on B200: tiles double-buffered per warp; strips in ONE large stage per warp (the
    // other warps of the SM cover the copy: bigger chunks, fewer producer steps)
    plan->stages = (opts && opts->stages > 0) ? opts->stages : (plan->narrow ? 2 : 1);
    MEMO_REQUIRE(plan->stages <= MAX_STAGES, "stages must be <= %d", MAX_STAGES);
    plan->ctas_per_sm = (opts && opts->ctas_per_sm > 0) ? opts->ctas_per_sm : 0;
    const long long row_bytes = (long long)ld * 4;

    // scratch: warps reserve it in chunks; a request that does not fit abandons fewer
    // rows than it asks for, and every warp leaves its last chunk unfinished:
    // 2 out_cap + warps * chunk rows suffice.  Chunks of 1-4 K rows when the output
    // is large enough for that slack not to matter.
    const long long max_warps = (long long)device_sm_count() * 32;
    long long chunk = 4096;
    while (chunk > 256 && chunk * max_warps > out_cap) chunk >>= 1;
    while (chunk < 4ll * C) chunk <<= 1;
    plan->chunk = (uint32_t)chunk;
    plan->scr_cap = out_cap > 0 ? (2 * out_cap + max_warps * chunk + 64) & ~3ll : 0;
    // every chunk a strip crosses into costs one BlockRec
    plan->pool_cap = (uint32_t)(plan->scr_cap / (chunk / 2) + max_warps + 16);

    // a warp's stages must fit its share of the SM's shared memory (next to its queue of
    // flagged rows, row list, barriers and descriptors)
    const long long queue_bytes = plan->narrow ? 4 * 40 * (long long)((2 * C + 2) | 1) : 0;   // NARROW_QCAP entries
    const long long budget = (220 * 1024 / plan->warps - 1024 - queue_bytes - 2 * (MAX_TILE_ROWS + 4)) / plan->stages;
    long long T;
    if (plan->narrow) {
        // tiles: whole warp steps of 32 * rpl rows, ~4.5-9 KB of DAP per tile
        const long long step = 32ll * plan->rpl;
        long long it = (opts && opts->rows_per_tile > 0) ? (opts->rows_per_tile + step - 1) / step
                                                          : (5632 + step * row_bytes / 2) / (step * row_bytes);
        if (it < 1) it = 1;
        while (it > 1 && (it * step > MAX_TILE_ROWS || (it * step + 2) * row_bytes + 144 > budget)) --it;
        T = it * step;
        // strips of R tiles (~1 K rows): one scratch block chain and one gather unit each
        long long G = (opts && opts->emit_buf_records > 0) ? opts->emit_buf_records : (1024 + T - 1) / T;
        if (G < 1) G = 1;
        MEMO_REQUIRE(G * T <= SCR_MAX_UNIT_ROWS, "strip of %lld x %lld rows: at most %lld rows per unit", G, T,
                     SCR_MAX_UNIT_ROWS);
        plan->R = (int)G;
        plan->stage_bytes = (uint32_t)align_up((size_t)((T + 2) * row_bytes + 16), 128);
    } else {
        // strips of R rows stream through the ring in chunks of T rows: as many as keep four
        // CTAs of `warps` warps on an SM (~9 KB per warp)
        const long long per_warp = (227 * 1024 / 4 - 1024) / plan->warps - 8 * MAX_STAGES -
                                   (long long)sizeof(TileDesc) * MAX_STAGES - 128;
        T = (opts && opts->rows_per_tile > 0) ? opts->rows_per_tile
                                              : (per_warp / plan->stages - 128 - 32 - 128 * plan->kpl) / row_bytes;
        if (T > MAX_TILE_ROWS) T = MAX_TILE_ROWS;
        if (T * row_bytes + 160 + 128 * plan->kpl > budget) T = (budget - 160 - 128 * plan->kpl) / row_bytes;
        if (T < 1) T = 1;
        // Up to 460 compare rows per strip, a whole number of chunks (predecessor row included): the
        // sorted row is built from scratch once per strip (~300 warp instructions; at 115 rows per
        // strip that was 2.6 per row, 4 % of the kernel).  Short inputs keep short strips: every
        // resident warp (4 x 6 per SM) should see >= 48 strips, or the last strips -- and the dense
        // first strip of a record, one warp's work -- become the kernel's tail (94 genomes x 10 Mbp:
        // 0.85 ms with 115-row strips, 1.37 ms with 460; 31 M rows, a shard of chr1 on 8 GPUs:
        // 2.486 ms with 160 - 230 rows, 2.529 with 460; 62 M rows: 4.84 ms with 300, 4.875 with 460).
        static const long long strip_rows_max = env_int("MEMO_WIDE_STRIP_ROWS", 460);
        long long strip_rows = rows / ((long long)device_sm_count() * 24 * 48);
        if (strip_rows < 115) strip_rows = 115;
        if (strip_rows > strip_rows_max) strip_rows = strip_rows_max;
        long long R = (opts && opts->emit_buf_records > 0) ? opts->emit_buf_records
                                                           : (strip_rows / T > 1 ? strip_rows / T : 1) * T - 1;
        if (R < 1) R = 1;
        MEMO_REQUIRE(R <= SCR_MAX_UNIT_ROWS, "strip of %lld rows: at most %lld rows per unit", R, SCR_MAX_UNIT_ROWS);
        plan->R = (int)R;
        // (slots of lanes past the last column read up to 128 * kpl bytes beyond a row)
        plan->stage_bytes = (uint32_t)align_up((size_t)(T * row_bytes + 32 + 128 * plan->kpl), 128);
    }
    plan->T = (int)T;
    size_t o = (size_t)plan->stages * plan->stage_bytes;
    plan->off_bars = (uint32_t)o;    o += 8 * MAX_STAGES;
    plan->off_descs = (uint32_t)o;   o += sizeof(TileDesc) * MAX_STAGES;
    plan->off_stg = (uint32_t)o;     o += (size_t)queue_bytes;
    plan->off_list = (uint32_t)o;    o += plan->narrow ? 2 * (size_t)(T + 4) : 0;
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
            if (nt < 1) nt = 1;
            nt = (nt + plan->R - 1) / plan->R;
        } else {
            nt = (s.n_rows - primed + plan->R - 1) / plan->R;
        }
        t += nt > 0 ? nt : 1;          // a one-row run still owns its chr-end rows
    }
    if (tstart_host) tstart_host[n_seg] = t;
    plan->n_units = t;
    plan->n_blocks = (t + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const size_t ns = (size_t)(t > 0 ? t : 1);
    size_t off = 0;
    plan->off_segs = off;    off = align_up(off + sizeof(memo_segment_t) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
    plan->off_tstart = off;  off = align_up(off + sizeof(long long) * (size_t)(n_seg + 1), 256);
    plan->off_cnt = off;     off = align_up(off + 4 * ns, 256);
    plan->off_off = off;     off = align_up(off + 8 * ns, 256);
    plan->off_partial = off; off = align_up(off + 8 * (size_t)(plan->n_blocks > 0 ? plan->n_blocks : 1), 256);
    plan->off_ctrl = off;    off = align_up(off + 256, 256);
    plan->off_fcnt = off;    off = align_up(off + 4 * ns, 256);
    plan->off_next = off;    off = align_up(off + 4 * ns, 256);
    plan->off_pool = off;    off = align_up(off + sizeof(BlockRec) * (size_t)plan->pool_cap, 256);
    plan->off_pos0 = off;    off = align_up(off + 4 * ns, 256);
    plan->off_aux = off;     off = align_up(off + 4 * ns, 256);
    plan->off_scratch = off; off = align_up(off + 4 * SCR_WORDS * (size_t)(plan->scr_cap + SCR_DUMP_ROWS) + 48, 256);
    plan->total = off;
    return MEMO_OK;
}

// Resident CTAs per SM of a streaming kernel for a block size and dynamic shared-memory size on
// the current device.  The kernel's shared-memory limit is raised to the device maximum once per
// (device, kernel) and never lowered; the occupancy query runs once per shape (it used to run,
// with the attribute call, on every launch: 362 chunks for chr1 x 94 at 256 MB).
int stream_ctas_per_sm(stream_kernel_t kern, int threads, size_t smem) {
    struct Entry { int dev; stream_kernel_t k; int threads; size_t smem; int n; };
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
        if (cache[i].threads == threads && cache[i].smem == smem) return cache[i].n;
    }
    int n = 0;
    if (kern == nullptr) {                  // the gather kernel (static shared memory only)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, strip_gather_kernel, threads, smem) != cudaSuccess) n = 0;
    } else {
        if (!raised &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess) n = 0;
    }
    if (n_cache == 64) n_cache = 0;
    cache[n_cache++] = Entry{dev, kern, threads, smem, n};
    return n;
}

// which build a shape takes: 0 = lane-per-row tiles (narrow rows), 2 = the strip kernel with
// scratch + gather (index_wide.cu), 1 = the single-kernel strip build (index_wide2.cu: ordered
// in-place writes; measured slower on B200 -- DESIGN.md 4.3 -- so only on request: kernel_variant 3)
int pick_build(int32_t C, int32_t ld, const memo_index_opts_t* opts) {
    const int variant = opts ? opts->kernel_variant : 0;
    if (variant == 3 && wide2_supported(C, ld)) return 1;
    const bool narrow = ld == C && variant == 0 && select_narrow_kernel(C, true, nullptr) != nullptr;
    return narrow ? 0 : 2;
}

}  // namespace

void profile_begin(cudaStream_t stream) {
    if (!g_profile.on) return;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, stream);
    g_profile.events.push_back(ev);
}

void profile_end(cudaStream_t stream) {
    if (!g_profile.on || (g_profile.events.size() & 1) == 0) return;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) {
        cudaEventDestroy(g_profile.events.back());
        g_profile.events.pop_back();
        return;
    }
    cudaEventRecord(ev, stream);
    g_profile.events.push_back(ev);
}

}  // namespace memo

extern "C" {

int memo_profile_enable(int32_t on) {
    memo::g_profile.on = on != 0;
    return MEMO_OK;
}

int memo_profile_collect(double* stream_kernel_ms, int32_t* n_builds) {
    using namespace memo;
    double total = 0.0;
    const size_t n = g_profile.events.size() / 2;
    for (size_t i = 0; i < n; ++i) {
        float ms = 0.f;
        MEMO_CUDA_TRY(cudaEventSynchronize(g_profile.events[2 * i + 1]));
        MEMO_CUDA_TRY(cudaEventElapsedTime(&ms, g_profile.events[2 * i], g_profile.events[2 * i + 1]));
        total += ms;
        cudaEventDestroy(g_profile.events[2 * i]);
        cudaEventDestroy(g_profile.events[2 * i + 1]);
    }
    g_profile.events.clear();
    if (stream_kernel_ms) *stream_kernel_ms = total;
    if (n_builds) *n_builds = (int32_t)n;
    return MEMO_OK;
}

size_t memo_index_workspace_bytes(int64_t rows, int32_t n_cols, int32_t ld, int64_t out_cap,
                                  const memo_segment_t* segs, int32_t n_seg,
                                  const memo_index_opts_t* opts) {
    memo::FastPlan plan;
    if (memo::make_fast_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, nullptr) != MEMO_OK)
        return 0;
    const size_t general = memo::general_workspace_bytes(rows, n_cols, segs, n_seg, opts);
    if (general == 0) return 0;
    size_t fast = plan.total;
    if (memo::pick_build(n_cols, ld, opts) == 1) {
        fast = memo::wide2_workspace_bytes(rows, n_cols, ld, out_cap, segs, n_seg, opts);
        if (fast == 0) return 0;
    }
    return fast > general ? fast : general;
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
    MEMO_REQUIRE(n_cols >= 1 && out_cap >= 0, "bad n_cols / out_cap");
    if (pick_build(n_cols, ld, opts) == 1)
        return launch_wide2(dap, rows, n_cols, ld, segs, n_seg, opts, out_start, out_end, out_order, out_cap,
                            seg_out_end, result, workspace, workspace_bytes, stream);
    FastPlan plan;
    std::vector<long long> tstart_v((size_t)n_seg + 1);
    long long* tstart = tstart_v.data();
    int rc = make_fast_plan(rows, n_cols, ld, out_cap, segs, n_seg, opts, &plan, tstart);
    if (rc != MEMO_OK) return rc;
    if (workspace_bytes < plan.total || workspace == nullptr) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, plan.total);
        return MEMO_ERR_WORKSPACE;
    }
    if (n_seg == 0 || plan.n_units == 0) {
        MEMO_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(int64_t) * MEMO_RES_SLOTS, stream));
        return MEMO_OK;
    }

    char* ws = static_cast<char*>(workspace);
    for (int first = 0; first < n_seg; first += PREP_SEGS) {
        PrepArgs a;
        a.n = n_seg - first < PREP_SEGS ? n_seg - first : PREP_SEGS;
        a.first = first;
        for (int i = 0; i < a.n; ++i) a.segs[i] = segs[first + i];
        for (int i = 0; i <= a.n; ++i) a.tstart[i] = tstart[first + i];
        a.d_segs = reinterpret_cast<memo_segment_t*>(ws + plan.off_segs);
        a.d_tstart = reinterpret_cast<long long*>(ws + plan.off_tstart);
        a.result = result;
        a.ctrl = reinterpret_cast<unsigned int*>(ws + plan.off_ctrl);
        prep_kernel<<<1, 96, 0, stream>>>(a);
        MEMO_LAUNCH_CHECK(1);
    }
    const bool order = opts ? (opts->order_mode != 0) : true;

    FastParams P{};
    P.dap = dap;
    P.total_bytes = (long long)rows * ld * 4;
    P.C = n_cols; P.ld = ld;
    P.segs = reinterpret_cast<const memo_segment_t*>(ws + plan.off_segs);
    P.seg_tile_start = reinterpret_cast<const long long*>(ws + plan.off_tstart);
    P.n_seg = n_seg; P.n_tiles = plan.n_units; P.T = plan.T; P.R = plan.R;
    P.stages = plan.stages; P.stage_bytes = plan.stage_bytes; P.warp_smem = plan.warp_smem;
    P.off_bars = plan.off_bars; P.off_descs = plan.off_descs;
    P.off_stg = plan.off_stg; P.off_list = plan.off_list;
    P.scr = reinterpret_cast<uint32_t*>(ws + plan.off_scratch);
    P.out_cap = out_cap; P.scr_cap = plan.scr_cap; P.chunk = plan.chunk;
    P.tile_cnt = reinterpret_cast<uint32_t*>(ws + plan.off_cnt);
    P.tile_off = reinterpret_cast<unsigned long long*>(ws + plan.off_off);
    P.cursor = reinterpret_cast<unsigned long long*>(ws + plan.off_ctrl);
    P.strip_counter = reinterpret_cast<unsigned long long*>(ws + plan.off_ctrl + 128);
    P.pool_counter = reinterpret_cast<unsigned int*>(ws + plan.off_ctrl + 192);
    P.first_cnt = reinterpret_cast<uint32_t*>(ws + plan.off_fcnt);
    P.unit_pos0 = reinterpret_cast<uint32_t*>(ws + plan.off_pos0);
    P.unit_aux = reinterpret_cast<uint32_t*>(ws + plan.off_aux);
    P.unit_next = reinterpret_cast<int32_t*>(ws + plan.off_next);
    P.pool = reinterpret_cast<BlockRec*>(ws + plan.off_pool);
    P.pool_cap = plan.pool_cap;
    // (measured at chr1 x 94, kernel ms: no prefetch 19.19, plain prefetch 19.09, prefetch evict_last +
    //  copies evict_first 18.84 -- what is built in --, evict_first copies without prefetch 19.50)
    static const int prefetch = env_int("MEMO_WIDE_PREFETCH", 1) != 0;
    P.prefetch = prefetch;
    unsigned int* done = reinterpret_cast<unsigned int*>(ws + plan.off_ctrl + 64);
    unsigned long long* partial = reinterpret_cast<unsigned long long*>(ws + plan.off_partial);
    P.result = result;

    stream_kernel_t kern = plan.narrow ? select_narrow_kernel(n_cols, order, nullptr)
                                       : select_wide_kernel(n_cols, order, nullptr);
    if (!kern) {
        set_error("no kernel for n_cols=%d", n_cols);
        return MEMO_ERR_UNSUPPORTED;
    }
    const int threads = plan.warps * 32;
    const int per_sm_q = stream_ctas_per_sm(kern, threads, plan.smem);
    int per_sm = per_sm_q;
    if (per_sm < 1) {
        set_error("stream kernel does not fit on an SM (smem %zu B)", plan.smem);
        return MEMO_ERR_UNSUPPORTED;
    }
    if (plan.ctas_per_sm > 0 && plan.ctas_per_sm < per_sm) per_sm = plan.ctas_per_sm;
    long long grid = (long long)device_sm_count() * per_sm;
    const long long need = (plan.n_units + plan.warps - 1) / plan.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    profile_begin(stream);
    kern<<<(unsigned)grid, threads, plan.smem, stream>>>(P);
    MEMO_LAUNCH_CHECK(1);
    profile_end(stream);
    tile_scan_kernel<<<(unsigned)plan.n_blocks, SCAN_THREADS, 0, stream>>>(P.tile_cnt, plan.n_units, partial,
                                                                          done, result);
    MEMO_LAUNCH_CHECK(1);
    {
        // one wave of resident CTAs walks n_blocks x split work items (>= 6 items per CTA where the
        // input is large enough: the slowest CTA then decides a few % of the kernel, not a wave)
        const int occ = stream_ctas_per_sm(nullptr, SCAN_THREADS, 0);
        const long long slots = (long long)(occ > 0 ? occ : 4) * device_sm_count();
        long long split = (6 * slots + plan.n_blocks - 1) / plan.n_blocks;
        if (split < 1) split = 1;
        if (split > 32) split = 32;
        long long grid = plan.n_blocks * split < slots ? plan.n_blocks * split : slots;
        strip_gather_kernel<<<(unsigned)grid, SCAN_THREADS, 0, stream>>>(
            P.tile_cnt, P.tile_off, P.unit_pos0, P.unit_aux, P.first_cnt, P.unit_next, P.pool, P.pool_cap,
            plan.n_units, partial,
            P.seg_tile_start, n_seg, P.scr, out_start, out_end, out_order, out_cap, plan.scr_cap,
            seg_out_end, (int)plan.n_blocks, (int)split);
    }
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // extern "C"
