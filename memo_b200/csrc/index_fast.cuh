// Shared pieces of the single-pass index build (index_build.cu: generic warp-stream
// kernel, scan, gather, host plan; index_narrow.cu: the lane-per-row kernel for
// narrow DAP rows).
#pragma once
#include "common.cuh"

namespace memo {
namespace {

constexpr int MAX_STAGES = 4;
constexpr int MAX_TILE_ROWS = 960;
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;   // tiles per scan / gather block
constexpr int SCR_WORDS = 2;                            // words per scratch row
constexpr uint32_t SCR_ROW_CHR = 0xFFFFu;               // row field of a chr-end row
constexpr long long SCR_MAX_UNIT_ROWS = 65000;          // DAP rows per unit: the row field has 16 bits
constexpr long long SCR_DUMP_ROWS = 4096;              // rows behind the scratch area that take the stores of chunks past its end

struct alignas(16) TileDesc {
    int n;              // rows in the stage (narrow: T compare rows after row 0)
    int off;            // wide: word offset of row 0 inside the stage data; narrow: strip
    uint32_t pos_h;     // record-relative position of row 0
    uint32_t rec_len;
    int flags;          // WD_* (index_fast.cuh)
    int r_lo, r_hi;     // narrow: live compare rows of the tile are r_lo..r_hi (1-based); wide: r_lo = strip
    int pad;            // narrow: tile row of the run's last row
};

// stage descriptor flags
constexpr int WD_FIRST = 1;    // first stage of a strip (wide: row 0 is the strip's predecessor row)
constexpr int WD_LAST = 2;     // last stage of the strip
constexpr int WD_CHR = 4;      // chr-end rows follow (wide: the strip; narrow: the run, see WD_RUNLAST)
constexpr int WD_END = 8;      // no more strips
constexpr int WD_RUNLAST = 16; // narrow: last tile of its run

struct BlockRec {          // a block of consecutive scratch rows continuing a strip's output
    unsigned long long off;
    uint32_t cnt;
    int32_t next;          // index of the next BlockRec of the strip, -1 = none
};

struct FastParams {
    const int32_t* dap;
    long long total_bytes;             // rows * ld * 4
    int32_t C;
    int32_t ld;
    const memo_segment_t* segs;        // device copy
    const long long* seg_tile_start;   // device [n_seg + 1]
    int32_t n_seg;
    long long n_tiles;                 // work units (strips)
    int32_t T;                         // narrow: compare rows per tile; wide: rows per chunk
    int32_t R;                         // wide: compare rows per strip; narrow: tiles per strip
    int32_t stages;
    uint32_t stage_bytes;
    uint32_t warp_smem;                // shared-memory bytes per warp
    uint32_t off_bars, off_descs, off_stg, off_list;   // inside the warp's region
    uint32_t* scr;                     // scratch index rows (unordered blocks): {end, row << 16 | order} x scr_cap
                                       // (row = BED start - the unit's unit_pos0, SCR_ROW_CHR = a chr-end row:
                                       //  BED start = the unit's unit_aux)
    long long out_cap;                 // 0 = count only
    long long scr_cap;                 // rows of the scratch area
    uint32_t chunk;                    // scratch rows a warp reserves per atomicAdd
    uint32_t* tile_cnt;                // [n_tiles] index rows of the unit
    unsigned long long* tile_off;      // [n_tiles] scratch row of the unit's (first) block
    // wide: a strip whose output crosses scratch chunks continues in further blocks
    uint32_t* unit_pos0;               // [n_tiles] position the unit's row numbers count from
    uint32_t* unit_aux;                // [n_tiles] BED start of the unit's chr-end rows (the record length)
    uint32_t* first_cnt;               // [n_tiles] rows of the first block
    int32_t* unit_next;                // [n_tiles] next block (index into pool) or -1
    BlockRec* pool;                    // [pool_cap] further blocks
    uint32_t pool_cap;
    unsigned int* pool_counter;
    unsigned long long* cursor;        // scratch allocation cursor
    unsigned long long* strip_counter; // wide: next strip to hand out
    int32_t prefetch;                  // wide: prefetch a strip's next chunk into L2 (evict_last; the copies are
                                       // evict_first either way)
    int64_t* result;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared (TMA engine), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// the same range on its way into L2 only (no shared-memory destination, no completion)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// ... with L2 eviction priorities: the prefetched lines are kept (evict_last) until the copy that
// consumes them marks them evict_first
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_prefetch_l2_hint(const void* src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
        : "memory");
}


// one scratch row: MEM end + (row << 16 | order), row = BED start - the unit's pos0 or SCR_ROW_CHR
__device__ __forceinline__ void scr_store(uint32_t* dst, uint32_t end, uint32_t row, uint32_t order) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(end, (row << 16) | order);
}

// Output of one work unit (strip) of a warp.  A warp reserves P.chunk rows of the
// scratch area with one atomicAdd and appends the index rows of its strips to
// them, so that the single allocation cursor sees one atomic per few thousand
// index rows.  A strip's output is therefore one block of consecutive scratch
// rows, or -- when it does not fit the rest of the warp's chunk -- a chain of
// blocks: the first one is described by the strip's own tile_off / first_cnt /
// unit_next entries, later ones by BlockRec pool records.  A request that does
// not fit abandons the rest of the chunk (fewer rows than the request itself, so
// at most half of the reserved rows are ever wasted).  All members are warp
// uniform; every method must be called by the whole warp.
struct StripOut {
    uint32_t* w_ptr = nullptr;                     // next scratch row of the warp's chunk
    uint32_t w_left = 0;                           // rows left in the chunk
    unsigned long long w_end = 0;                  // scratch row index after the chunk
    unsigned long long blk_off = 0;                // open block of the strip
    uint32_t blk_left0 = 0;                        // w_left when the open block began
    uint32_t total = 0;                            // rows of the strip's closed blocks
    int nblk = 0, last_rec = -1;                   // blocks closed so far; pool index of the last one
    long long strip = 0;

    __device__ __forceinline__ void begin(long long strip_id) {
        strip = strip_id;
        nblk = 0;
        blk_off = w_end - w_left;
        blk_left0 = w_left;
        total = 0;
    }
    __device__ __forceinline__ void close_block(const FastParams& P, int lane) {
        const uint32_t blk_cnt = blk_left0 - w_left;
        total += blk_cnt;
        if (nblk == 0) {
            if (lane == 0) {
                P.tile_off[strip] = blk_off;
                P.first_cnt[strip] = blk_cnt;
                P.unit_next[strip] = -1;
            }
        } else {
            int idx = 0;
            if (lane == 0) {
                idx = (int)atomicAdd(P.pool_counter, 1u);
                if ((uint32_t)idx < P.pool_cap) {
                    BlockRec rec;
                    rec.off = blk_off; rec.cnt = blk_cnt; rec.next = -1;
                    P.pool[idx] = rec;
                    if (nblk == 1) P.unit_next[strip] = idx; else P.pool[last_rec].next = idx;
                }
            }
            last_rec = __shfl_sync(FULL, idx, 0);
        }
        ++nblk;
    }
    // next chunk: the strip continues in a new block
    __device__ __forceinline__ void new_chunk(const FastParams& P, uint32_t n, int lane) {
        close_block(P, lane);
        const unsigned long long want = n > P.chunk ? n : P.chunk;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.cursor, want);
        base = __shfl_sync(FULL, base, 0);
        w_left = (uint32_t)want;
        w_end = base + want;
        // a chunk past the end of the scratch area writes to the dump rows behind it (its rows are
        // counted, the gather skips them, the caller retries with a larger out_cap)
        w_ptr = P.scr + (w_end <= (unsigned long long)P.scr_cap ? base : (unsigned long long)P.scr_cap) * SCR_WORDS;
        blk_off = base;
        blk_left0 = w_left;
    }
    // `n` consecutive scratch rows (8-byte records, scr_record) for the strip's next index
    // rows (n <= SCR_DUMP_ROWS); returns the first.  Always writable: see new_chunk.
    __device__ __forceinline__ uint32_t* reserve(const FastParams& P, uint32_t n, int lane) {
        if (n > w_left) new_chunk(P, n, lane);
        uint32_t* const at = w_ptr;
        w_ptr += SCR_WORDS * n;
        w_left -= n;
        return at;
    }
    // pos0: the position the strip's row numbers count from; aux: BED start of its chr-end rows
    __device__ __forceinline__ void end(const FastParams& P, int lane, uint32_t pos0, uint32_t aux) {
        close_block(P, lane);
        if (lane == 0) {
            P.tile_cnt[strip] = total;
            P.unit_pos0[strip] = pos0;
            P.unit_aux[strip] = aux;
        }
    }
};

typedef void (*stream_kernel_t)(const FastParams);

}  // namespace

// index_wide.cu: strip kernel for any n_cols <= 512 (nullptr beyond); *kpl = sorted
// positions per lane.
stream_kernel_t select_wide_kernel(int n_cols, bool order, int* kpl);

// index_wide2.cu: the single-kernel strip build (TMA tensor staging, phase A row scan,
// decoupled look-back straight into the ordered output) for rows of up to 256 columns
bool wide2_supported(int32_t n_cols, int32_t ld);
size_t wide2_workspace_bytes(int64_t rows, int32_t n_cols, int32_t ld, int64_t out_cap, const memo_segment_t* segs,
                             int32_t n_seg, const memo_index_opts_t* opts);
int launch_wide2(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld, const memo_segment_t* segs,
                 int32_t n_seg, const memo_index_opts_t* opts, int32_t* out_start, uint32_t* out_end,
                 int32_t* out_order, int64_t out_cap, int64_t* seg_out_end, int64_t* result,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream);

// index_build.cu: CUDA events around the streaming kernel while memo_profile_enable is on
void profile_begin(cudaStream_t stream);
void profile_end(cudaStream_t stream);

// index_narrow.cu: kernel for n_cols == ld == CT (compile-time) rows, or nullptr.
// *rows_per_lane receives the number of consecutive rows a lane scans per step
// (tile bases must be multiples of it so that they are 16-byte aligned).
stream_kernel_t select_narrow_kernel(int n_cols, bool order, int* rows_per_lane);

}  // namespace memo
