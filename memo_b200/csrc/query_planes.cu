// k-mer conservation (n_docs <= 255) and membership queries on sm_100a: bit-plane tiles.
//
// Replaces the reference's src/memo_query.py memo_init :42-55, memo_query :57-63
// and the argmax of print_res :70 (see query.cu for the definition):
//
//   conservation[p] = min{ f3 : clip(f2 - s - (k-1)) <= p < clip(f1 - s) }, else n_docs
//
// The reference paints a [W, N+1] byte matrix, k-1 byte stores per index row.  A
// shared-memory atomic per painted cell is what bounds a direct transcription
// (ATOMS retires ~2 lanes per clock per SM).  Here a tile of TP window positions
// is kept as one BIT-PLANE per order value: plane o has bit p set iff some row
// with f3 == o covers position p.  A row covers at most k-1 consecutive
// positions, i.e. it ORs one or two 32-bit words (k <= 33) instead of painting
// 30 cells.  When the tile's rows are in, one lane per 32-position word column
// walks the planes in ascending order (first plane with the bit set = the
// minimum), keeps the result as B binary digit planes (n_docs < 2^B) and expands those
// into the 32 output bytes of its positions: two 128-bit coalesced stores per lane.
//
// Membership (-m) uses the same planes (one per genome) and row walk; its result IS a
// bitmap, transposed: one lane per word column transposes 32 planes x 32 positions in
// registers, the words go through the (already consumed) plane storage and out as rows
// of ceil(n_docs/32) words per position.
//
// Every WARP is an independent stream (no __syncthreads, no bounds pass): it
// takes runs of consecutive tiles from an atomic counter, finds the first row of
// the run with one cooperative 32-ary search and from then on walks the rows
// forward 128 at a time (one 128-bit load per array and lane, next batch in
// flight) -- the first row of the next tile is seen while the current tile's halo
// rows go by.  A tile that turns out heavy (a dense stretch of the index, e.g. the
// start of a record) is handed on in pieces through a small record table in the
// workspace, so that no single warp decides the kernel's time.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace memo {
namespace {

constexpr int QP_WARPS = 4;          // warps per CTA
constexpr int QP_WORDS = 1280;       // shared-memory words per warp (all planes of its tile)
constexpr int QP_WORDS_M = 2560;     // ... membership (needs 32 more for the skewed read-out)
constexpr int QP_HEAVY = 2048;       // rows after which a tile is handed on in (up to 8) pieces
constexpr int QP_QCAP = 256;         // heavy tiles that can be handed on per launch (the others are walked as they are)
constexpr size_t QP_WS_HEADER = 256; // workspace: next run | 128 bytes on: number of heavy tiles | their records

struct HeavyTile;

struct PlaneParams {
    const int32_t* f1;
    const uint32_t* f2;
    const int32_t* f3;
    long long n_rows, s, W;
    int n_docs;
    int n_k;                       // k values answered by this launch (a sweep shares the tile hand-out,
    int ks[16];                    //  the row search and the rows in L1/L2); result i at out + i * out_stride
    long long out_stride;          // bytes between the results of consecutive k values
    int NOP;                       // planes per tile: n_docs rounded up to a multiple of the group size
    int WPT;                       // words per plane
    int TP;                        // positions per tile = 32 WPT
    int run;                       // tiles per run
    int heavy_rows;                // rows after which a tile is handed on in pieces
    long long n_tiles, n_runs;
    void* out;                     // uint8 [W] | membership: uint32 [W, NOP / 32]
    int32_t* status;
    unsigned long long* counter;   // next run
    unsigned int* n_heavy;         // heavy tiles handed on so far (its own 128-byte line)
    HeavyTile* heavy;              // [QP_QCAP]
};

struct __align__(16) HeavyTile {   // a tile handed on in pieces (32 bytes, zero = not written yet)
    // (window position + 1) << 28 | positions << 12 | words per piece << 4 | pieces (<= 8);
    // readers wait for it to become non-zero
    unsigned long long desc;
    unsigned int next;             // next piece to take (may run past the number of pieces)
    unsigned int pad[5];
};

// first row with f1 > key (n if none): 32 probes per round
__device__ __forceinline__ long long warp_upper_bound(const int32_t* __restrict__ f1, long long n,
                                                      long long key, int lane) {
    long long lo = 0, hi = n;                       // the answer lies in [lo, hi]
    while (hi - lo > 32) {
        const long long span = hi - lo;
        // probe l at lo + span (l+1) / 32 - 1: ascending in l, first >= lo, last = hi - 1
        const long long idx = lo + ((span * (lane + 1)) >> 5) - 1;
        const unsigned b = __ballot_sync(FULL, (long long)f1[idx] > key);
        if (b == 0u) return hi;                      // row hi - 1 is not past the key
        const int j = __ffs(b) - 1;
        hi = lo + ((span * (j + 1)) >> 5) - 1;
        if (j > 0) lo = lo + ((span * j) >> 5);
    }
    const long long idx = lo + lane;
    const unsigned b = __ballot_sync(FULL, idx >= hi || (long long)f1[idx] > key);
    return b ? lo + (__ffs(b) - 1) : hi;
}

// 32 x 32 bit matrix transpose in registers: afterwards bit o of a[i] = former bit i of a[o]
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= t << j;
            a[k + j] ^= t;
        }
        m ^= m << (j >> 1);
    }
}

struct RowBatch {                  // 128 consecutive rows: rows 4 lane .. 4 lane + 3 of the batch
    int32_t a1[4];
    uint32_t a2[4];
    uint32_t a3[4];
};

// B = binary digits of the result that can be non-zero (n_docs < 2^B); GS = planes per
// unrolled group of the read-out (P.NOP is a multiple of it)
template <int B, int GS, bool MEMB>
__global__ void __launch_bounds__(QP_WARPS * 32) query_planes_kernel(const PlaneParams P) {
    constexpr int GB = GS == 8 ? 3 : 2;              // digits fixed by the position inside a group
    __shared__ __align__(16) uint32_t smem[QP_WARPS][MEMB ? QP_WORDS_M + 32 : QP_WORDS];
    // conservation read-out: byte v of a digit plane -> the 8 result bytes it contributes to
    // (bit i of v in byte i), looked up instead of spread by multiplies
    __shared__ uint2 spread_lut[MEMB ? 1 : 256];
    if (!MEMB) {
        for (int v = threadIdx.x; v < 256; v += QP_WARPS * 32)
            spread_lut[v] = make_uint2((((uint32_t)v & 0xFu) * 0x00204081u) & 0x01010101u,
                                       (((uint32_t)v >> 4) * 0x00204081u) & 0x01010101u);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    uint32_t* const planes = smem[threadIdx.x >> 5];
    const int WPT = P.WPT, n_docs = P.n_docs;
    const int n_words = P.NOP * WPT;
    uint32_t km1 = 0;                                // k - 1 of the k value being answered
    int halo = 0;                                    // rows starting up to tile end + halo can cover
    unsigned char* out_k = static_cast<unsigned char*>(P.out);
    const long long n_rows = P.n_rows;

    // One tile: window positions [t0, t0 + tn), rows from r on (the first row with f1 > s + t0).
    // Returns the row where the next tile starts (the first one past this tile), or -1 if the
    // tile turned out heavy and was handed on in pieces (splittable, more than 32 positions).
    // All arithmetic is 32-bit relative to the tile base (the window ends below 2^31 - 2^17:
    // launch_query_planes checks).  Rows are read 128 at a time from the 4-aligned row below r
    // (one 128-bit load per array and lane; f1, f2, f3 are 16-byte aligned): the up to three
    // rows before r have f1 <= base and drop out as "not in"; rows past the end of the index
    // read as f1 = INT_MAX.
    auto do_tile = [&](const long long t0, const int tn, const long long r, const bool splittable) -> long long {
        const int base = (int)(P.s + t0);
        const uint32_t bk = (uint32_t)base + km1;
        const uint32_t lim_in = (uint32_t)(tn + halo);
        const long long ra = r & ~3ll;
        const int32_t* const pf1 = P.f1 + ra;
        const uint32_t* const pf2 = P.f2 + ra;
        const int32_t* const pf3 = P.f3 + ra;
        const int n32 = (int)min(n_rows - ra, 0x40000000ll);
        auto load = [&](RowBatch& bt, int i) {
            const int idx = i + 4 * lane;
            if (idx + 3 < n32) {
                const int4 v1 = *reinterpret_cast<const int4*>(pf1 + idx);
                const uint4 v2 = *reinterpret_cast<const uint4*>(pf2 + idx);
                const int4 v3 = *reinterpret_cast<const int4*>(pf3 + idx);
                bt.a1[0] = v1.x; bt.a1[1] = v1.y; bt.a1[2] = v1.z; bt.a1[3] = v1.w;
                bt.a2[0] = v2.x; bt.a2[1] = v2.y; bt.a2[2] = v2.z; bt.a2[3] = v2.w;
                bt.a3[0] = (uint32_t)v3.x; bt.a3[1] = (uint32_t)v3.y; bt.a3[2] = (uint32_t)v3.z; bt.a3[3] = (uint32_t)v3.w;
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const bool ok = idx + h < n32;
                    bt.a1[h] = ok ? pf1[idx + h] : 0x7FFFFFFF;
                    bt.a2[h] = ok ? pf2[idx + h] : 0u;
                    bt.a3[h] = ok ? (uint32_t)pf3[idx + h] : 0u;
                }
            }
        };
        RowBatch cur, nxt;
        load(cur, 0);
        // ---- all planes := empty
        for (int i = lane * 4; i < n_words; i += 128)
            *reinterpret_cast<uint4*>(planes + i) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        // ---- rows with base < f1 <= base + tn + halo, in order
        int i = 0, n_before = 0;                       // rows walked; rows not past the tile end
        bool bad = false;
        for (;;) {
            load(nxt, i + 128);
            int c = 0;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int x = cur.a1[h] - base;
                const bool in = (uint32_t)(x - 1) < lim_in;          // 0 < x <= tn + halo
                c += x <= tn ? 1 : 0;
                // the row covers [a, b) of the tile: a = clip(f2 - base - (k-1)), b = clip(x)
                uint32_t a = min(cur.a2[h] - bk, (uint32_t)tn);
                if (cur.a2[h] < bk) a = 0u;
                const uint32_t b = (uint32_t)min(x, tn);
                const uint32_t ord = cur.a3[h];
                // (conservation: order n_docs is valid and means "not covered")
                bad = bad || (in && (MEMB ? ord >= (uint32_t)n_docs : ord > (uint32_t)n_docs));
                if (in && b > a && ord < (uint32_t)n_docs) {
                    uint32_t* const pl = planes + ord * WPT;
                    const uint32_t w0 = a >> 5, w1 = (b - 1u) >> 5;
                    const uint32_t m0 = 0xFFFFFFFFu << (a & 31u);
                    const uint32_t m1 = 0xFFFFFFFFu >> (31u - ((b - 1u) & 31u));
                    if (w0 == w1) {
                        atomicOr(pl + w0, m0 & m1);
                    } else {
                        atomicOr(pl + w0, m0);
                        for (uint32_t w = w0 + 1; w < w1; ++w) atomicOr(pl + w, 0xFFFFFFFFu);
                        atomicOr(pl + w1, m1);
                    }
                }
            }
            n_before += __reduce_add_sync(FULL, c);
            // rows are in f1 order: the batch's last row tells whether the tile's rows are through
            if (__shfl_sync(FULL, cur.a1[3], 31) - base > (int)lim_in) break;
            cur = nxt;
            i += 128;
            if (splittable && i == P.heavy_rows && tn > 32) {
                // a heavy tile (a dense stretch of the index): hand it on as up to 8 pieces of
                // whole 32-position words; a piece that is still heavy is split again
                const unsigned words = (unsigned)(tn + 31) >> 5;
                const unsigned piece = (words + 7u) >> 3;
                const unsigned n_sub = (words + piece - 1u) / piece;
                unsigned slot = 0xFFFFFFFFu;
                if (lane == 0) {
                    slot = atomicAdd(P.n_heavy, 1u);
                    if (slot < (unsigned)QP_QCAP) {
                        *reinterpret_cast<volatile unsigned long long*>(&P.heavy[slot].desc) =
                            (((unsigned long long)t0 + 1ull) << 28) | ((unsigned long long)tn << 12) |
                            (unsigned long long)(piece << 4) | n_sub;
                    }
                }
                slot = __shfl_sync(FULL, slot, 0);
                if (slot < (unsigned)QP_QCAP) return -1;
            }
        }
        if (bad) *P.status = 1;
        __syncwarp();
        if (MEMB) {
            // ---- planes -> rows of NW words per position.  One group of 32 genomes at a time,
            //      last group first: lane wc transposes the 32 x 32 bits of its word column and
            //      parks the 32 words in the group's own (now consumed) plane storage, skewed by
            //      one word per column (conflict free; the skew spills into the storage of the
            //      group above, which is done with); then the words go out in position order.
            const int NW = P.NOP >> 5;
            const uint32_t last_mask = (n_docs & 31) ? ((1u << (n_docs & 31)) - 1u) : 0xFFFFFFFFu;
            uint32_t* const ob = reinterpret_cast<uint32_t*>(out_k) + t0 * NW;
            const bool act = lane < WPT && 32 * lane < tn;
            for (int g = NW - 1; g >= 0; --g) {
                uint32_t* const reg = planes + 32 * g * WPT;
                uint32_t A[32];
#pragma unroll
                for (int o = 0; o < 32; ++o) A[o] = act ? reg[o * WPT + lane] : 0u;
                __syncwarp();
                transpose32(A);
                const uint32_t gm = g == NW - 1 ? last_mask : 0xFFFFFFFFu;
                if (act) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) reg[lane * 33 + q] = ~A[q] & gm;
                }
                __syncwarp();
                {
                    uint32_t* dst = ob + lane * NW + g;
                    const uint32_t* src = reg + lane;
                    for (int u = lane; u < tn; u += 32, dst += 32 * NW, src += 33) *dst = *src;
                }
                __syncwarp();
            }
            return ra + (n_before < n32 ? n_before : n32);
        }
        // ---- one lane per 32-position word column: the first plane with the bit set is the
        //      minimum; kept as B binary digit planes.  Few word columns (many planes): the
        //      planes are split between the two half-warps and merged afterwards.
        const int halves = WPT <= 16 ? 2 : 1;
        const int lpi = 32 / halves;                       // word columns per sweep
        const int half = halves == 2 ? lane >> 4 : 0;
        const int g_cnt = P.NOP / GS / halves;             // plane groups per lane
        for (int wc0 = 0; 32 * wc0 < tn; wc0 += lpi) {
            const int wc = wc0 + (lane & (lpi - 1));
            uint32_t seen = 0u, dg[B];
#pragma unroll
            for (int bb = 0; bb < B; ++bb) dg[bb] = 0u;
            // (idle lanes of a short sweep read the last column)
            const uint32_t* col = planes + min(wc, WPT - 1) + half * g_cnt * GS * WPT;
            for (int g = half * g_cnt; g < (half + 1) * g_cnt; ++g, col += GS * WPT) {
                const uint32_t before = seen;
#pragma unroll
                for (int o = 0; o < GS; ++o) {
                    const uint32_t M = col[o * WPT];
                    const uint32_t S = M & ~seen;          // positions whose minimum is GS g + o
                    seen |= M;
#pragma unroll
                    for (int bb = 0; bb < GB && bb < B; ++bb)
                        if (o & (1 << bb)) dg[bb] |= S;
                }
                const uint32_t U = seen & ~before;          // positions decided in this group
#pragma unroll
                for (int bb = GB; bb < B; ++bb) dg[bb] |= ((g >> (bb - GB)) & 1) ? U : 0u;
            }
            if (halves == 2) {                             // the upper planes count where no lower one does
                const uint32_t seen_hi = __shfl_down_sync(FULL, seen, 16);
#pragma unroll
                for (int bb = 0; bb < B; ++bb) dg[bb] |= __shfl_down_sync(FULL, dg[bb], 16) & ~seen;
                seen |= seen_hi;
            }
            if (half != 0 || 32 * wc >= tn) continue;
            const uint32_t none = ~seen;                   // nothing covers: n_docs
#pragma unroll
            for (int bb = 0; bb < B; ++bb) dg[bb] |= ((n_docs >> bb) & 1) ? none : 0u;
            // digit planes -> bytes: output word j holds positions 4j .. 4j+3
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                  // byte j of every digit plane: positions 8j .. 8j+7
                uint32_t lo = 0u, hi = 0u;
#pragma unroll
                for (int bb = 0; bb < B; ++bb) {
                    const uint2 sp = spread_lut[(dg[bb] >> (8 * j)) & 0xFFu];
                    lo += sp.x << bb;                       // (digits do not overlap: add = or)
                    hi += sp.y << bb;
                }
                w[2 * j] = lo;
                w[2 * j + 1] = hi;
            }
            uint8_t* const o = out_k + t0 + 32 * wc;
            if (32 * wc + 32 <= tn) {
                reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
                const int n = tn - 32 * wc;
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (4 * j + q < n) o[4 * j + q] = (uint8_t)(w[j] >> (8 * q));
            }
        }
        __syncwarp();
        return ra + (n_before < n32 ? n_before : n32);
    };

    // takes one piece of a heavy tile, if there is one left: its window position.  Every warp
    // walks the records in order and leaves a record for good once its pieces are gone, so a
    // record sees at most one failed take per warp.
    unsigned hv_cur = 0, hv_seen = 0;                // hv_seen: number of records, read one run ahead
    auto ld_heavy = [&]() -> unsigned {
        return *reinterpret_cast<const volatile unsigned int*>(P.n_heavy);
    };
    auto take_piece = [&](bool fresh, int& tn_out) -> long long {
        unsigned nh = fresh ? ld_heavy() : __shfl_sync(FULL, hv_seen, 0);       // (read by lane 0)
        if (nh > (unsigned)QP_QCAP) nh = QP_QCAP;
        while (hv_cur < nh) {
            // 32 records per look: one 16-byte read of {desc, next} each (a same-address atomic
            // per warp and record would cost more than the pieces are worth)
            const unsigned idx = hv_cur + lane;
            unsigned long long desc = 0, nx = 0;
            if (idx < nh) {
                do {
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(desc), "=l"(nx) : "l"(P.heavy + idx));
                } while (desc == 0ull);
            }
            const unsigned live = __ballot_sync(FULL, idx < nh && (unsigned)nx < (unsigned)(desc & 0xFull));
            if (live == 0u) {
                hv_cur = min(hv_cur + 32u, nh);
                continue;
            }
            const int src = __ffs(live) - 1;                 // the records before it are exhausted
            unsigned j = 0;
            if (lane == src) j = atomicAdd(&P.heavy[idx].next, 1u);
            j = __shfl_sync(FULL, j, src);
            desc = __shfl_sync(FULL, desc, src);
            hv_cur += src;
            if (j < (unsigned)(desc & 0xFull)) {
                const int plen = 32 * (int)((desc >> 4) & 0xFFull);
                const int total = (int)((desc >> 12) & 0xFFFFull);
                tn_out = min(plen, total - plen * (int)j);
                return (long long)(desc >> 28) - 1 + (long long)plen * j;
            }
            ++hv_cur;                                        // lost the last piece to another warp
        }
        return -1;
    };

    unsigned long long look = 0;
    if (lane == 0) look = atomicAdd(P.counter, 1ull);
    long long t_cur = 0, t_end = 0, r_run = 0;       // the tiles left of the warp's run
    bool finishing = false;                          // no runs left: only pieces
    for (;;) {
        long long t0, r;
        int tn;
        const bool from_run = t_cur < t_end;
        if (from_run) {
            t0 = t_cur * P.TP;
            tn = (int)min((long long)P.TP, P.W - t0);
            r = r_run >= 0 ? r_run : warp_upper_bound(P.f1, n_rows, P.s + t0, lane);
        } else {
            // between runs: pieces of heavy tiles go first.  Whoever queued pieces comes by
            // here afterwards, so none is left behind.
            t0 = take_piece(finishing, tn);
            if (t0 >= 0) {
                r = warp_upper_bound(P.f1, n_rows, P.s + t0, lane);
            } else if (finishing) {
                break;
            } else {
                const long long run = (long long)__shfl_sync(FULL, look, 0);
                if (run >= P.n_runs) {
                    finishing = true;
                } else {
                    if (lane == 0) {
                        look = atomicAdd(P.counter, 1ull);              // one run ahead: hides the latency
                        hv_seen = ld_heavy();
                    }
                    t_cur = run * P.run;
                    t_end = min(t_cur + (long long)P.run, P.n_tiles);
                    r_run = -1;
                }
                continue;
            }
        }
        // every k of the sweep over the same tile and rows (a heavy tile is handed on at the first
        // k: its pieces answer all of them)
        long long r_next = -1;
        for (int ki = 0; ki < P.n_k; ++ki) {
            const int k = P.ks[ki];
            km1 = (uint32_t)(k - 1);
            halo = (k > 2 ? k : 2) - 2;
            out_k = static_cast<unsigned char*>(P.out) + (long long)ki * P.out_stride;
            r_next = do_tile(t0, tn, r, ki == 0);
            if (r_next < 0) break;
        }
        if (from_run) {
            r_run = r_next;
            ++t_cur;
        }
    }
}

typedef void (*planes_kernel_t)(const PlaneParams);

}  // namespace

size_t query_planes_workspace_bytes() { return QP_WS_HEADER + sizeof(HeavyTile) * QP_QCAP; }

// Largest n_docs the membership variant takes (one word column of all planes must fit)
int query_planes_max_membership_docs() { return QP_WORDS_M; }

int launch_query_planes(int membership, const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                        int64_t n_rows, int64_t q_start, int64_t q_end, const int32_t* ks, int32_t n_k,
                        int32_t n_docs, void* out, int64_t out_stride, int32_t* status, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
    PlaneParams P;
    P.f1 = f1; P.f2 = f2; P.f3 = f3;
    P.n_rows = n_rows; P.s = q_start; P.W = q_end - q_start;
    P.n_docs = n_docs;
    MEMO_REQUIRE(n_k >= 1 && n_k <= 16, "1 .. 16 k values per launch");
    P.n_k = n_k;
    for (int i = 0; i < 16; ++i) P.ks[i] = i < n_k ? ks[i] : 0;
    P.out_stride = out_stride;
    int wpt;
    if (membership) {
        P.NOP = (n_docs + 31) / 32 * 32;                 // groups of 32 genomes = output words
        wpt = QP_WORDS_M / P.NOP;
        if (wpt > 32) wpt = 32;                          // one lane per word column
        if (wpt < 1) {
            set_error("n_docs = %d too large for the membership planes", n_docs);
            return MEMO_ERR_UNSUPPORTED;
        }
    } else {
        const int gs = n_docs <= 16 ? 4 : 8;
        P.NOP = (n_docs + gs - 1) / gs * gs;
        if (QP_WORDS / P.NOP <= 16) P.NOP = (n_docs + 2 * gs - 1) / (2 * gs) * (2 * gs);   // halves of the planes
        wpt = QP_WORDS / P.NOP;
        if (wpt >= 32) wpt = wpt / 32 * 32;             // whole warps of word columns
        if (wpt > 128) wpt = 128;
    }
    P.WPT = wpt;
    P.TP = 32 * wpt;
    P.n_tiles = (P.W + P.TP - 1) / P.TP;
    P.out = out; P.status = status;
    // a tile is heavy when it has walked many times the average tile's rows
    P.heavy_rows = QP_HEAVY;
    if (P.n_tiles > 0 && 3 * (n_rows / P.n_tiles) > P.heavy_rows)
        P.heavy_rows = (int)min(3 * (n_rows / P.n_tiles), (long long)(1 << 30)) / 128 * 128;
    static const int env_heavy = getenv("MEMO_QUERY_HEAVY") ? atoi(getenv("MEMO_QUERY_HEAVY")) : -1;   // tuning: 0 = never
    static const int env_run = getenv("MEMO_QUERY_RUN") ? atoi(getenv("MEMO_QUERY_RUN")) : 0;
    if (env_heavy >= 0) P.heavy_rows = env_heavy / 128 * 128;
    const size_t need = query_planes_workspace_bytes();
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return MEMO_ERR_WORKSPACE;
    }
    P.counter = static_cast<unsigned long long*>(workspace);
    P.n_heavy = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 128);
    P.heavy = reinterpret_cast<HeavyTile*>(static_cast<char*>(workspace) + QP_WS_HEADER);
    MEMO_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
    int bits = 1;
    while ((1 << bits) <= n_docs) ++bits;            // n_docs < 2^bits
    planes_kernel_t fn;
    int slot;
    if (membership) { fn = query_planes_kernel<1, 4, true>; slot = 6; }
    else if (bits <= 3) { fn = query_planes_kernel<3, 4, false>; slot = 0; }
    else if (bits <= 4) { fn = query_planes_kernel<4, 4, false>; slot = 1; }
    else if (n_docs <= 16) { fn = query_planes_kernel<5, 4, false>; slot = 2; }
    else if (bits <= 6) { fn = query_planes_kernel<6, 8, false>; slot = 3; }
    else if (bits <= 7) { fn = query_planes_kernel<7, 8, false>; slot = 4; }
    else { fn = query_planes_kernel<8, 8, false>; slot = 5; }
    // resident CTAs per SM of the variant on the current device (one query per device and variant)
    int ctas = 0;
    {
        static std::mutex mu;
        static int per_sm[64][7];
        int dev = 0;
        MEMO_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || per_sm[dev][slot] == 0) {
            int n = 0;
            MEMO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, QP_WARPS * 32, 0));
            ctas = n > 0 ? n : 1;
            if (dev >= 0 && dev < 64) per_sm[dev][slot] = ctas;
        } else {
            ctas = per_sm[dev][slot];
        }
    }
    const long long warps = (long long)device_sm_count() * ctas * QP_WARPS;
    // runs of consecutive tiles share one row search; about 3 runs per warp keep the tail short
    long long run = P.n_tiles / (warps * 3);
    P.run = (int)(run < 1 ? 1 : run > 16 ? 16 : run);
    if (env_run > 0) P.run = env_run;                // tuning
    P.n_runs = (P.n_tiles + P.run - 1) / P.run;
    long long grid = (long long)device_sm_count() * ctas;
    const long long enough = (P.n_runs + QP_WARPS - 1) / QP_WARPS;
    if (grid > enough) grid = enough;
    fn<<<(unsigned)grid, QP_WARPS * 32, 0, stream>>>(P);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // namespace memo
