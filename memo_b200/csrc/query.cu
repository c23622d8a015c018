// k-mer conservation / membership query over a pivot window on sm_100a.
//
// Replaces the reference's src/memo_query.py: memo_init :42-55 (re-centre,
// shadow cast by k-1, clip, keep end < start), memo_query :57-63 (paint) and
// the argmax of print_res :70.  Per window position p (relative to q_start):
//
//   conservation[p] = min{ f3 : clip(f2 - s - (k-1)) <= p < clip(f1 - s) }, else n_docs
//   membership[p]   = all-ones mask with bit f3 cleared for every such row
//
// Only rows with p < f1 - s <= p + k - 1 can cover p (f2 >= f1), so a tile of T
// positions needs the rows with f1 in (s + t0, s + t0 + T + k - 2]; rows are in
// index order (f1 ascending), so that is one contiguous range found by binary
// search (bounds kernel).  The tile lives in shared memory; painting uses
// shared-memory atomics, the result is written once, coalesced.
#include <stdlib.h>

#include "common.cuh"

namespace memo {
namespace {

constexpr int QT_MIN = 2048;   // smallest tile (window positions): the bounds workspace is sized for it
constexpr int QTHREADS = 256;

// lo[t] = first row with f1 > s + t*QT              (t = 0..n_tiles-1)
// hi[t] = first row with f1 > s + (t+1)*QT + k - 2
__global__ void query_bounds_kernel(const int32_t* __restrict__ f1, long long n_rows, long long s,
                                    int k, int QT, long long n_tiles, long long* __restrict__ lo,
                                    long long* __restrict__ hi) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const long long key_lo = s + t * (long long)QT;
    const long long key_hi = s + (t + 1) * (long long)QT + k - 2;
    long long a = 0, b = n_rows;
    while (a < b) {
        const long long m = (a + b) >> 1;
        if ((long long)f1[m] > key_lo) b = m; else a = m + 1;
    }
    lo[t] = a;
    b = n_rows;                                   // hi >= lo
    while (a < b) {
        const long long m = (a + b) >> 1;
        if ((long long)f1[m] > key_hi) b = m; else a = m + 1;
    }
    hi[t] = a;
}

template <typename OutT, int QT, int TH>
__global__ void __launch_bounds__(TH)
query_conservation_kernel(const int32_t* __restrict__ f1, const uint32_t* __restrict__ f2,
                          const int32_t* __restrict__ f3, const long long* __restrict__ lo,
                          const long long* __restrict__ hi, long long s, long long W, int k,
                          int n_docs, OutT* __restrict__ out, int32_t* status) {
    __shared__ uint32_t tile[QT];
    const long long n_tiles = (W + QT - 1) / QT;
    long long t = blockIdx.x;
    if (t >= n_tiles) return;
    long long r0 = lo[t], r1 = hi[t];
    for (; t < n_tiles; t += gridDim.x) {
        const long long t0 = t * QT;
        const long long t1 = min(t0 + (long long)QT, W);
        // the tile's first batch of rows and the next tile's row range are requested before
        // the tile is initialised: their latency overlaps the shared-memory work
        long long r = r0 + threadIdx.x;
        const bool have = r < r1;
        const int32_t v1 = have ? f1[r] : 0;
        const uint32_t v2 = have ? f2[r] : 0u;
        const int32_t v3 = have ? f3[r] : 0;
        const long long tn = t + gridDim.x;
        const long long nr0 = tn < n_tiles ? lo[tn] : 0, nr1 = tn < n_tiles ? hi[tn] : 0;
        for (int i = threadIdx.x; i < QT; i += TH) tile[i] = (uint32_t)n_docs;
        __syncthreads();
        // Row r paints [cend, start).  If the row before it starts at the same position, has
        // an order <= this one and ends no earlier, it covers the tail [its cend, start) with
        // a value at least as small: this row then only paints up to that cend.  (Rows of one
        // position come in ascending order with descending ends, so a position's rows paint
        // every window position once instead of once per row.)
        auto paint = [&](long long r, int32_t a1, uint32_t a2, int32_t ord) {
            const long long start = (long long)a1 - s;                    // > t0 by construction
            const long long cend = (long long)a2 - s - (k - 1);
            if (ord < 0 || ord > n_docs) { *status = 1; return; }
            const long long a = max(cend, t0);                            // clip to [0, W] and to the tile
            long long b = min(start, t1);
            if (r > 0 && b > a) {
                const int32_t p1 = f1[r - 1], p3 = f3[r - 1];
                const uint32_t p2 = f2[r - 1];
                if (p1 == a1 && p3 >= 0 && p3 <= ord && p2 >= a2) b = min(b, (long long)p2 - s - (k - 1));
            }
            for (long long q = a; q < b; ++q) atomicMin(&tile[q - t0], (uint32_t)ord);
        };
        if (have) paint(r, v1, v2, v3);
        // four rows per thread in flight
        for (r += TH; r < r1; r += 4 * TH) {
            int32_t w1[4], w3[4];
            uint32_t w2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long rj = r + j * TH;
                const bool ok = rj < r1;
                w1[j] = ok ? f1[rj] : 0;
                w2[j] = ok ? f2[rj] : 0u;
                w3[j] = ok ? f3[rj] : 0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r + j * TH < r1) paint(r + j * TH, w1[j], w2[j], w3[j]);
        }
        __syncthreads();
        const int n = (int)(t1 - t0);
        if (sizeof(OutT) == 1 && n == QT) {
            // 16 positions per thread -> one 128-bit store
            uint4* dst = reinterpret_cast<uint4*>(out + t0);
            for (int i = threadIdx.x; i < QT / 16; i += TH) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 v = reinterpret_cast<const uint4*>(tile)[i * 4 + j];
                    w[j] = (v.x & 0xFF) | ((v.y & 0xFF) << 8) | ((v.z & 0xFF) << 16) | ((v.w & 0xFF) << 24);
                }
                dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += TH) out[t0 + i] = (OutT)tile[i];
        }
        __syncthreads();
        r0 = nr0;
        r1 = nr1;
    }
}

// Membership: NW = ceil(n_docs/32) words per position; tile of QM positions.
constexpr int QM_WORDS = 8192;   // shared-memory words per tile

__global__ void __launch_bounds__(QTHREADS)
query_membership_kernel(const int32_t* __restrict__ f1, const uint32_t* __restrict__ f2,
                        const int32_t* __restrict__ f3, long long n_rows, long long s, long long W,
                        int k, int n_docs, int NW, int TP, uint32_t* __restrict__ out,
                        int32_t* status) {
    __shared__ uint32_t tile[QM_WORDS];
    __shared__ long long range[2];
    const long long n_tiles = (W + TP - 1) / TP;
    const uint32_t last_mask = (n_docs % 32) ? ((1u << (n_docs % 32)) - 1u) : 0xFFFFFFFFu;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long t0 = t * TP;
        const long long t1 = min(t0 + (long long)TP, W);
        const int n = (int)(t1 - t0);
        for (int i = threadIdx.x; i < n * NW; i += QTHREADS)
            tile[i] = ((i % NW) == NW - 1) ? last_mask : 0xFFFFFFFFu;
        if (threadIdx.x < 2) {
            // rows with f1 in (s + t0, s + t1 + k - 2]
            const long long key = threadIdx.x == 0 ? (s + t0) : (s + t1 + k - 2);
            long long a = 0, b = n_rows;
            while (a < b) {
                const long long m = (a + b) >> 1;
                if ((long long)f1[m] > key) b = m; else a = m + 1;
            }
            range[threadIdx.x] = a;
        }
        __syncthreads();
        const long long r0 = range[0], r1 = range[1];
        for (long long r = r0 + threadIdx.x; r < r1; r += QTHREADS) {
            const long long start = (long long)f1[r] - s;
            const long long cend = (long long)f2[r] - s - (k - 1);
            const int32_t ord = f3[r];
            if (ord < 0 || ord >= n_docs) { *status = 1; continue; }
            const long long a = max(cend, t0);
            const long long b = min(start, t1);
            const uint32_t clr = ~(1u << (ord & 31));
            for (long long q = a; q < b; ++q) atomicAnd(&tile[(q - t0) * NW + (ord >> 5)], clr);
        }
        __syncthreads();
        uint32_t* dst = out + t0 * NW;
        for (int i = threadIdx.x; i < n * NW; i += QTHREADS) dst[i] = tile[i];
        __syncthreads();
    }
}

}  // namespace
}  // namespace memo

extern "C" {

size_t memo_query_workspace_bytes(int64_t window_len) {
    if (window_len < 0) window_len = 0;
    const int64_t n_tiles = (window_len + memo::QT_MIN - 1) / memo::QT_MIN;
    const size_t tiles = memo::align_up(sizeof(long long) * (size_t)(n_tiles + 1), 256) * 2;
    const size_t planes = memo::query_planes_workspace_bytes();
    return tiles > planes ? tiles : planes;
}

int memo_query_conservation(const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                            int64_t n_rows, int64_t q_start, int64_t q_end, int32_t k,
                            int32_t n_docs, void* out, int32_t out_u16, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(q_end >= q_start && q_start >= 0, "bad window [%lld, %lld)", (long long)q_start, (long long)q_end);
    MEMO_REQUIRE(k >= 1, "k must be >= 1");
    MEMO_REQUIRE(n_docs >= 1 && (out_u16 ? n_docs <= 65535 : n_docs <= 255),
                 "n_docs = %d does not fit the output type", n_docs);
    MEMO_REQUIRE(n_rows >= 0 && status != nullptr, "bad rows/status");
    const long long W = q_end - q_start;
    MEMO_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    if (W == 0) return MEMO_OK;
    MEMO_REQUIRE(out != nullptr, "out must not be NULL");
    MEMO_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
    // uint8 results (n_docs <= 255): bit-plane tiles, one stream per warp (query_planes.cu)
    if (!out_u16 && q_end + k < (1ll << 31) - (1ll << 17) &&
        !(getenv("MEMO_QUERY_PLANES") && atoi(getenv("MEMO_QUERY_PLANES")) == 0) &&
        ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) | reinterpret_cast<uintptr_t>(f3)) & 15) == 0)
        return launch_query_planes(0, f1, f2, f3, n_rows, q_start, q_end, &k, 1, n_docs, out, 0, status, workspace,
                                   workspace_bytes, stream);
    // tiles of 8192 positions; dense indexes (many rows per position) get smaller tiles so
    // that the work per tile stays small against the number of tiles per CTA
    constexpr int TH = QTHREADS;
    const int QT = (n_rows > W / 2) ? 2048 : 8192;
    const long long n_tiles = (W + QT - 1) / QT;
    const size_t half = align_up(sizeof(long long) * (size_t)(n_tiles + 1), 256);
    if (workspace == nullptr || workspace_bytes < 2 * half) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, 2 * half);
        return MEMO_ERR_WORKSPACE;
    }
    long long* lo = static_cast<long long*>(workspace);
    long long* hi = reinterpret_cast<long long*>(static_cast<char*>(workspace) + half);
    query_bounds_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(f1, n_rows, q_start, k, QT, n_tiles, lo, hi);
    MEMO_LAUNCH_CHECK(1);
    const int sms = device_sm_count();
    long long grid = (long long)sms * (QT == 8192 ? 5 : 8);     // by shared memory / by threads
    if (grid > n_tiles) grid = n_tiles;
#define MEMO_QLAUNCH(TT, QQ)                                                                \
    query_conservation_kernel<TT, QQ, TH><<<(unsigned)grid, TH, 0, stream>>>(               \
        f1, f2, f3, lo, hi, q_start, W, k, n_docs, static_cast<TT*>(out), status)
    if (out_u16) { if (QT == 8192) MEMO_QLAUNCH(uint16_t, 8192); else MEMO_QLAUNCH(uint16_t, 2048); }
    else         { if (QT == 8192) MEMO_QLAUNCH(uint8_t, 8192); else MEMO_QLAUNCH(uint8_t, 2048); }
#undef MEMO_QLAUNCH
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

int memo_query_sweep(int32_t membership, const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                     int64_t n_rows, int64_t q_start, int64_t q_end, const int32_t* ks, int32_t n_k,
                     int32_t n_docs, void* out, int32_t* status, void* workspace, size_t workspace_bytes,
                     void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(q_end >= q_start && q_start >= 0, "bad window [%lld, %lld)", (long long)q_start, (long long)q_end);
    MEMO_REQUIRE(ks != nullptr && n_k >= 1 && n_k <= 16, "1 .. 16 k values per sweep");
    int32_t k_max = 1;
    for (int i = 0; i < n_k; ++i) {
        MEMO_REQUIRE(ks[i] >= 1, "k must be >= 1");
        if (ks[i] > k_max) k_max = ks[i];
    }
    MEMO_REQUIRE(n_docs >= 1 && (membership ? n_docs <= query_planes_max_membership_docs() : n_docs <= 255),
                 "n_docs = %d: the sweep runs on the bit-plane kernel only", n_docs);
    MEMO_REQUIRE(n_rows >= 0 && status != nullptr, "bad rows/status");
    MEMO_REQUIRE(q_end + k_max < (1ll << 31) - (1ll << 17), "window too far for 32-bit tile arithmetic");
    const long long W = q_end - q_start;
    MEMO_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    if (W == 0) return MEMO_OK;
    MEMO_REQUIRE(out != nullptr, "out must not be NULL");
    const int64_t stride = membership ? (int64_t)W * 4 * ((n_docs + 31) / 32) : (int64_t)((W + 15) / 16 * 16);
    MEMO_REQUIRE(((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) | reinterpret_cast<uintptr_t>(f3) |
                   reinterpret_cast<uintptr_t>(out) | (uintptr_t)stride) & 15) == 0,
                 "rows and results must be 16-byte aligned");
    return launch_query_planes(membership, f1, f2, f3, n_rows, q_start, q_end, ks, n_k, n_docs, out, stride, status,
                               workspace, workspace_bytes, stream);
}

int memo_query_membership(const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                          int64_t n_rows, int64_t q_start, int64_t q_end, int32_t k,
                          int32_t n_docs, uint32_t* out_bits, int32_t* status, void* workspace,
                          size_t workspace_bytes, void* stream_) {
    using namespace memo;
    (void)workspace; (void)workspace_bytes;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(q_end >= q_start && q_start >= 0, "bad window [%lld, %lld)", (long long)q_start, (long long)q_end);
    MEMO_REQUIRE(k >= 1, "k must be >= 1");
    MEMO_REQUIRE(n_docs >= 1, "n_docs must be >= 1");
    MEMO_REQUIRE(n_rows >= 0 && status != nullptr, "bad rows/status");
    const long long W = q_end - q_start;
    const int NW = (n_docs + 31) / 32;
    MEMO_REQUIRE(NW <= QM_WORDS / 32, "n_docs = %d too large for the membership tile", n_docs);
    MEMO_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    if (W == 0) return MEMO_OK;
    MEMO_REQUIRE(out_bits != nullptr, "out_bits must not be NULL");
    // bit-plane tiles, one stream per warp (query_planes.cu)
    if (n_docs <= query_planes_max_membership_docs() && q_end + k < (1ll << 31) - (1ll << 17) &&
        workspace != nullptr && workspace_bytes >= query_planes_workspace_bytes() &&
        ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) | reinterpret_cast<uintptr_t>(f3) |
          reinterpret_cast<uintptr_t>(out_bits)) & 15) == 0 &&
        !(getenv("MEMO_QUERY_PLANES") && atoi(getenv("MEMO_QUERY_PLANES")) == 0))
        return launch_query_planes(1, f1, f2, f3, n_rows, q_start, q_end, &k, 1, n_docs, out_bits, 0, status, workspace,
                                   workspace_bytes, stream);
    const int TP = QM_WORDS / NW;
    const long long n_tiles = (W + TP - 1) / TP;
    const int sms = device_sm_count();
    long long grid = (long long)sms * 5;
    if (grid > n_tiles) grid = n_tiles;
    query_membership_kernel<<<(unsigned)grid, QTHREADS, 0, stream>>>(
        f1, f2, f3, n_rows, q_start, W, k, n_docs, NW, TP, out_bits, status);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // extern "C"
