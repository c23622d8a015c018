// Synthetic HPRC-shaped document-array profile (measurement only).
// Integer-only, bit-identical to oracle/memo_oracle.py:synth_dap:
//   h  = splitmix64(seed ^ p*K1 ^ c*K2)
//   d  = short geometric draw, or (prob 2^-9) a long draw with geometric decay
//   MS[p][c] = min( max_{q<=p}(d[q][c] + q) - p, rec_len - p )
// d < 2^15, so the prefix maximum only needs the previous LOOKBACK_CHUNKS
// chunks of SY_CHUNK rows.
#include "common.cuh"

namespace memo {
namespace {

constexpr int SY_CHUNK = 4096;
constexpr int SY_LOOKBACK_CHUNKS = 5;     // 5 * 4096 = 20480 >= 18432 > max draw
constexpr int SY_SUB_MAX = 64;            // rows per shared-memory sub-tile (at most)
constexpr int SY_THREADS = 256;

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ long long draw(unsigned long long seed, long long p, int c, int dense) {
    const unsigned long long h = splitmix64(seed ^ ((unsigned long long)p * 0x9E3779B97F4A7C15ull) ^
                                            ((unsigned long long)c * 0xC2B2AE3D27D4EB4Full));
    const long long d_s = 12 + (__ffsll((long long)(h | (1ull << 20))) - 1);
    long long d_l;
    bool is_long;
    if (dense) {
        is_long = ((h >> 20) & 1023ull) == 0;
        d_l = ((long long)(__ffsll((long long)((h >> 32) | (1ull << 16)))) << 8) + (long long)((h >> 48) & 255ull);
    } else {
        is_long = ((h >> 20) & 511ull) == 0;
        d_l = ((long long)(__ffsll((long long)((h >> 32) | (1ull << 16)))) << 10) + (long long)((h >> 48) & 1023ull);
    }
    return is_long ? max(d_s, d_l) : d_s;
}

// chunkmax[chunk][c] = max over the chunk's rows of d + q.  Chunks are indexed
// from (row0 / SY_CHUNK) - SY_LOOKBACK_CHUNKS; rows < 0 contribute nothing.
__global__ void synth_chunkmax_kernel(long long chunk_first, long long n_chunks, int C,
                                      long long rec_len, unsigned long long seed, int dense,
                                      long long* __restrict__ chunkmax) {
    const long long chunk = chunk_first + blockIdx.x;
    const int c = blockIdx.y;
    long long m = -1;
    if (chunk >= 0) {
        const long long q0 = chunk * SY_CHUNK;
        for (int i = threadIdx.x; i < SY_CHUNK; i += blockDim.x) {
            const long long q = q0 + i;
            if (q < rec_len) m = max(m, draw(seed, q, c, dense) + q);
        }
    }
    __shared__ long long red[SY_THREADS];
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) chunkmax[(long long)blockIdx.x * C + c] = red[0];
}

__global__ void __launch_bounds__(SY_THREADS)
synth_fill_kernel(int32_t* __restrict__ dap, long long row0, long long rows, int C, int ld,
                  long long rec_len, unsigned long long seed, int dense, long long chunk_first,
                  const long long* __restrict__ chunkmax, int SY_SUB) {
    extern __shared__ long long sm[];          // [SY_SUB][C] reach values, then [C] running
    long long* tile = sm;
    long long* running = sm + (size_t)SY_SUB * C;
    const long long chunk = row0 / SY_CHUNK + blockIdx.x;     // absolute chunk index
    const long long ci = chunk - chunk_first;                 // index into chunkmax
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        long long m = -1;
        for (int b = 1; b <= SY_LOOKBACK_CHUNKS; ++b) m = max(m, chunkmax[(ci - b) * C + c]);
        running[c] = m;
    }
    __syncthreads();
    const long long q_begin = chunk * SY_CHUNK;
    const long long q_end = min(q_begin + SY_CHUNK, min(rec_len, row0 + rows));
    for (long long q0 = q_begin; q0 < q_end; q0 += SY_SUB) {
        const int nr = (int)min((long long)SY_SUB, q_end - q0);
        for (int i = threadIdx.x; i < nr * C; i += blockDim.x) {
            const int r = i / C, c = i % C;
            tile[i] = draw(seed, q0 + r, c, dense) + (q0 + r);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            long long m = running[c];
            for (int r = 0; r < nr; ++r) {
                m = max(m, tile[r * C + c]);
                tile[r * C + c] = m;
            }
            running[c] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nr * C; i += blockDim.x) {
            const int r = i / C, c = i % C;
            const long long p = q0 + r;
            if (p >= row0) {
                const long long ms = min(tile[i] - p, rec_len - p);
                dap[(p - row0) * (long long)ld + c] = (int32_t)ms;
            }
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace memo

extern "C" {

size_t memo_synth_workspace_bytes(int64_t rows, int32_t n_cols) {
    if (rows < 0 || n_cols < 1) return 0;
    const int64_t n_chunks = rows / memo::SY_CHUNK + 2 + memo::SY_LOOKBACK_CHUNKS;
    return memo::align_up(sizeof(long long) * (size_t)n_chunks * (size_t)n_cols, 256);
}

int memo_synth_dap(int32_t* dap, int64_t row0, int64_t rows, int32_t n_cols, int32_t ld,
                   int64_t rec_len, uint64_t seed, int32_t dense, void* workspace,
                   size_t workspace_bytes, void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(rows >= 0 && row0 >= 0 && row0 + rows <= rec_len, "rows [%lld, %lld) outside record of %lld",
                 (long long)row0, (long long)(row0 + rows), (long long)rec_len);
    MEMO_REQUIRE(n_cols >= 1 && ld >= n_cols && n_cols <= 2048, "bad n_cols/ld");
    if (rows == 0) return MEMO_OK;
    MEMO_REQUIRE(dap != nullptr, "dap must not be NULL");
    const long long c_lo = row0 / SY_CHUNK;
    const long long c_hi = (row0 + rows - 1) / SY_CHUNK;
    const long long chunk_first = c_lo - SY_LOOKBACK_CHUNKS;
    const long long n_chunks = c_hi - chunk_first + 1;
    const size_t need = sizeof(long long) * (size_t)n_chunks * (size_t)n_cols;
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return MEMO_ERR_WORKSPACE;
    }
    long long* chunkmax = static_cast<long long*>(workspace);
    synth_chunkmax_kernel<<<dim3((unsigned)n_chunks, (unsigned)n_cols), SY_THREADS, 0, stream>>>(
        chunk_first, n_chunks, n_cols, rec_len, seed, dense, chunkmax);
    MEMO_LAUNCH_CHECK(1);
    int sub = (int)((200 * 1024 / sizeof(long long) - n_cols) / n_cols);
    if (sub > SY_SUB_MAX) sub = SY_SUB_MAX;
    MEMO_REQUIRE(sub >= 1, "n_cols too large for the generator");
    const size_t smem = sizeof(long long) * ((size_t)sub * n_cols + n_cols);
    MEMO_CUDA_TRY(cudaFuncSetAttribute(synth_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    synth_fill_kernel<<<(unsigned)(c_hi - c_lo + 1), SY_THREADS, smem, stream>>>(
        dap, row0, rows, n_cols, ld, rec_len, seed, dense, chunk_first, chunkmax, sub);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // extern "C"
