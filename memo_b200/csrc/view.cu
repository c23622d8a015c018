// `memo view` binning on sm_100a: per-bin histogram of a conservation vector.
//
// Replaces the reference's src/plot_conservation.py preprocess_data :46-58: the window is
// cut into n_bins position bins (edges = int(linspace(0, positions, n_bins + 1)), computed
// on the host exactly as :52) and every bin counts how many of its positions hold each
// conservation value 0 .. n_docs (the Counter of :55; the host divides by the bin size,
// :56).  HBM bound: 1 B (or 2) per position read once, n_bins * (n_docs + 1) counters out.
//
// grid = (slices, bins): a CTA takes one slice of SLICE positions of one bin, counts it in
// a shared-memory histogram (lanes holding the same value add once: match_any, the data
// are a few long runs of equal values) and adds that to the bin's global counters.
#include "common.cuh"

namespace memo {
namespace {

constexpr int VW_THREADS = 256;
constexpr int VW_SLICE = 1 << 16;

template <typename T>
__global__ void __launch_bounds__(VW_THREADS)
view_bins_kernel(const T* __restrict__ vals, const long long* __restrict__ edges, int n_values,
                 unsigned long long* __restrict__ counts, int32_t* __restrict__ status) {
    extern __shared__ unsigned int hist[];
    const int bin = blockIdx.y;
    const long long b0 = edges[bin], b1 = edges[bin + 1];
    const long long lo = b0 + (long long)blockIdx.x * VW_SLICE;
    if (lo >= b1) return;
    const long long hi = min(lo + (long long)VW_SLICE, b1);
    for (int i = threadIdx.x; i < n_values; i += VW_THREADS) hist[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    bool bad = false;
    for (long long p0 = lo + (threadIdx.x & ~31); p0 < hi; p0 += VW_THREADS) {
        const long long p = p0 + lane;
        const bool in = p < hi;
        const unsigned v = in ? (unsigned)vals[p] : 0xFFFFFFFFu;
        const unsigned active = __ballot_sync(FULL, in);
        if (in) {
            const unsigned same = __match_any_sync(active, v);
            if (v >= (unsigned)n_values) bad = true;                 // (a value above n_docs: the reference
            else if (lane == __ffs(same) - 1) atomicAdd(&hist[v], (unsigned)__popc(same));   //  raises KeyError-free 0s; flagged)
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_values; i += VW_THREADS)
        if (hist[i]) atomicAdd(&counts[(long long)bin * n_values + i], (unsigned long long)hist[i]);
    if (bad) *status = 1;
}

}  // namespace
}  // namespace memo

extern "C" {

int memo_view_bins(const void* vals, int32_t is_u16, int64_t n, int32_t n_docs, int32_t n_bins,
                   const int64_t* edges, uint64_t* counts, int32_t* status, void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(n >= 0 && n_docs >= 0 && n_bins >= 1, "bad view_bins shape");
    MEMO_REQUIRE(edges != nullptr && counts != nullptr && status != nullptr, "NULL argument");
    const int n_values = n_docs + 1;
    MEMO_REQUIRE(n_values <= 12000, "n_docs too large for the shared-memory histogram");
    MEMO_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)n_bins * n_values, stream));
    MEMO_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    if (n == 0) return MEMO_OK;
    // bins are near equal (linspace): ceil(n / n_bins) + 1 positions at most
    const long long bin_len = n / n_bins + 2;
    const unsigned slices = (unsigned)((bin_len + VW_SLICE - 1) / VW_SLICE);
    const dim3 grid(slices, (unsigned)n_bins);
    const size_t smem = sizeof(unsigned int) * (size_t)n_values;
    if (is_u16)
        view_bins_kernel<uint16_t><<<grid, VW_THREADS, smem, stream>>>(
            static_cast<const uint16_t*>(vals), reinterpret_cast<const long long*>(edges), n_values,
            reinterpret_cast<unsigned long long*>(counts), status);
    else
        view_bins_kernel<uint8_t><<<grid, VW_THREADS, smem, stream>>>(
            static_cast<const uint8_t*>(vals), reinterpret_cast<const long long*>(edges), n_values,
            reinterpret_cast<unsigned long long*>(counts), status);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // extern "C"
