// Warp-level descending sort of keys held in registers (shared by the index kernels).
#pragma once
#include "common.cuh"

namespace memo {
namespace {

template <int KPL>
__device__ __forceinline__ void local_sort_desc(uint32_t (&a)[KPL]) {
#pragma unroll
    for (int round = 0; round < KPL; ++round) {
#pragma unroll
        for (int i = round & 1; i + 1 < KPL; i += 2) {
            uint32_t hi = max(a[i], a[i + 1]);
            uint32_t lo = min(a[i], a[i + 1]);
            a[i] = hi;
            a[i + 1] = lo;
        }
    }
}

// Sort G*KPL keys held by a group of G lanes (KPL per lane, any order) so that
// key i = lg*KPL + k is the i-th largest.  Bitonic network over lanes with
// merge-split exchanges (each lane keeps a descending run).
template <int G, int KPL>
__device__ __forceinline__ void group_sort_desc(uint32_t (&a)[KPL], int lg) {
    local_sort_desc<KPL>(a);
#pragma unroll
    for (int size = 2; size <= G; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const bool desc = (lg & size) == 0;
            const bool lower = (lg & stride) == 0;
            const bool keep_large = (lower == desc);
            uint32_t b[KPL];
#pragma unroll
            for (int k = 0; k < KPL; ++k) b[k] = __shfl_xor_sync(FULL, a[KPL - 1 - k], stride);
#pragma unroll
            for (int k = 0; k < KPL; ++k) a[k] = keep_large ? max(a[k], b[k]) : min(a[k], b[k]);
            local_sort_desc<KPL>(a);
        }
    }
}

}  // namespace
}  // namespace memo
