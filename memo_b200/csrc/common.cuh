// Shared helpers for libmemo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/memo_b200.h"

namespace memo {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

void set_error(const char* fmt, ...);

#define MEMO_CUDA_TRY(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::memo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                 \
            return MEMO_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)

#define MEMO_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::memo::set_error(__VA_ARGS__);  \
            return MEMO_ERR_ARG;             \
        }                                    \
    } while (0)

// every kernel launch of the library is counted (memo_launch_count: bench.py's gpu_launches)
void note_launches(int n);

#define MEMO_LAUNCH_CHECK(n)                 \
    do {                                     \
        ::memo::note_launches(n);            \
        MEMO_CUDA_TRY(cudaGetLastError());   \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int device_sm_count();

// scratch the general (three-pass) index build needs (index_general.cu)
size_t general_workspace_bytes(int64_t rows, int32_t n_cols, const memo_segment_t* segs,
                               int32_t n_seg, const memo_index_opts_t* opts);

// query_planes.cu: conservation query with uint8 results (n_docs <= 255) and membership
// query (n_docs <= query_planes_max_membership_docs()); windows ending below 2^31 - 2^17,
// f1 / f2 / f3 16-byte aligned
size_t query_planes_workspace_bytes();
int query_planes_max_membership_docs();
int launch_query_planes(int membership, const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                        int64_t n_rows, int64_t q_start, int64_t q_end, const int32_t* ks, int32_t n_k,
                        int32_t n_docs, void* out, int64_t out_stride, int32_t* status, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);

}  // namespace memo
