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

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int device_sm_count();

}  // namespace memo
