// Error reporting and device queries shared by the libmemo_b200.so entry points.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace memo {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        return 148;
    return sms;
}

}  // namespace memo

extern "C" {

int memo_abi_version(void) { return MEMO_B200_ABI_VERSION; }

const char* memo_last_error(void) { return memo::g_error; }

int64_t memo_launch_count(int32_t reset) {
    return reset ? memo::g_launches.exchange(0) : memo::g_launches.load();
}

int memo_device_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        memo::set_error("no CUDA device");
        return MEMO_ERR_CUDA;
    }
    return memo::device_sm_count();
}

}  // extern "C"
