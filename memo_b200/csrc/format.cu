// Text formatters on device: query results and the BED rows of the index.
//
// Replace the text emitters of the reference's src/memo_query.py print_res
// :65-71: conservation = one decimal integer per line (print(*rec, sep='\n')),
// membership = n_docs space-separated 0/1 digits per line (np.savetxt '%i').
// The bytes produced are exactly the reference's file contents.
#include "common.cuh"

namespace memo {
namespace {

constexpr int FT = 256;          // threads per block
constexpr int FV = 16;           // values per thread
constexpr int FB = FT * FV;      // values per block

__device__ __forceinline__ int dec_len(uint32_t v) {
    return v < 10 ? 2 : v < 100 ? 3 : v < 1000 ? 4 : v < 10000 ? 5 : 6;   // digits + '\n'
}

template <typename T>
__global__ void __launch_bounds__(FT) fmt_len_kernel(const T* __restrict__ vals, long long n,
                                                     unsigned long long* __restrict__ blocksum) {
    const long long base = (long long)blockIdx.x * FB + (long long)threadIdx.x * FV;
    unsigned len = 0;
    for (int i = 0; i < FV; ++i)
        if (base + i < n) len += dec_len(vals[base + i]);
    __shared__ unsigned red[FT / 32];
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(FULL, len, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < FT / 32; ++i) t += red[i];
        blocksum[blockIdx.x] = t;
    }
}

// exclusive scan of blocksum[nb] in place (single block), total -> *out_len
__global__ void fmt_scan_kernel(unsigned long long* blocksum, long long nb, long long* out_len) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long b0 = 0; b0 < nb; b0 += blockDim.x) {
        const long long i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? blocksum[i] : 0;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
        const unsigned long long c = carry;
        if (i < nb) blocksum[i] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_len = (long long)carry;
}

template <typename T>
__global__ void __launch_bounds__(FT) fmt_write_kernel(const T* __restrict__ vals, long long n,
                                                       const unsigned long long* __restrict__ blockoff,
                                                       char* __restrict__ out) {
    const long long base = (long long)blockIdx.x * FB + (long long)threadIdx.x * FV;
    uint32_t v[FV];
    unsigned len = 0;
    for (int i = 0; i < FV; ++i) {
        v[i] = (base + i < n) ? (uint32_t)vals[base + i] : 0;
        if (base + i < n) len += dec_len(v[i]);
    }
    // exclusive prefix of len within the block
    unsigned incl = len;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ unsigned wsum[FT / 32];
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    char* dst = out + blockoff[blockIdx.x] + woff + incl - len;
    for (int i = 0; i < FV; ++i) {
        if (base + i >= n) break;
        const int l = dec_len(v[i]);
        uint32_t x = v[i];
        dst[l - 1] = '\n';
        for (int d = l - 2; d >= 0; --d) {
            dst[d] = (char)('0' + x % 10);
            x /= 10;
        }
        dst += l;
    }
}

__global__ void fmt_membership_kernel(const uint32_t* __restrict__ bits, long long W, int n_docs,
                                      int NW, char* __restrict__ out) {
    const long long total = W * (long long)n_docs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / n_docs;
        const int j = (int)(i % n_docs);
        const uint32_t w = bits[p * NW + (j >> 5)];
        char2 c;
        c.x = (char)('0' + ((w >> (j & 31)) & 1u));
        c.y = (j == n_docs - 1) ? '\n' : ' ';
        reinterpret_cast<char2*>(out)[i] = c;
    }
}

// ---------------------------------------------------------------- BED rows
// One index row as the reference prints it (src/dap_to_bed.py:105,
// print('\t'.join(map(str, [header, start, end, annot])))): name TAB start TAB end TAB order LF.
// All rows of a call belong to one record (one name).
__device__ __forceinline__ int dec_digits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
         : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
__device__ __forceinline__ char* put_dec(char* dst, uint32_t v, char term) {
    const int l = dec_digits(v);
    for (int d = l - 1; d >= 0; --d) {
        dst[d] = (char)('0' + v % 10u);
        v /= 10u;
    }
    dst[l] = term;
    return dst + l + 1;
}

constexpr int BV = 8;            // rows per thread
constexpr int BB = FT * BV;      // rows per block

__global__ void __launch_bounds__(FT) bed_len_kernel(const int32_t* __restrict__ f1, const uint32_t* __restrict__ f2,
                                                     const int32_t* __restrict__ f3, long long n, int name_len,
                                                     unsigned long long* __restrict__ blocksum) {
    const long long base = (long long)blockIdx.x * BB + (long long)threadIdx.x * BV;
    unsigned len = 0;
    for (int i = 0; i < BV; ++i)
        if (base + i < n)
            len += (unsigned)(name_len + 4 + dec_digits((uint32_t)f1[base + i]) + dec_digits(f2[base + i]) +
                              dec_digits((uint32_t)f3[base + i]));
    __shared__ unsigned red[FT / 32];
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(FULL, len, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < FT / 32; ++i) t += red[i];
        blocksum[blockIdx.x] = t;
    }
}

struct BedName { char c[256]; };

__global__ void __launch_bounds__(FT) bed_write_kernel(const int32_t* __restrict__ f1, const uint32_t* __restrict__ f2,
                                                       const int32_t* __restrict__ f3, long long n, int name_len,
                                                       const BedName name,
                                                       const unsigned long long* __restrict__ blockoff,
                                                       char* __restrict__ out) {
    const long long base = (long long)blockIdx.x * BB + (long long)threadIdx.x * BV;
    uint32_t a[BV], b[BV], c[BV];
    unsigned len = 0;
    for (int i = 0; i < BV; ++i) {
        const bool live = base + i < n;
        a[i] = live ? (uint32_t)f1[base + i] : 0u;
        b[i] = live ? f2[base + i] : 0u;
        c[i] = live ? (uint32_t)f3[base + i] : 0u;
        if (live) len += (unsigned)(name_len + 4 + dec_digits(a[i]) + dec_digits(b[i]) + dec_digits(c[i]));
    }
    unsigned incl = len;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ unsigned wsum[FT / 32];
    __shared__ char s_name[256];
    if (threadIdx.x < name_len) s_name[threadIdx.x] = name.c[threadIdx.x];
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    char* dst = out + blockoff[blockIdx.x] + woff + incl - len;
    for (int i = 0; i < BV; ++i) {
        if (base + i >= n) break;
        for (int j = 0; j < name_len; ++j) dst[j] = s_name[j];
        dst[name_len] = '\t';
        dst = put_dec(dst + name_len + 1, a[i], '\t');
        dst = put_dec(dst, b[i], '\t');
        dst = put_dec(dst, c[i], '\n');
    }
}

}  // namespace
}  // namespace memo

extern "C" {

size_t memo_format_bed_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return memo::align_up(sizeof(unsigned long long) * (size_t)((n + memo::BB - 1) / memo::BB + 1), 256);
}

size_t memo_format_bed_max_bytes(int64_t n, int32_t name_len) {
    if (n < 0) n = 0;
    return (size_t)n * (size_t)(name_len + 4 + 10 + 10 + 10);
}

int memo_format_bed(const int32_t* f1, const uint32_t* f2, const int32_t* f3, int64_t n, const char* name,
                    int32_t name_len, char* out_text, int64_t* out_len, void* workspace, size_t workspace_bytes,
                    void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(n >= 0 && out_len != nullptr, "bad n/out_len");
    MEMO_REQUIRE(name_len >= 0 && name_len <= 256 && (name_len == 0 || name != nullptr),
                 "record name of %d bytes (at most 256 on the device formatter)", name_len);
    if (n == 0) {
        MEMO_CUDA_TRY(cudaMemsetAsync(out_len, 0, sizeof(int64_t), stream));
        return MEMO_OK;
    }
    MEMO_REQUIRE(f1 && f2 && f3 && out_text, "NULL buffer");
    const long long nb = (n + BB - 1) / BB;
    if (workspace == nullptr || workspace_bytes < sizeof(unsigned long long) * (size_t)nb) {
        set_error("workspace too small");
        return MEMO_ERR_WORKSPACE;
    }
    BedName nm;
    for (int i = 0; i < 256; ++i) nm.c[i] = i < name_len ? name[i] : 0;
    unsigned long long* bs = static_cast<unsigned long long*>(workspace);
    bed_len_kernel<<<(unsigned)nb, FT, 0, stream>>>(f1, f2, f3, n, name_len, bs);
    fmt_scan_kernel<<<1, 1024, 0, stream>>>(bs, nb, reinterpret_cast<long long*>(out_len));
    bed_write_kernel<<<(unsigned)nb, FT, 0, stream>>>(f1, f2, f3, n, name_len, nm, bs, out_text);
    MEMO_LAUNCH_CHECK(3);
    return MEMO_OK;
}

size_t memo_format_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return memo::align_up(sizeof(unsigned long long) * (size_t)((n + memo::FB - 1) / memo::FB + 1), 256);
}

int memo_format_conservation(const void* vals, int32_t is_u16, int64_t n, char* out_text,
                             int64_t* out_len, void* workspace, size_t workspace_bytes,
                             void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(n >= 0 && out_len != nullptr, "bad n/out_len");
    if (n == 0) {
        MEMO_CUDA_TRY(cudaMemsetAsync(out_len, 0, sizeof(int64_t), stream));
        return MEMO_OK;
    }
    MEMO_REQUIRE(vals && out_text, "NULL buffer");
    const long long nb = (n + FB - 1) / FB;
    if (workspace == nullptr || workspace_bytes < sizeof(unsigned long long) * (size_t)nb) {
        set_error("workspace too small");
        return MEMO_ERR_WORKSPACE;
    }
    unsigned long long* bs = static_cast<unsigned long long*>(workspace);
    if (is_u16) fmt_len_kernel<uint16_t><<<(unsigned)nb, FT, 0, stream>>>(static_cast<const uint16_t*>(vals), n, bs);
    else fmt_len_kernel<uint8_t><<<(unsigned)nb, FT, 0, stream>>>(static_cast<const uint8_t*>(vals), n, bs);
    fmt_scan_kernel<<<1, 1024, 0, stream>>>(bs, nb, reinterpret_cast<long long*>(out_len));
    if (is_u16) fmt_write_kernel<uint16_t><<<(unsigned)nb, FT, 0, stream>>>(static_cast<const uint16_t*>(vals), n, bs, out_text);
    else fmt_write_kernel<uint8_t><<<(unsigned)nb, FT, 0, stream>>>(static_cast<const uint8_t*>(vals), n, bs, out_text);
    MEMO_LAUNCH_CHECK(3);
    return MEMO_OK;
}

int memo_format_membership(const uint32_t* bits, int64_t W, int32_t n_docs, char* out_text,
                           void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(W >= 0 && n_docs >= 1, "bad W/n_docs");
    if (W == 0) return MEMO_OK;
    MEMO_REQUIRE(bits && out_text, "NULL buffer");
    MEMO_REQUIRE((reinterpret_cast<uintptr_t>(out_text) & 1) == 0, "out_text must be 2-byte aligned");
    const int NW = (n_docs + 31) / 32;
    const long long total = W * (long long)n_docs;
    long long grid = (total + 255) / 256;
    const long long cap = (long long)device_sm_count() * 32;
    if (grid > cap) grid = cap;
    fmt_membership_kernel<<<(unsigned)grid, 256, 0, stream>>>(bits, W, n_docs, NW, out_text);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

}  // extern "C"
