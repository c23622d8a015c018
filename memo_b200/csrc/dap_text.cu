// dap.txt on the device: text -> int32 DAP rows (sm_100a).
//
// Replaces the row parse of the reference's src/dap_to_bed.py: read_file :14-18 (line by
// line) and get_new_record :85-88 (`map(int, row.split(' '))`, first field = position).  The
// host text parse is what bounds the drop-in end to end (~0.26 GB/s with pyarrow,
// ~0.03 GB/s in the reference); here the raw bytes of whole lines go to the device and are
// parsed there: HBM/PCIe bound, ~1 B read per byte, 4 n_cols B written per line.
//
// The text is cut into tiles of 1 KB, one per warp (a lane owns 32 consecutive bytes):
//   text_count_kernel   per tile: newlines, and the spaces after its last newline
//   text_scan_kernel    one CTA: row index and field index at every tile's first byte
//                       (a segmented scan: a newline resets the field count)
//   text_parse_kernel   per tile: the same scan across the lanes, then every lane walks its
//                       bytes; a token belongs to the lane it starts in (digits are read on
//                       past the lane's bytes), value -> out[row][field - 1]; field 0 is the
//                       position and must equal pos_first + row (index.sh:83 `nl -v0`).
// Everything int() / the row layout would reject is flagged (result[1]): a character that
// is no digit, space or newline; an empty field; a line with another number of fields;
// positions that are not consecutive; values >= 2^31.
#include "common.cuh"

namespace memo {
namespace {

constexpr int TX_TILE = 1024;          // bytes per warp
constexpr int TX_LANE = 32;            // bytes per lane
constexpr int TX_WARPS = 8;            // warps per CTA

struct Agg {                           // newlines | spaces after the last newline (all spaces if none)
    int nl, sp;
    bool has;
};
__device__ __forceinline__ Agg combine(const Agg& a, const Agg& b) {     // a before b
    Agg r;
    r.nl = a.nl + b.nl;
    r.has = a.has || b.has;
    r.sp = b.has ? b.sp : a.sp + b.sp;
    return r;
}
__device__ __forceinline__ Agg shfl_up_agg(const Agg& a, int d) {
    Agg r;
    r.nl = __shfl_up_sync(FULL, a.nl, d);
    r.sp = __shfl_up_sync(FULL, a.sp, d);
    r.has = __shfl_up_sync(FULL, (int)a.has, d) != 0;
    return r;
}
// aggregate of the lane's 32 bytes at text[g0 ..) (bytes past n count as nothing)
__device__ __forceinline__ Agg lane_agg(const uint8_t* __restrict__ text, long long g0, long long n) {
    Agg a{0, 0, false};
    if (g0 >= n) return a;
    const uint4 w0 = *reinterpret_cast<const uint4*>(text + g0);
    const uint4 w1 = *reinterpret_cast<const uint4*>(text + g0 + 16);
    const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const int lim = (int)min((long long)TX_LANE, n - g0);
#pragma unroll
    for (int i = 0; i < TX_LANE; ++i) {
        const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        if (i < lim) {
            if (c == '\n') { ++a.nl; a.has = true; a.sp = 0; }
            else if (c == ' ') ++a.sp;
        }
    }
    return a;
}
// inclusive scan of the lanes' aggregates
__device__ __forceinline__ Agg warp_scan(Agg a, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Agg up = shfl_up_agg(a, d);
        if (lane >= d) a = combine(up, a);
    }
    return a;
}

__global__ void __launch_bounds__(TX_WARPS * 32)
text_count_kernel(const uint8_t* __restrict__ text, long long n, long long n_tiles, int* __restrict__ t_nl,
                  int* __restrict__ t_sp, unsigned char* __restrict__ t_has) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * TX_WARPS + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const Agg a = warp_scan(lane_agg(text, tile * TX_TILE + lane * TX_LANE, n), lane);
    if (lane == 31) {
        t_nl[tile] = a.nl;
        t_sp[tile] = a.sp;
        t_has[tile] = a.has ? 1 : 0;
    }
}

// row / field index at the first byte of every tile; result[0] = number of lines
__global__ void __launch_bounds__(1024)
text_scan_kernel(long long n_tiles, const int* __restrict__ t_nl, const int* __restrict__ t_sp,
                 const unsigned char* __restrict__ t_has, long long* __restrict__ t_row, int* __restrict__ t_field,
                 int64_t* result) {
    __shared__ long long s_nl[1024];
    __shared__ int s_sp[1024];
    __shared__ unsigned char s_has[1024];
    const long long per = (n_tiles + 1023) / 1024;
    const long long lo = min(n_tiles, (long long)threadIdx.x * per), hi = min(n_tiles, lo + per);
    long long nl = 0;
    int sp = 0;
    bool has = false;
    for (long long t = lo; t < hi; ++t) {
        nl += t_nl[t];
        sp = t_has[t] ? t_sp[t] : sp + t_sp[t];
        has = has || t_has[t];
    }
    s_nl[threadIdx.x] = nl; s_sp[threadIdx.x] = sp; s_has[threadIdx.x] = has ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) {                      // exclusive scan over the 1024 spans
        long long rn = 0;
        int rs = 0;
        for (int i = 0; i < 1024; ++i) {
            const long long a = s_nl[i];
            const int b = s_sp[i];
            const bool h = s_has[i] != 0;
            s_nl[i] = rn; s_sp[i] = rs;
            rn += a;
            rs = h ? b : rs + b;
        }
        result[0] = rn;
    }
    __syncthreads();
    nl = s_nl[threadIdx.x];
    sp = s_sp[threadIdx.x];
    for (long long t = lo; t < hi; ++t) {
        t_row[t] = nl;
        t_field[t] = sp;
        nl += t_nl[t];
        sp = t_has[t] ? t_sp[t] : sp + t_sp[t];
    }
}

constexpr int ERR_CHAR = 1, ERR_FIELDS = 2, ERR_POS = 4, ERR_RANGE = 8, ERR_ROWS = 16;

__global__ void __launch_bounds__(TX_WARPS * 32)
text_parse_kernel(const uint8_t* __restrict__ text, long long n, long long n_tiles, const long long* __restrict__ t_row,
                  const int* __restrict__ t_field, int n_cols, long long pos_first, int32_t* __restrict__ out,
                  long long max_rows, int ld, int64_t* result) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * TX_WARPS + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const long long g0 = tile * TX_TILE + lane * TX_LANE;
    const Agg mine = lane_agg(text, g0, n);
    Agg incl = warp_scan(mine, lane);
    Agg before = shfl_up_agg(incl, 1);                 // aggregate of the lanes before this one
    if (lane == 0) before = Agg{0, 0, false};
    long long row = t_row[tile] + before.nl;
    int field = before.has ? before.sp : t_field[tile] + before.sp;
    if (g0 >= n) return;
    int err = 0;
    const long long end = min(g0 + TX_LANE, n);
    uint32_t prev = g0 > 0 ? text[g0 - 1] : (uint32_t)'\n';
    for (long long g = g0; g < end; ++g) {
        const uint32_t c = text[g];
        if (c == '\n' || c == ' ') {
            if (prev == '\n' || prev == ' ') err |= ERR_CHAR;          // an empty field: int('')
            if (c == '\n') {
                if (field != n_cols) err |= ERR_FIELDS;
                ++row;
                field = 0;
            } else {
                ++field;
            }
        } else if (c - '0' <= 9u) {
            if (prev == '\n' || prev == ' ') {                         // a token starts here
                unsigned long long v = c - '0';
                int digits = 1;
                for (long long j = g + 1; j < n; ++j) {
                    const uint32_t d = text[j] - '0';
                    if (d > 9u) break;
                    if (++digits > 10) break;
                    v = v * 10ull + d;
                }
                if (digits > 10 || v > 2147483647ull) {
                    err |= ERR_RANGE;
                } else if (field == 0) {
                    if ((long long)v != pos_first + row) err |= ERR_POS;
                } else if (field <= n_cols) {
                    if (row < max_rows) out[row * (long long)ld + field - 1] = (int32_t)v;
                    else err |= ERR_ROWS;
                }                                                       // (field > n_cols: flagged at the newline)
            }
        } else {
            err |= ERR_CHAR;
        }
        prev = c;
    }
    if (err) atomicOr(reinterpret_cast<unsigned long long*>(&result[1]), (unsigned long long)err);
}

}  // namespace
}  // namespace memo

extern "C" {

size_t memo_dap_text_workspace_bytes(int64_t n_bytes) {
    if (n_bytes < 0) n_bytes = 0;
    const size_t nt = (size_t)((n_bytes + memo::TX_TILE - 1) / memo::TX_TILE) + 1;
    return memo::align_up(4 * nt, 256) * 3 + memo::align_up(8 * nt, 256) + memo::align_up(nt, 256);
}

int memo_dap_text_parse(const uint8_t* text, int64_t n_bytes, int32_t n_cols, int64_t pos_first, int32_t* out,
                        int64_t max_rows, int32_t ld, int64_t* result, void* workspace, size_t workspace_bytes,
                        void* stream_) {
    using namespace memo;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MEMO_REQUIRE(n_bytes >= 0 && n_cols >= 1 && ld >= n_cols && max_rows >= 0, "bad dap text shape");
    MEMO_REQUIRE(result != nullptr && (n_bytes == 0 || (text != nullptr && out != nullptr)), "NULL argument");
    MEMO_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15) == 0, "text must be 16-byte aligned");
    MEMO_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(int64_t) * 4, stream));
    if (n_bytes == 0) return MEMO_OK;
    const long long nt = (n_bytes + TX_TILE - 1) / TX_TILE;
    const size_t need = memo_dap_text_workspace_bytes(n_bytes);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return MEMO_ERR_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    const size_t a4 = align_up(4 * (size_t)(nt + 1), 256);
    int* t_nl = reinterpret_cast<int*>(ws);
    int* t_sp = reinterpret_cast<int*>(ws + a4);
    int* t_field = reinterpret_cast<int*>(ws + 2 * a4);
    long long* t_row = reinterpret_cast<long long*>(ws + 3 * a4);
    unsigned char* t_has = reinterpret_cast<unsigned char*>(ws + 3 * a4 + align_up(8 * (size_t)(nt + 1), 256));
    const unsigned grid = (unsigned)((nt + TX_WARPS - 1) / TX_WARPS);
    text_count_kernel<<<grid, TX_WARPS * 32, 0, stream>>>(text, n_bytes, nt, t_nl, t_sp, t_has);
    text_scan_kernel<<<1, 1024, 0, stream>>>(nt, t_nl, t_sp, t_has, t_row, t_field, result);
    text_parse_kernel<<<grid, TX_WARPS * 32, 0, stream>>>(text, n_bytes, nt, t_row, t_field, n_cols, pos_first, out,
                                                          max_rows, ld, result);
    MEMO_LAUNCH_CHECK(3);
    return MEMO_OK;
}

}  // extern "C"
