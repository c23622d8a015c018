// DAP -> MEMO index rows on sm_100a: the GENERAL build, exact for arbitrary
// non-negative integer input (three passes).  The single-pass tile kernel for
// matching statistics lives in index_build.cu; this one is only run when that
// kernel reports the input as irregular.
//
// Replaces the hot loop of the reference's src/dap_to_bed.py (--mem --overlap
// [--order]): get_new_record :85-91, dap_to_mem :116-134, print_interval /
// overlaps :93-109.  See DESIGN.md "index build" for the derivation; summary:
//
//   E[r][j] = p(r) + S[r][j]          S = row sorted descending (--order) or raw
//   flag    = E[r][j] >  E[r-1][j]    (== S[r-1][j] <= S[r][j], :123)
//   c[j]    = E at the last flagged row of column j (the dict of :107)
//   emit (p, min(c[j], E[r][j]), j+1) at flagged rows iff that min >= p
//
// Work decomposition: a *strip* is R consecutive rows of one record run, walked
// serially by a group of G lanes (G * KPL >= n_cols slots); a warp owns 32/G
// consecutive strips (one *ticket*), handed out in row order by an atomic
// counter.  The sorted row lives in registers and is updated incrementally:
// for matching statistics E only changes where a new MEM starts, so most rows
// need one compare per cell and nothing else.  Index rows are staged per strip
// in shared memory; the ticket's row count goes through a single-pass
// decoupled look-back (status word = 2-bit state + 62-bit count) to get its
// offset in the ordered output, then the staged rows are flushed coalesced.
// A strip that overflows its staging buffer only counts, and is replayed with
// direct global stores once its offset is known.
#include "common.cuh"
#include "warp_sort.cuh"

namespace memo {
namespace {

struct IndexParams {
    const int32_t* dap;
    int64_t rows;
    int32_t C;
    int32_t ld;
    const memo_segment_t* segs;        // device copy
    const int64_t* seg_ticket_start;   // device [n_seg + 1]
    int32_t n_seg;
    int64_t n_tickets;
    int32_t R;                         // rows per strip
    int32_t K;                         // staged records per strip
    int32_t mode;                      // 0 emit, 1 aggregate only
    const uint32_t* carry_in;          // [n_tickets * NG, C] or null
    uint32_t* agg_out;                 // mode 1: [n_tickets * NG, C]
    int32_t* out_start;
    uint32_t* out_end;
    int32_t* out_order;
    int64_t out_cap;
    int64_t* seg_out_end;
    unsigned long long* status;        // [n_tickets]
    unsigned long long* ticket_counter;
    int64_t* result;
};

constexpr unsigned long long ST_AGG = 1ull << 62;
constexpr unsigned long long ST_PREFIX = 2ull << 62;
constexpr unsigned long long ST_VALUE = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <int G, int KPL, bool ORDER>
__global__ void __launch_bounds__(256) index_kernel(const IndexParams P) {
    constexpr int NG = 32 / G;
    // incremental update pays off only for wide groups; narrow ones re-sort
    constexpr int INCR_MAX = (G >= 32) ? 4 : (G >= 16 ? 2 : 0);
    constexpr int U = (KPL <= 3) ? 4 : (KPL <= 4 ? 2 : 1);   // rows fetched ahead

    extern __shared__ uint32_t smem_u32[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / G;
    const int lg = lane % G;
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (g * G));
    const unsigned ltmask = (1u << lane) - 1u;
    const int K = P.K;
    uint32_t* const wbuf = smem_u32 + (size_t)warp * NG * K * 3;
    uint32_t* const ebuf = wbuf + (size_t)g * K * 3;
    const int C = P.C;

    // static slot geometry
    bool raw_valid[KPL];   // raw (striped) slot k of this lane is a real column
    bool out_valid[KPL];   // output slot k of this lane maps to j <= C
    int out_j[KPL];        // 1-based order / genome id of output slot k
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        raw_valid[k] = (k * G + lg) < C;
        const int i = ORDER ? (lg * KPL + k) : (k * G + lg);
        out_valid[k] = i < C;
        out_j[k] = i + 1;
    }

    bool irregular = false;

    for (;;) {
        long long ticket = 0;
        if (lane == 0) ticket = (long long)atomicAdd(P.ticket_counter, 1ull);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= P.n_tickets) break;

        // ticket -> record run
        int s_lo = 0, s_hi = P.n_seg - 1;
        while (s_lo < s_hi) {
            const int mid = (s_lo + s_hi + 1) >> 1;
            if (P.seg_ticket_start[mid] <= ticket) s_lo = mid; else s_hi = mid - 1;
        }
        const memo_segment_t seg = P.segs[s_lo];
        const long long lt = ticket - P.seg_ticket_start[s_lo];
        const bool last_ticket_of_seg = (ticket + 1 == P.seg_ticket_start[s_lo + 1]);

        const long long rs = (lt * NG + g) * (long long)P.R;           // strip rows [rs, re)
        const long long re = min(rs + (long long)P.R, (long long)seg.n_rows);
        const bool has = rs < seg.n_rows;
        const bool primed = (seg.flags & MEMO_SEG_PRIMED) && rs == 0;
        const long long d0 = rs - (primed ? 0 : 1);                    // init row
        const long long strip_id = ticket * NG + g;
        const bool chr_end = has && re == seg.n_rows && (seg.flags & MEMO_SEG_CHR_END);

        uint32_t gcount = 0;         // records produced by this group's strip
        long long gbase = 0;         // global offset of the strip (direct mode)
        bool direct = false;

        for (int pass = 0; pass < 2; ++pass) {
            // ---------------- init from row d0
            uint32_t prevE[KPL], A[KPL], cA[KPL];
            // a record-first row stores its MEMs (:130): every column counts as flagged
            unsigned seen = primed ? 0xFFFFFFFFu : 0u;
            {
                const int32_t* rowp = P.dap + (seg.row_begin + (has ? d0 : 0)) * (long long)P.ld;
                const uint32_t p_init = (uint32_t)seg.pos0 + (uint32_t)(has ? d0 : 0);
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    uint32_t v = 0;
                    if (has && raw_valid[k]) v = (uint32_t)__ldg(rowp + k * G + lg) + p_init;
                    prevE[k] = v;
                    A[k] = v;
                }
                if (ORDER) group_sort_desc<G, KPL>(A, lg);
#pragma unroll
                for (int k = 0; k < KPL; ++k) cA[k] = A[k];
                if (P.carry_in != nullptr && has && !primed) {
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        if (out_valid[k]) {
                            const uint32_t ci = P.carry_in[strip_id * C + out_j[k] - 1];
                            if (ci != NONE32) cA[k] = ci;
                        }
                    }
                }
            }
            gcount = 0;

            auto emit = [&](uint32_t p, const bool (&em)[KPL], const uint32_t (&endv)[KPL]) {
                unsigned bal[KPL];
                unsigned any = 0;
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    bal[k] = __ballot_sync(FULL, em[k]);
                    any |= bal[k];
                }
                if (any == 0) return;
                int total = 0;
                int rank[KPL];
                if (ORDER) {
                    int lower = 0;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        lower += __popc(bal[k] & gmask & ltmask);
                        total += __popc(bal[k] & gmask);
                    }
                    int run = 0;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        rank[k] = lower + run;
                        run += em[k] ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        rank[k] = total + __popc(bal[k] & gmask & ltmask);
                        total += __popc(bal[k] & gmask);
                    }
                }
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    if (em[k]) {
                        const uint32_t idx = gcount + (uint32_t)rank[k];
                        if (direct) {
                            const long long gi = gbase + idx;
                            if (gi < P.out_cap) {
                                P.out_start[gi] = (int32_t)p;
                                P.out_end[gi] = endv[k];
                                P.out_order[gi] = out_j[k];
                            }
                        } else if (idx < (uint32_t)K) {
                            ebuf[3 * idx + 0] = p;
                            ebuf[3 * idx + 1] = endv[k];
                            ebuf[3 * idx + 2] = (uint32_t)out_j[k];
                        }
                    }
                }
                gcount += (uint32_t)total;
            };

            // ---------------- rows d0+1 .. re-1
            const int32_t* const base = P.dap + seg.row_begin * (long long)P.ld + lg;
            for (int t0 = 0; t0 < P.R; t0 += U) {
                uint32_t buf[U][KPL];
                bool ract[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long r = d0 + 1 + t0 + u;
                    ract[u] = has && r < re && (t0 + u) < P.R;
                    const int32_t* rowp = base + r * (long long)P.ld;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        buf[u][k] = 0;
                        if (ract[u] && raw_valid[k]) buf[u][k] = (uint32_t)__ldg(rowp + k * G);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t p = (uint32_t)seg.pos0 + (uint32_t)(d0 + 1 + t0 + u);
                    uint32_t E[KPL];
                    bool ch[KPL];
                    bool chg = false;
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        E[k] = (ract[u] && raw_valid[k]) ? buf[u][k] + p : prevE[k];
                        ch[k] = E[k] != prevE[k];
                        chg |= ch[k];
                    }
                    if (!__any_sync(FULL, chg)) continue;

                    uint32_t Aold[KPL];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) Aold[k] = A[k];
                    if (ORDER) {
                        bool resort = true;
                        unsigned balc[KPL];
                        if (INCR_MAX > 0) {
                            int tot = 0;
                            bool down = false;
#pragma unroll
                            for (int k = 0; k < KPL; ++k) {
                                balc[k] = __ballot_sync(FULL, ch[k]) & gmask;
                                tot += __popc(balc[k]);
                                down |= ch[k] && (E[k] < prevE[k]);
                            }
                            resort = __any_sync(FULL, tot > INCR_MAX || down);
                        }
                        if (resort) {
#pragma unroll
                            for (int k = 0; k < KPL; ++k) A[k] = E[k];
                            group_sort_desc<G, KPL>(A, lg);
                        } else {
#pragma unroll
                            for (int k = 0; k < KPL; ++k) {
                                unsigned m = balc[k];
                                while (__any_sync(FULL, m != 0)) {
                                    const bool act = m != 0;
                                    const int src = act ? (__ffs(m) - 1) : lane;
                                    const uint32_t x = __shfl_sync(FULL, prevE[k], src);
                                    const uint32_t y = __shfl_sync(FULL, E[k], src);
                                    m &= m - 1;
                                    // remove one x, insert y (> x): r_new = #{A >= y}, r_old = #{A > x}
                                    int cnt = 0;
#pragma unroll
                                    for (int kk = 0; kk < KPL; ++kk)
                                        cnt += (A[kk] >= y ? 1 : 0) + (A[kk] > x ? 65536 : 0);
#pragma unroll
                                    for (int o = G / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
                                    const int r_new = cnt & 0xFFFF;
                                    const int r_old = cnt >> 16;
                                    const uint32_t up = __shfl_up_sync(FULL, A[KPL - 1], 1);
                                    uint32_t B[KPL];
#pragma unroll
                                    for (int kk = 0; kk < KPL; ++kk) {
                                        const int i = lg * KPL + kk;
                                        const uint32_t sh = (kk == 0) ? up : A[kk > 0 ? kk - 1 : 0];
                                        B[kk] = (!act || i < r_new || i > r_old) ? A[kk]
                                                                                 : (i == r_new ? y : sh);
                                    }
#pragma unroll
                                    for (int kk = 0; kk < KPL; ++kk) A[kk] = B[kk];
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < KPL; ++k) A[k] = E[k];
                    }
#pragma unroll
                    for (int k = 0; k < KPL; ++k) prevE[k] = E[k];

                    bool em[KPL];
                    uint32_t endv[KPL];
#pragma unroll
                    for (int k = 0; k < KPL; ++k) {
                        const bool fl = A[k] > Aold[k];
                        irregular |= A[k] < Aold[k];
                        const uint32_t e = min(cA[k], A[k]);
                        endv[k] = e;
                        em[k] = fl && e >= p;
                        if (fl) {
                            cA[k] = A[k];
                            seen |= 1u << k;
                        }
                    }
                    if (P.mode == 0) emit(p, em, endv);
                }
            }

            if (P.mode == 1) {
                // per-strip aggregate: end of the last flagged MEM per output column
                if (has) {
#pragma unroll
                    for (int k = 0; k < KPL; ++k)
                        if (out_valid[k])
                            P.agg_out[strip_id * C + out_j[k] - 1] = ((seen >> k) & 1u) ? cA[k] : NONE32;
                }
                break;
            }

            // ---------------- chr-end rows (dap_to_bed.py:126-128,133-134)
            {
                const uint32_t n = (uint32_t)seg.rec_len;
                bool em[KPL];
                uint32_t endv[KPL];
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const uint32_t e = min(cA[k], 2u * n);
                    endv[k] = e;
                    em[k] = chr_end && out_valid[k] && e >= n;
                }
                emit(n, em, endv);
            }

            if (pass == 1) break;

            // ---------------- ordered offset: decoupled look-back over tickets
            unsigned long long my = (lg == 0) ? (unsigned long long)gcount : 0ull;
            // exclusive prefix over the groups of this warp + warp total
            unsigned long long incl = my;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned long long total = __shfl_sync(FULL, incl, 31);
            const unsigned long long gexcl = __shfl_sync(FULL, incl - my, g * G);
            if (lane == 0)
                st_status(P.status + ticket, (ticket == 0 ? ST_PREFIX : ST_AGG) | total);
            unsigned long long excl = 0;
            if (ticket > 0) {
                long long idx = ticket - 1;
                for (;;) {
                    const long long mine = idx - lane;
                    unsigned long long w = ST_PREFIX;   // before ticket 0: prefix 0
                    if (mine >= 0) {
                        do { w = ld_status(P.status + mine); } while ((w >> 62) == 0);
                    }
                    const unsigned pm = __ballot_sync(FULL, (w >> 62) == 2);
                    unsigned long long v = w & ST_VALUE;
                    if (pm) {
                        const int first = __ffs(pm) - 1;
                        if (lane > first) v = 0;
                    }
                    excl += warp_sum_u64(v);
                    if (pm) break;
                    idx -= 32;
                }
                if (lane == 0) st_status(P.status + ticket, ST_PREFIX | (excl + total));
            }
            const long long out_end_of_ticket = (long long)(excl + total);
            if (lane == 0) {
                if (last_ticket_of_seg) P.seg_out_end[s_lo] = out_end_of_ticket;
                if (ticket + 1 == P.n_tickets) P.result[MEMO_RES_N_OUT] = out_end_of_ticket;
            }
            gbase = (long long)(excl + gexcl);

            const bool overflow = __any_sync(FULL, gcount > (uint32_t)K);
            if (!overflow) {
                // flush staged rows, one group buffer at a time, coalesced
#pragma unroll 1
                for (int gi = 0; gi < NG; ++gi) {
                    const uint32_t cnt = __shfl_sync(FULL, gcount, gi * G);
                    const long long off = __shfl_sync(FULL, gbase, gi * G);
                    const uint32_t* src = wbuf + (size_t)gi * K * 3;
                    __syncwarp();
                    for (uint32_t i = lane; i < cnt; i += 32) {
                        const long long gi_out = off + i;
                        if (gi_out < P.out_cap) {
                            P.out_start[gi_out] = (int32_t)src[3 * i + 0];
                            P.out_end[gi_out] = src[3 * i + 1];
                            P.out_order[gi_out] = (int32_t)src[3 * i + 2];
                        }
                    }
                }
                __syncwarp();
                break;
            }
            // replay with direct stores
            if (lane == 0) atomicAdd((unsigned long long*)(P.result + MEMO_RES_REPLAYS), 1ull);
            direct = true;
            __syncwarp();
        }
    }
    if (__any_sync(FULL, irregular) && lane == 0) P.result[MEMO_RES_IRREGULAR] = 1;
}

// carry[strip][j] = aggregate of the closest earlier strip of the same run that
// flagged column j (NONE32 if none).  One thread per (run, column).
__global__ void carry_scan_kernel(const memo_segment_t* segs, const int64_t* seg_ticket_start,
                                  int n_seg, int NG, int R, int C, const uint32_t* agg,
                                  uint32_t* carry, const uint32_t* shard_in, uint32_t* shard_out) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tid >= (long long)n_seg * C) return;
    const int s = (int)(tid / C);
    const int j = (int)(tid % C);
    const long long first = seg_ticket_start[s] * NG;
    const long long n_strips = (segs[s].n_rows + R - 1) / R;
    uint32_t running = NONE32;
    if (s == 0 && !(segs[0].flags & MEMO_SEG_PRIMED) && shard_in != nullptr) running = shard_in[j];
    for (long long i = 0; i < n_strips; ++i) {
        carry[(first + i) * C + j] = running;
        const uint32_t a = agg[(first + i) * C + j];
        if (a != NONE32) running = a;
    }
    if (s == n_seg - 1 && shard_out != nullptr) shard_out[j] = running;
}

struct Geometry {
    int G, KPL;
};

bool pick_geometry(int C, Geometry* geo) {
    static const Geometry table[] = {
        {4, 1}, {4, 2}, {4, 3}, {8, 2}, {8, 3}, {8, 4}, {16, 3}, {16, 4},
        {32, 3}, {32, 4}, {32, 6}, {32, 8}, {32, 16},
    };
    for (const Geometry& t : table)
        if (t.G * t.KPL >= C) {
            *geo = t;
            return true;
        }
    return false;
}

typedef void (*index_kernel_t)(const IndexParams);

template <int G, int KPL>
index_kernel_t kernel_for(bool order) {
    return order ? index_kernel<G, KPL, true> : index_kernel<G, KPL, false>;
}

index_kernel_t select_kernel(const Geometry& g, bool order) {
#define MEMO_CASE(GG, KK) \
    if (g.G == GG && g.KPL == KK) return kernel_for<GG, KK>(order);
    MEMO_CASE(4, 1) MEMO_CASE(4, 2) MEMO_CASE(4, 3) MEMO_CASE(8, 2) MEMO_CASE(8, 3) MEMO_CASE(8, 4)
    MEMO_CASE(16, 3) MEMO_CASE(16, 4) MEMO_CASE(32, 3) MEMO_CASE(32, 4) MEMO_CASE(32, 6)
    MEMO_CASE(32, 8) MEMO_CASE(32, 16)
#undef MEMO_CASE
    return nullptr;
}

struct Plan {
    Geometry geo;
    int NG;
    int R, K, warps, ctas_per_sm;
    int64_t n_tickets;
    // workspace layout (byte offsets)
    size_t off_segs, off_tstart, off_status, off_counter, off_agg, off_carry, total;
};

int make_plan(int64_t rows, int32_t C, const memo_segment_t* segs, int32_t n_seg,
              const memo_index_opts_t* opts, Plan* plan, int64_t* tstart_host /* n_seg+1 or null */) {
    MEMO_REQUIRE(C >= 1, "n_cols must be >= 1");
    MEMO_REQUIRE(n_seg >= 0 && (n_seg == 0 || segs != nullptr), "bad segment table");
    if (!pick_geometry(C, &plan->geo)) {
        set_error("n_cols = %d not supported (max 512)", C);
        return MEMO_ERR_UNSUPPORTED;
    }
    plan->NG = 32 / plan->geo.G;
    plan->R = (opts && opts->rows_per_tile > 0) ? opts->rows_per_tile : 256;
    plan->warps = (opts && opts->warps_per_cta > 0) ? opts->warps_per_cta : 8;
    MEMO_REQUIRE(plan->warps >= 1 && plan->warps <= 8, "warps_per_cta must be 1..8");
    plan->ctas_per_sm = (opts && opts->ctas_per_sm > 0) ? opts->ctas_per_sm : 0;
    if (opts && opts->emit_buf_records > 0) {
        plan->K = opts->emit_buf_records;
    } else {
        // ~3x the HPRC-shaped density (0.6 % of cells), at least 32 records
        const long long cells = (long long)plan->R * C;
        long long k = cells / 48;
        if (k < 32) k = 32;
        if (k > 1024) k = 1024;
        plan->K = (int)k;
    }
    int64_t t = 0;
    int64_t prev_end = 0;
    for (int i = 0; i < n_seg; ++i) {
        const memo_segment_t& s = segs[i];
        MEMO_REQUIRE(s.n_rows > 0, "segment %d has no rows", i);
        MEMO_REQUIRE(s.row_begin >= prev_end && s.row_begin + s.n_rows <= rows,
                     "segment %d out of order or out of range", i);
        MEMO_REQUIRE((s.flags & MEMO_SEG_PRIMED) || s.row_begin >= 1,
                     "segment %d: continuation run needs a halo row before it", i);
        MEMO_REQUIRE(s.pos0 >= 0 && s.rec_len >= 0 && (int64_t)s.pos0 + s.n_rows <= 2147483647LL,
                     "segment %d: positions exceed int32", i);
        prev_end = s.row_begin + s.n_rows;
        if (tstart_host) tstart_host[i] = t;
        const int64_t per = (int64_t)plan->NG * plan->R;
        t += (s.n_rows + per - 1) / per;
    }
    if (tstart_host) tstart_host[n_seg] = t;
    plan->n_tickets = t;
    size_t off = 0;
    plan->off_segs = off;    off = align_up(off + sizeof(memo_segment_t) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
    plan->off_tstart = off;  off = align_up(off + sizeof(int64_t) * (size_t)(n_seg + 1), 256);
    plan->off_status = off;  off = align_up(off + sizeof(unsigned long long) * (size_t)(t > 0 ? t : 1), 256);
    plan->off_counter = off; off = align_up(off + 256, 256);
    const size_t strip_words = (size_t)t * plan->NG * (size_t)C;
    plan->off_agg = off;     off = align_up(off + 4 * (strip_words ? strip_words : 1), 256);
    plan->off_carry = off;   off = align_up(off + 4 * (strip_words ? strip_words : 1), 256);
    plan->total = off;
    return MEMO_OK;
}

int launch_index(const Plan& plan, IndexParams& P, bool order, cudaStream_t stream) {
    index_kernel_t kern = select_kernel(plan.geo, order);
    if (!kern) {
        set_error("no kernel for geometry G=%d KPL=%d", plan.geo.G, plan.geo.KPL);
        return MEMO_ERR_UNSUPPORTED;
    }
    const int threads = plan.warps * 32;
    const size_t smem = (size_t)plan.warps * plan.NG * plan.K * 3 * sizeof(uint32_t);
    MEMO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MEMO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) {
        set_error("index kernel does not fit on an SM (smem %zu B)", smem);
        return MEMO_ERR_UNSUPPORTED;
    }
    if (plan.ctas_per_sm > 0 && plan.ctas_per_sm < per_sm) per_sm = plan.ctas_per_sm;
    const int sms = device_sm_count();
    long long grid = (long long)sms * per_sm;
    const long long need = (plan.n_tickets + plan.warps - 1) / plan.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    MEMO_CUDA_TRY(cudaMemsetAsync(P.status, 0, sizeof(unsigned long long) * (size_t)(plan.n_tickets > 0 ? plan.n_tickets : 1), stream));
    MEMO_CUDA_TRY(cudaMemsetAsync(P.ticket_counter, 0, 256, stream));
    kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    MEMO_LAUNCH_CHECK(1);
    return MEMO_OK;
}

int build_common(bool general, const int32_t* dap, int64_t rows, int32_t C, int32_t ld,
                 const memo_segment_t* segs, int32_t n_seg, const memo_index_opts_t* opts,
                 const uint32_t* shard_carry_in, uint32_t* shard_carry_out,
                 int32_t* out_start, uint32_t* out_end, int32_t* out_order, int64_t out_cap,
                 int64_t* seg_out_end, int64_t* result, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
    MEMO_REQUIRE(rows >= 0 && ld >= C, "bad dap shape (rows=%lld, n_cols=%d, ld=%d)", (long long)rows, C, ld);
    MEMO_REQUIRE(result != nullptr, "result must not be NULL");
    MEMO_REQUIRE(out_cap == 0 || (out_start && out_end && out_order), "out_* NULL with out_cap > 0");
    MEMO_REQUIRE((reinterpret_cast<uintptr_t>(dap) & 15) == 0, "dap must be 16-byte aligned");
    MEMO_REQUIRE(n_seg == 0 || seg_out_end != nullptr, "seg_out_end must not be NULL");
    Plan plan;
    int64_t* tstart = new int64_t[(size_t)n_seg + 1];
    int rc = make_plan(rows, C, segs, n_seg, opts, &plan, tstart);
    if (rc != MEMO_OK) { delete[] tstart; return rc; }
    if (workspace_bytes < plan.total || workspace == nullptr) {
        delete[] tstart;
        set_error("workspace too small: %zu < %zu", workspace_bytes, plan.total);
        return MEMO_ERR_WORKSPACE;
    }
    cudaError_t e0 = cudaMemsetAsync(result, 0, sizeof(int64_t) * MEMO_RES_SLOTS, stream);
    if (e0 != cudaSuccess) { delete[] tstart; set_error("memset result: %s", cudaGetErrorString(e0)); return MEMO_ERR_CUDA; }
    if (n_seg == 0 || plan.n_tickets == 0) { delete[] tstart; return MEMO_OK; }

    char* ws = static_cast<char*>(workspace);
    // pageable host -> device copies are staged by the runtime before returning,
    // so the host buffers may be released right after these calls
    cudaError_t e1 = cudaMemcpyAsync(ws + plan.off_segs, segs, sizeof(memo_segment_t) * (size_t)n_seg,
                                     cudaMemcpyHostToDevice, stream);
    cudaError_t e2 = cudaMemcpyAsync(ws + plan.off_tstart, tstart, sizeof(int64_t) * (size_t)(n_seg + 1),
                                     cudaMemcpyHostToDevice, stream);
    delete[] tstart;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("segment table upload failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return MEMO_ERR_CUDA;
    }
    const bool order = opts ? (opts->order_mode != 0) : true;

    IndexParams P{};
    P.dap = dap; P.rows = rows; P.C = C; P.ld = ld;
    P.segs = reinterpret_cast<const memo_segment_t*>(ws + plan.off_segs);
    P.seg_ticket_start = reinterpret_cast<const int64_t*>(ws + plan.off_tstart);
    P.n_seg = n_seg; P.n_tickets = plan.n_tickets; P.R = plan.R; P.K = plan.K;
    P.out_start = out_start; P.out_end = out_end; P.out_order = out_order; P.out_cap = out_cap;
    P.seg_out_end = seg_out_end;
    P.status = reinterpret_cast<unsigned long long*>(ws + plan.off_status);
    P.ticket_counter = reinterpret_cast<unsigned long long*>(ws + plan.off_counter);
    P.result = result;
    P.agg_out = reinterpret_cast<uint32_t*>(ws + plan.off_agg);
    P.carry_in = nullptr;
    P.mode = 0;

    if (general) {
        P.mode = 1;
        rc = launch_index(plan, P, order, stream);
        if (rc != MEMO_OK) return rc;
        uint32_t* carry = reinterpret_cast<uint32_t*>(ws + plan.off_carry);
        const long long nthreads = (long long)n_seg * C;
        carry_scan_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, stream>>>(
            P.segs, P.seg_ticket_start, n_seg, plan.NG, plan.R, C, P.agg_out, carry,
            shard_carry_in, shard_carry_out);
        MEMO_LAUNCH_CHECK(1);
        P.mode = 0;
        P.carry_in = carry;
    }
    return launch_index(plan, P, order, stream);
}

}  // namespace
}  // namespace memo

namespace memo {
size_t general_workspace_bytes(int64_t rows, int32_t n_cols, const memo_segment_t* segs,
                               int32_t n_seg, const memo_index_opts_t* opts) {
    Plan plan;
    if (make_plan(rows, n_cols, segs, n_seg, opts, &plan, nullptr) != MEMO_OK) return 0;
    return plan.total;
}
}  // namespace memo

extern "C" {

int memo_index_build_general(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld,
                             const memo_segment_t* segs, int32_t n_seg,
                             const memo_index_opts_t* opts, const uint32_t* shard_carry_in,
                             uint32_t* shard_carry_out, int32_t* out_start, uint32_t* out_end,
                             int32_t* out_order, int64_t out_cap, int64_t* seg_out_end,
                             int64_t* result, void* workspace, size_t workspace_bytes,
                             void* stream) {
    return memo::build_common(true, dap, rows, n_cols, ld, segs, n_seg, opts, shard_carry_in,
                              shard_carry_out, out_start, out_end, out_order, out_cap,
                              seg_out_end, result, workspace, workspace_bytes,
                              static_cast<cudaStream_t>(stream));
}

}  // extern "C"
