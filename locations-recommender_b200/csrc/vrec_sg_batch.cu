// SG batch path: StochasticRecommender.makeRecommendations for MANY start vertices of one graph
// (BASELINE config 4, "all persons"), stochastic/StochasticRecommender.scala:66-141 +
// stochastic/StochasticRecommenderMain.scala:64-76.
//
// Structure used (exact, not an approximation).  Let Z be the vertices without in-edges -- in the
// reference's graphs every person vertex: the edge families are person->place, person->category,
// place->place, category->place (stochastic/StochasticGraphBuilder.scala:8-28).  For a start vertex
// t in Z:
//   * iteration 0 starts from x0 = 1/N and gives x1[a] = 0.85 * sigma(x0)[a] for every vertex a with
//     in-edges -- the same for every t; x1[t] = 0.15, x1[z] = 0 for the other z in Z;
//   * from then on x[z] = 0 for z in Z \ {t} and x[t] = 0.15, so a row's sum only has the terms of
//     its sources with in-edges ("active" vertices A, ~10^4 places + categories) plus the start
//     vertex's own edge.  Terms x[z] * w = +0.0 leave an fp64 sum unchanged, so dropping them keeps
//     the canonical summation order bit for bit: a term keeps the lane (k % 32) and segment
//     (k / 1024) of its position k in the FULL row.
// The path requires the active sources of every row to precede its Z sources (ids of categories and
// places are smaller than person ids in the reference's data: SampleGeneratorMain.scala:36-37,54);
// otherwise, or for start vertices with in-edges, vrec_sg_query falls back to the per-query kernels.
//
// One CTA owns T start vertices at a time: their x vectors over A live in shared memory, the
// reduced graph (10 B per edge, L2-resident) is streamed once per iteration for all T, and the
// step() loop (:92-106), the convergence test (:130-141) and the ranked top-N all run inside the
// kernel -- no launch per iteration or per query.
//
// Layout of the reduced graph: the active vertices are renumbered by ceil(in-edges / 32) and stored
// in slices of 32 rows (sliced ELLPACK, one row per lane, coalesced).  Inside a slice a row's terms
// are stored canonical lane by canonical lane (term k -> step (k % 32) * R + k / 32, R = terms per
// lane), so ONE THREAD sums a whole row in exactly the canonical order: the 32 lane sums one after
// the other, combined by a binary-counter stack that reproduces the xor-butterfly tree
// ((L0+L1)+(L2+L3))+... -- no shuffles, no dependent cross-lane latency.  Padding terms are
// x[0] * 0.0 = +0.0.  Rows with more than 1024 active in-edges keep the CSR form and a whole warp.
#include <algorithm>
#include <chrono>
#include <climits>

#include "vrec_sg.cuh"

namespace {

constexpr int BT = 768;             // threads per CTA
constexpr int BW = BT / 32;
constexpr int TOPN_FAST = 16;       // max_recs up to this use the warp-local selection
constexpr size_t SMEM_LIMIT = 227 * 1024 - 12 * 1024;   // dynamic part; ~9 KB of statics + reserve

struct SgBatchArgs {
    int n_a, n_chunks;
    int n_fast_chunks, n_slow;        // slices summed one row per thread; rows behind them, one per warp
    // sliced layout: slice c holds rows [32c, 32c+32); chunk_r[c] = terms of its longest row (> 1024: CSR path)
    const int *chunk_r;
    const long long *chunk_ptr;
    const unsigned short *sell_src;
    const double *sell_w;
    // CSR of the same reduced graph (rows that receive the start vertex's term, rows > 1024 terms)
    const int *r_rowptr, *r_src, *full_len;
    const double *r_w;
    const int *z_rowptr, *z_row, *z_pos;
    const double *z_w;
    const double *x1a;
    const long long *act_id;
    const int *q_vertex;
    int n_q, max_it;
    double eps2;
    const unsigned *cand_mask;        // bit per active vertex: member of the place filter
    int max_recs;
    long long *out_id;
    double *out_prob;
    int *out_count, *out_it, *out_conv;
    double *scratch;
    int *counter;
};

__device__ __forceinline__ unsigned ld_nc_u16(const unsigned short *p) {
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// x of the T start vertices of a CTA is interleaved in shared memory (xs[i * T + t]): one 128-bit
// load fetches both values of a source when T = 2 (fewer bank-conflict wavefronts than two 64-bit loads).
template <int T>
__device__ __forceinline__ void load_x(const double *xs, unsigned c, double (&xv)[T]) {
    if constexpr (T == 2) {
        const double2 v = *reinterpret_cast<const double2 *>(xs + 2 * c);
        xv[0] = v.x;
        xv[1] = v.y;
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) xv[t] = xs[c * T + t];
    }
}

// One thread, one row of a slice whose longest row has n_max = 32 * (R - 1) + rem terms (1 <= rem <= 32):
// canonical lane l has R terms if l < rem, else R - 1.  Storage: first the (R - 1) leading terms of every
// lane, lane by lane (step l * (R - 1) + c2), then the rem last terms (step 32 * (R - 1) + l).  The
// lanes are summed 0..31 in turn and merged by a binary-counter stack = the xor-butterfly tree.
__host__ __device__ constexpr int trailing_ones(int l) { return (l & 1) ? 1 + trailing_ones(l >> 1) : 0; }

template <int T, int L>
__device__ __forceinline__ void tree_push(double (&st)[T][6], const double (&acc)[T]) {
    constexpr int TZ = trailing_ones(L);
#pragma unroll
    for (int t = 0; t < T; ++t) {
        double v = acc[t];
#pragma unroll
        for (int bit = 0; bit < TZ; ++bit) v = xadd(st[t][bit], v);
        st[t][TZ] = v;
    }
}

// lanes [L0, L0 + GL) of the row: all their loads first (GL * R <= 8 in flight), then the sums
template <int T, int R, int GL, int L0>
struct SellGroup {
    static __device__ __forceinline__ void run(const unsigned short *__restrict__ src, const double *__restrict__ w,
                                               const double *xs, int n_a, int rem, double (&st)[T][6]) {
        unsigned c[GL][R];
        double ww[GL][R];
#pragma unroll
        for (int g = 0; g < GL; ++g) {
            const int l = L0 + g;
#pragma unroll
            for (int c2 = 0; c2 < R - 1; ++c2) {
                c[g][c2] = ld_nc_u16(src + (l * (R - 1) + c2) * 32);
                ww[g][c2] = ld_nc_f64(w + (l * (R - 1) + c2) * 32);
            }
            const bool last = l < rem;
            c[g][R - 1] = last ? ld_nc_u16(src + (32 * (R - 1) + l) * 32) : 0u;
            ww[g][R - 1] = last ? ld_nc_f64(w + (32 * (R - 1) + l) * 32) : 0.0;
        }
        tail<0>(c, ww, xs, n_a, rem, st);
        SellGroup<T, R, GL, L0 + GL>::run(src, w, xs, n_a, rem, st);
    }
    template <int G>
    static __device__ __forceinline__ void tail(const unsigned (&c)[GL][R], const double (&ww)[GL][R], const double *xs,
                                                int n_a, int rem, double (&st)[T][6]) {
        if constexpr (G < GL) {
            double acc[T];
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = 0.0;
#pragma unroll
            for (int c2 = 0; c2 < R - 1; ++c2) {
                double xv[T];
                load_x<T>(xs, c[G][c2], xv);
#pragma unroll
                for (int t = 0; t < T; ++t) acc[t] = xadd(acc[t], xmul(xv[t], ww[G][c2]));
            }
            if (L0 + G < rem) {
                double xv[T];
                load_x<T>(xs, c[G][R - 1], xv);
#pragma unroll
                for (int t = 0; t < T; ++t) acc[t] = xadd(acc[t], xmul(xv[t], ww[G][R - 1]));
            }
            tree_push<T, L0 + G>(st, acc);
            tail<G + 1>(c, ww, xs, n_a, rem, st);
        }
    }
};
template <int T, int R, int GL>
struct SellGroup<T, R, GL, 32> {
    static __device__ __forceinline__ void run(const unsigned short *, const double *, const double *, int, int,
                                               double (&)[T][6]) {}
};

template <int T, int R>
__device__ __forceinline__ void sell_row_sum(const unsigned short *__restrict__ src, const double *__restrict__ w,
                                             const double *xs, int n_a, int rem, double (&out)[T]) {
    constexpr int GL = R == 1 ? 8 : (R == 2 ? 4 : 2);
    double st[T][6];
    SellGroup<T, R, GL, 0>::run(src, w, xs, n_a, rem, st);
#pragma unroll
    for (int t = 0; t < T; ++t) out[t] = st[t][5];
}

// The same for any R <= 32 (rows with more than 128 active in-edges): runtime inner loop, the
// canonical lane L as a template parameter so that the tree stack stays in registers.
template <int T, int L>
struct SellLane {
    static __device__ __forceinline__ void run(const unsigned short *__restrict__ src, const double *__restrict__ w,
                                               const double *xs, int n_a, int R, int rem, double (&st)[T][6]) {
        double acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.0;
        const unsigned short *ps = src + (size_t)L * (R - 1) * 32;
        const double *pw = w + (size_t)L * (R - 1) * 32;
#pragma unroll 4
        for (int c2 = 0; c2 < R - 1; ++c2) {
            unsigned c = ld_nc_u16(ps + c2 * 32);
            double ww = ld_nc_f64(pw + c2 * 32);
            double xv[T];
            load_x<T>(xs, c, xv);
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = xadd(acc[t], xmul(xv[t], ww));
        }
        if (L < rem) {
            unsigned c = ld_nc_u16(src + (size_t)(32 * (R - 1) + L) * 32);
            double ww = ld_nc_f64(w + (size_t)(32 * (R - 1) + L) * 32);
            double xv[T];
            load_x<T>(xs, c, xv);
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = xadd(acc[t], xmul(xv[t], ww));
        }
        tree_push<T, L>(st, acc);
        SellLane<T, L + 1>::run(src, w, xs, n_a, R, rem, st);
    }
};
template <int T>
struct SellLane<T, 32> {
    static __device__ __forceinline__ void run(const unsigned short *, const double *, const double *, int, int, int,
                                               double (&)[T][6]) {}
};

template <int T>
__device__ __forceinline__ void sell_row_sum_any(const unsigned short *__restrict__ src, const double *__restrict__ w,
                                                 const double *xs, int n_a, int R, int rem, double (&out)[T]) {
    double st[T][6];
    SellLane<T, 0>::run(src, w, xs, n_a, R, rem, st);
#pragma unroll
    for (int t = 0; t < T; ++t) out[t] = st[t][5];
}

// Lane-strided partial sums of the CSR terms [s, s+len): this lane's canonical lane sum (terms
// k = lane, lane+32, ... in ascending order), four loads in flight.  NT x vectors with values at
// xs[c * S + t], t < NT (S = vectors interleaved in shared memory).
template <int NT, int S>
__device__ __forceinline__ void csr_lane_sum(const SgBatchArgs &a, const double *xs, int s, int len, int lane,
                                             double (&acc)[NT]) {
    const int *__restrict__ src = a.r_src + s;
    const double *__restrict__ w = a.r_w + s;
    int k = lane;
    for (; k + 96 < len; k += 128) {
        int c[4];
        double ww[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c[q] = __ldg(src + k + 32 * q);
            ww[q] = __ldg(w + k + 32 * q);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double xv[NT];
            if constexpr (NT == S) {
                load_x<NT>(xs, (unsigned)c[q], xv);
            } else {
#pragma unroll
                for (int t = 0; t < NT; ++t) xv[t] = xs[c[q] * S + t];
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t] = xadd(acc[t], xmul(xv[t], ww[q]));
        }
    }
    for (; k < len; k += 32) {
        const int c = __ldg(src + k);
        const double ww = __ldg(w + k);
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t] = xadd(acc[t], xmul(xs[c * S + t], ww));
    }
}

// Row sum by a whole warp, for T x vectors: the reduced prefix of `row` (m terms at positions
// 0..m-1 of the full row).  Rows whose FULL length exceeds VREC_CANON_SEG are summed per 1024-term
// segment and the segment sums by the same lane rule; all-zero segments are
// skipped.
template <int T>
__device__ void warp_prefix_sigma(const SgBatchArgs &a, const double *xs, int row, int lane, double (&out)[T]) {
    const int s = a.r_rowptr[row], m = a.r_rowptr[row + 1] - s;
    const int fl = a.full_len[row];
    double acc[T];
    if (fl <= VREC_CANON_SEG) {
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.0;
        csr_lane_sum<T, T>(a, xs, s, m, lane, acc);
#pragma unroll
        for (int t = 0; t < T; ++t) out[t] = canon_butterfly(acc[t]);
        return;
    }
    const int nsp = (m + VREC_CANON_SEG - 1) / VREC_CANON_SEG;
    double acc2[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc2[t] = 0.0;
    for (int j = 0; j < nsp; ++j) {
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.0;
        csr_lane_sum<T, T>(a, xs, s + j * VREC_CANON_SEG, min(VREC_CANON_SEG, m - j * VREC_CANON_SEG), lane, acc);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            double part = canon_butterfly(acc[t]);
            if ((j & 31) == lane) acc2[t] = xadd(acc2[t], part);
        }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) out[t] = canon_butterfly(acc2[t]);
}

// The same for one x vector (values at xs[c * S]) plus the start vertex's own terms z[e..ee) (positions z_pos >= m,
// ascending, value 0.15 * w): they follow the prefix terms of their canonical lane and segment.
template <int S>
__device__ double warp_row_sigma(const SgBatchArgs &a, const double *xs, int row, int e, int ee, int lane) {
    const int s = a.r_rowptr[row], m = a.r_rowptr[row + 1] - s;
    const int fl = a.full_len[row];
    double acc[1];
    if (fl <= VREC_CANON_SEG) {
        acc[0] = 0.0;
        csr_lane_sum<1, S>(a, xs, s, m, lane, acc);
        for (int q = e; q < ee; ++q)
            if ((a.z_pos[q] & 31) == lane) acc[0] = xadd(acc[0], xmul(kAlpha, a.z_w[q]));
        return canon_butterfly(acc[0]);
    }
    const int nsp = (m + VREC_CANON_SEG - 1) / VREC_CANON_SEG;
    double acc2 = 0.0;
    int q = e;
    for (int j = 0; j < nsp; ++j) {
        acc[0] = 0.0;
        csr_lane_sum<1, S>(a, xs, s + j * VREC_CANON_SEG, min(VREC_CANON_SEG, m - j * VREC_CANON_SEG), lane, acc);
        while (q < ee && a.z_pos[q] / VREC_CANON_SEG == j) {
            if ((a.z_pos[q] & 31) == lane) acc[0] = xadd(acc[0], xmul(kAlpha, a.z_w[q]));
            ++q;
        }
        double part = canon_butterfly(acc[0]);
        if ((j & 31) == lane) acc2 = xadd(acc2, part);
    }
    while (q < ee) {
        const int j = a.z_pos[q] / VREC_CANON_SEG;
        double accz = 0.0;
        while (q < ee && a.z_pos[q] / VREC_CANON_SEG == j) {
            if ((a.z_pos[q] & 31) == lane) accz = xadd(accz, xmul(kAlpha, a.z_w[q]));
            ++q;
        }
        double part = canon_butterfly(accz);
        if ((j & 31) == lane) acc2 = xadd(acc2, part);
    }
    return canon_butterfly(acc2);
}

// (value desc, id asc) with the id looked up only when the values tie
__device__ __forceinline__ bool act_before(const SgBatchArgs &a, double va, int ia, double vb, int ib) {
    if (va != vb) return va > vb;
    if (ia < 0 || ib < 0) return ib < 0 && ia >= 0;
    return a.act_id[ia] < a.act_id[ib];
}

// warp-wide best of (v, i) pairs; i < 0 = none
__device__ __forceinline__ void warp_best(const SgBatchArgs &a, double &bv, int &bi) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (oi >= 0 && (bi < 0 || act_before(a, ov, oi, bv, bi))) {
            bv = ov;
            bi = oi;
        }
    }
}

template <int T>
__global__ void __launch_bounds__(BT, 1) sg_batch_kernel(const SgBatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_a = a.n_a, n_chunks = a.n_chunks;
    double *xs = reinterpret_cast<double *>(smem_raw);                      // [T][n_a]
    double *chunk_res = xs + (size_t)T * n_a;                               // [T][n_chunks]
    unsigned *flag = reinterpret_cast<unsigned *>(chunk_res + (size_t)T * n_chunks);   // [T][n_chunks] row bitmaps
    unsigned *cmask = flag + (size_t)T * n_chunks;                          // [n_chunks] place-filter bitmap
    double *slow_res = reinterpret_cast<double *>(cmask + n_chunks + (((T + 1) * n_chunks) & 1));   // [T][n_slow], 8-byte aligned
    const int n_fast = a.n_fast_chunks, n_slow = a.n_slow;
    __shared__ int s_grp, s_next_chunk;
    __shared__ int s_done[T], s_iter[T], s_conv[T], s_copy[T];
    __shared__ double s_lv[BW * TOPN_FAST];
    __shared__ int s_li[BW * TOPN_FAST];
    __shared__ double s_sel_v;
    __shared__ int s_sel_i, s_list_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double one_minus = 1 - kAlpha;                                    // :121
    double *nx = a.scratch + (size_t)blockIdx.x * T * n_a;                  // x' of this CTA's start vertices
    const int n_groups = (a.n_q + T - 1) / T;
    for (int i = tid; i < n_chunks; i += BT) cmask[i] = a.cand_mask[i];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_grp = atomicAdd(a.counter, 1);
        __syncthreads();
        const int grp = s_grp;
        if (grp >= n_groups) break;
        const int nt = min(T, a.n_q - grp * T);
        for (int i = tid; i < n_a; i += BT) {
            double v = a.x1a[i];
#pragma unroll
            for (int t = 0; t < T; ++t) xs[i * T + t] = v;
        }
        if (tid < T) {
            s_done[tid] = tid >= nt;
            s_iter[tid] = 0;
            s_conv[tid] = 0;
        }
        __syncthreads();
        // step(), :92-106; iteration 0 (x0 -> x1) is the shared x1a, not converged (checked on the host)
        for (int it = 1;; ++it) {
            if (it >= a.max_it) {                                           // :93-95
                if (tid < T && !s_done[tid]) {
                    s_iter[tid] = a.max_it;
                    s_conv[tid] = 0;
                    s_done[tid] = 1;
                }
                __syncthreads();
                break;
            }
            for (int i = tid; i < T * n_chunks; i += BT) flag[i] = 0u;
            if (tid == 0) s_next_chunk = 0;
            __syncthreads();
            // rows that receive the start vertex's own term
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (t >= nt || s_done[t]) continue;
                const int z = a.q_vertex[grp * T + t];
                const int e0 = a.z_rowptr[z], e1 = a.z_rowptr[z + 1];
                for (int e = e0 + warp; e < e1; e += BW) {
                    const int row = a.z_row[e];
                    if (e > e0 && a.z_row[e - 1] == row) continue;          // duplicate edge: summed with its first
                    int ee = e + 1;
                    while (ee < e1 && a.z_row[ee] == row) ++ee;
                    double sigma = warp_row_sigma<T>(a, xs + t, row, e, ee, lane);
                    if (lane == 0) {
                        nx[t * n_a + row] = xadd(xmul(0.0, kAlpha), xmul(sigma, one_minus));   // :120-122
                        atomicOr(&flag[t * n_chunks + (row >> 5)], 1u << (row & 31));
                    }
                }
            }
            __syncthreads();
            // every other row: calcNextX (:108-128) over the reduced graph.  Work items, longest first:
            // the rows behind the fast slices (one warp per row), then one slice of 32 rows per warp turn.
            for (;;) {
                int item = 0;
                if (lane == 0) item = atomicAdd(&s_next_chunk, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= n_slow + n_fast) break;
                if (item < n_slow) {
                    const int r = n_fast * 32 + item;
                    double sg[T];
                    warp_prefix_sigma<T>(a, xs, r, lane, sg);
                    if (lane == 0) {
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const bool own = (flag[t * n_chunks + (r >> 5)] >> (r & 31)) & 1u;
                            double v;
                            if (own) {
                                v = nx[t * n_a + r];
                            } else {
                                v = xadd(xmul(0.0, kAlpha), xmul(sg[t], one_minus));
                                nx[t * n_a + r] = v;
                            }
                            double d = xsub(v, xs[r * T + t]);
                            slow_res[t * n_slow + item] = xmul(d, d);
                        }
                    }
                    continue;
                }
                const int c = item - n_slow;
                const int r = c * 32 + lane;
                const int n_max = a.chunk_r[c];                  // terms of the longest row of the slice
                const int R = (n_max + 31) >> 5, rem = n_max - 32 * (R - 1);
                double sigma[T];
#pragma unroll
                for (int t = 0; t < T; ++t) sigma[t] = 0.0;
                if (R > 0) {
                    const long long base = a.chunk_ptr[c] + lane;
                    const unsigned short *ps = a.sell_src + base;
                    const double *pw = a.sell_w + base;
                    switch (R) {
                        case 1: sell_row_sum<T, 1>(ps, pw, xs, n_a, rem, sigma); break;
                        case 2: sell_row_sum<T, 2>(ps, pw, xs, n_a, rem, sigma); break;
                        case 3: sell_row_sum<T, 3>(ps, pw, xs, n_a, rem, sigma); break;
                        case 4: sell_row_sum<T, 4>(ps, pw, xs, n_a, rem, sigma); break;
                        default: sell_row_sum_any<T>(ps, pw, xs, n_a, R, rem, sigma); break;
                    }
                }
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    double sq = 0.0;
                    if (r < n_a) {
                        const bool own = (flag[t * n_chunks + c] >> lane) & 1u;
                        double v;
                        if (own) {
                            v = nx[t * n_a + r];
                        } else {
                            v = xadd(xmul(0.0, kAlpha), xmul(sigma[t], one_minus));
                            nx[t * n_a + r] = v;
                        }
                        double d = xsub(v, xs[r * T + t]);                    // isConverged, :131-139
                        sq = xmul(d, d);
                    }
                    sq = canon_butterfly(sq);
                    if (lane == 0) chunk_res[t * n_chunks + c] = sq;
                }
            }
            __syncthreads();
            if (warp < T) {
                const int t = warp;
                double acc = 0.0;
                for (int k = lane; k < n_fast; k += 32) acc = xadd(acc, chunk_res[t * n_chunks + k]);
                for (int k = lane; k < n_slow; k += 32) acc = xadd(acc, slow_res[t * n_slow + k]);
                acc = canon_butterfly(acc);
                if (lane == 0) {
                    s_copy[t] = !s_done[t];
                    if (!s_done[t] && acc <= a.eps2) {                      // :140
                        s_conv[t] = 1;
                        s_iter[t] = it;                                     // :100
                        s_done[t] = 1;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (!s_copy[t]) continue;
                for (int i = tid; i < n_a; i += BT) xs[i * T + t] = nx[t * n_a + i];
            }
            __syncthreads();
            bool all = true;
#pragma unroll
            for (int t = 0; t < T; ++t) all = all && s_done[t];
            if (all) break;
        }
        // makeRecommendations0's filter (:85-88) + printRecommendations' ranked top-N (Main :69-73).
        // Candidates = active vertices in the place filter with probability > 0.
        for (int t = 0; t < nt; ++t) {
            const int q = grp * T + t;
            const double *xt = xs + t;                           // values at xt[i * T]
            int count = 0;
            bool ranked = false;
            if (a.max_recs > 0 && a.max_recs <= TOPN_FAST) {
                // (1) every thread's largest candidate; (2) the max_recs-th largest of those 512 maxima
                // is a lower bound of the max_recs-th best (each maximum is a distinct candidate);
                // (3) the few candidates at or above it are listed; (4) warp 0 ranks the list exactly.
                const int K = a.max_recs;
                double mx = 0.0;
                for (int base = warp * 32; base < n_a; base += BT) {
                    const int i = base + lane;
                    const double v = (i < n_a && ((cmask[base >> 5] >> lane) & 1u)) ? xt[i * T] : 0.0;
                    mx = v > mx ? v : mx;
                }
                for (int r = 0; r < K; ++r) {                     // warp-local top K (values, with multiplicity)
                    double wm = mx;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        double o = __shfl_xor_sync(0xffffffffu, wm, off);
                        wm = o > wm ? o : wm;
                    }
                    const unsigned holders = __ballot_sync(0xffffffffu, mx == wm);
                    if (lane == __ffs(holders) - 1) mx = -1.0;     // remove one instance
                    if (lane == 0) s_lv[warp * TOPN_FAST + r] = wm;
                }
                if (tid == 0) s_list_n = 0;
                __syncthreads();
                if (warp == 0) {
                    double e[BW * TOPN_FAST / 32];
#pragma unroll
                    for (int j = 0; j < BW * TOPN_FAST / 32; ++j) {
                        const int idx = lane + 32 * j;
                        e[j] = (idx % TOPN_FAST) < K ? s_lv[idx] : -1.0;
                    }
                    double theta = 0.0;
                    for (int r = 0; r < K; ++r) {
                        double lm = e[0];
#pragma unroll
                        for (int j = 1; j < BW * TOPN_FAST / 32; ++j) lm = e[j] > lm ? e[j] : lm;
                        double wm = lm;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            double o = __shfl_xor_sync(0xffffffffu, wm, off);
                            wm = o > wm ? o : wm;
                        }
                        theta = wm;
                        const unsigned holders = __ballot_sync(0xffffffffu, lm == wm);
                        if (lane == __ffs(holders) - 1) {
                            bool gone = false;
#pragma unroll
                            for (int j = 0; j < BW * TOPN_FAST / 32; ++j)
                                if (!gone && e[j] == wm) {
                                    e[j] = -1.0;
                                    gone = true;
                                }
                        }
                    }
                    if (lane == 0) s_sel_v = theta;                // <= 0: fewer than K positive candidates
                }
                __syncthreads();
                const double theta = s_sel_v;
                __syncthreads();                                   // s_lv is reused as the list below
                for (int base = warp * 32; base < n_a; base += BT) {
                    const int i = base + lane;
                    if (i < n_a && ((cmask[base >> 5] >> lane) & 1u)) {
                        const double v = xt[i * T];
                        if (v > 0 && v >= theta) {
                            const int pos = atomicAdd(&s_list_n, 1);
                            if (pos < BW * TOPN_FAST) {
                                s_lv[pos] = v;
                                s_li[pos] = i;
                            }
                        }
                    }
                }
                __syncthreads();
                const int n_list = s_list_n;
                if (n_list <= BW * TOPN_FAST) {
                    ranked = true;
                    if (warp == 0) {
                        double ev[BW * TOPN_FAST / 32];
                        int ei[BW * TOPN_FAST / 32];
#pragma unroll
                        for (int j = 0; j < BW * TOPN_FAST / 32; ++j) {
                            const int idx = lane + 32 * j;
                            ei[j] = idx < n_list ? s_li[idx] : -1;
                            ev[j] = idx < n_list ? s_lv[idx] : 0.0;
                        }
                        for (int r = 0; r < K; ++r) {
                            double bv = 0.0;
                            int bi = -1;
#pragma unroll
                            for (int j = 0; j < BW * TOPN_FAST / 32; ++j)
                                if (ei[j] >= 0 && (bi < 0 || act_before(a, ev[j], ei[j], bv, bi))) {
                                    bv = ev[j];
                                    bi = ei[j];
                                }
                            warp_best(a, bv, bi);
                            if (bi < 0) break;
#pragma unroll
                            for (int j = 0; j < BW * TOPN_FAST / 32; ++j)
                                if (ei[j] == bi) ei[j] = -1;
                            if (lane == 0) {
                                a.out_id[(size_t)q * a.max_recs + r] = a.act_id[bi];
                                a.out_prob[(size_t)q * a.max_recs + r] = bv;
                            }
                            ++count;
                        }
                        if (lane == 0) s_sel_i = count;
                    }
                    __syncthreads();
                    count = s_sel_i;
                }
            }
            if (!ranked) {
                // any max_recs: one block-wide pick per rank
                double pv = 0.0;
                int pi = -1;
                for (int r = 0; r < a.max_recs; ++r) {
                    double bv = 0.0;
                    int bi = -1;
                    for (int base = warp * 32; base < n_a; base += BT) {
                        const int i = base + lane;
                        if (i < n_a && ((cmask[base >> 5] >> lane) & 1u)) {
                            double v = xt[i * T];
                            if (v > 0 && (pi < 0 || act_before(a, pv, pi, v, i)) && (bi < 0 || act_before(a, v, i, bv, bi))) {
                                bv = v;
                                bi = i;
                            }
                        }
                    }
                    warp_best(a, bv, bi);
                    if (lane == 0) {
                        s_lv[warp] = bv;
                        s_li[warp] = bi;
                    }
                    __syncthreads();
                    if (warp == 0) {
                        bv = lane < BW ? s_lv[lane] : 0.0;
                        bi = lane < BW ? s_li[lane] : -1;
                        warp_best(a, bv, bi);
                        if (lane == 0) {
                            s_sel_v = bv;
                            s_sel_i = bi;
                        }
                    }
                    __syncthreads();
                    pv = s_sel_v;
                    pi = s_sel_i;
                    if (pi < 0) break;
                    if (tid == 0) {
                        a.out_id[(size_t)q * a.max_recs + r] = a.act_id[pi];
                        a.out_prob[(size_t)q * a.max_recs + r] = pv;
                    }
                    ++count;
                }
            }
            if (tid == 0) {
                a.out_count[q] = count;
                a.out_it[q] = s_iter[t];
                a.out_conv[q] = s_conv[t];
            }
            __syncthreads();
        }
    }
}

__global__ void sg_gather_active_kernel(const double *__restrict__ x, const int *__restrict__ act_vertex, int n_a,
                                        double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_a) out[i] = x[act_vertex[i]];
}

size_t batch_smem(int T, int n_a, int n_slow) {
    size_t n_chunks = (size_t)(n_a + 31) / 32;
    return (size_t)T * n_a * 8 + (size_t)T * n_chunks * 8 + (size_t)T * n_chunks * 4 + (n_chunks + 1) * 4 +
           (size_t)T * n_slow * 8 + 16;
}

template <int T>
int launch_batch(vrec_ctx *ctx, const SgBatchArgs &a, int grid, size_t smem) {
    VREC_CUDA(cudaFuncSetAttribute(sg_batch_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sg_batch_kernel<T><<<grid, BT, smem, ctx->stream>>>(a);
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}

}  // namespace

// Load-time analysis of a host-built graph: active set (renumbered by row length), reduced graph in
// sliced and CSR form, out-edges of the in-degree-0 vertices with their positions in the full rows.
// Leaves batch.ok = false when the graph does not have the required shape.
int sg_batch_analyse(vrec_sg *g, const std::vector<int> &rowptr, const int *h_src, const double *h_w) {
    SgBatch &b = g->batch;
    b.analysed = true;
    b.ok = false;
    const int64_t N = g->N;
    if (N <= 0 || g->partitioned || g->nblocks != 1 || g->row_lo != 0 || g->row_hi != N) return VREC_OK;
    std::vector<char> active((size_t)N, 0);
    int n_a = 0;
    for (int64_t v = 0; v < N; ++v)
        if (rowptr[v + 1] > rowptr[v]) {
            active[v] = 1;
            ++n_a;
        }
    if (n_a == 0 || n_a == N || batch_smem(1, n_a, 0) > SMEM_LIMIT) return VREC_OK;
    // active prefix of every row; rows ordered by terms per canonical lane
    std::vector<int> act_vertex, m_of((size_t)N, 0);
    act_vertex.reserve((size_t)n_a);
    for (int64_t v = 0; v < N; ++v) {
        if (!active[v]) continue;
        const int s = rowptr[v], e = rowptr[v + 1];
        int m = 0;
        while (s + m < e && active[h_src[s + m]]) ++m;
        for (int k = s + m; k < e; ++k)
            if (active[h_src[k]]) return VREC_OK;              // an active source after a Z source: not this shape
        m_of[v] = m;
        act_vertex.push_back((int)v);
    }
    // rows by length, so that the rows of a slice are (almost) equally long
    std::stable_sort(act_vertex.begin(), act_vertex.end(), [&](int x, int y) { return m_of[x] < m_of[y]; });
    b.h_act_of.assign((size_t)N, -1);
    for (int a = 0; a < n_a; ++a) b.h_act_of[act_vertex[a]] = a;
    const int n_chunks = (n_a + 31) / 32;
    std::vector<int> r_rowptr((size_t)n_a + 1, 0), full_len((size_t)n_a), zcnt((size_t)N + 1, 0), chunk_r((size_t)n_chunks, 0);
    std::vector<long long> chunk_ptr((size_t)n_chunks + 1, 0), act_id((size_t)n_a);
    for (int a = 0; a < n_a; ++a) {
        const int v = act_vertex[a], s = rowptr[v], e = rowptr[v + 1];
        r_rowptr[a + 1] = r_rowptr[a] + m_of[v];
        full_len[a] = e - s;
        act_id[a] = (long long)g->h_ids[v];
        chunk_r[a >> 5] = std::max(chunk_r[a >> 5], m_of[v]);
        for (int k = s + m_of[v]; k < e; ++k) zcnt[h_src[k] + 1]++;
    }
    int n_fast_chunks = n_chunks;                            // slices in front of the first row with > 1024 terms
    for (int c = n_chunks - 1; c >= 0 && chunk_r[c] > VREC_CANON_SEG; --c) n_fast_chunks = c;
    const int n_slow = n_a - std::min(n_a, n_fast_chunks * 32);
    if (batch_smem(1, n_a, n_slow) > SMEM_LIMIT) return VREC_OK;
    for (int c = 0; c < n_chunks; ++c)
        chunk_ptr[c + 1] = chunk_ptr[c] + (chunk_r[c] <= VREC_CANON_SEG ? (long long)chunk_r[c] * 32 : 0);
    const long long sell_n = std::max<long long>(1, chunk_ptr[n_chunks]);
    const int r_nnz = r_rowptr[n_a];
    if (sell_n > 4LL * (r_nnz + 1024LL * n_chunks)) return VREC_OK;   // pathological padding: keep the per-query path
    std::vector<int> r_src((size_t)std::max(1, r_nnz));
    std::vector<double> r_w((size_t)std::max(1, r_nnz));
    std::vector<unsigned short> sell_src((size_t)sell_n, 0);
    std::vector<double> sell_w((size_t)sell_n, 0.0);
    for (int64_t i = 0; i < N; ++i) zcnt[i + 1] += zcnt[i];
    const int z_nnz = zcnt[N];
    std::vector<int> z_rowptr(zcnt), z_row((size_t)std::max(1, z_nnz)), z_pos((size_t)std::max(1, z_nnz));
    std::vector<double> z_w((size_t)std::max(1, z_nnz));
    std::vector<int> fill(zcnt.begin(), zcnt.end() - 1);
    for (int a = 0; a < n_a; ++a) {                            // ascending rows, ascending positions: lists end up sorted
        const int v = act_vertex[a], s = rowptr[v], e = rowptr[v + 1], m = m_of[v];
        const int n_max = chunk_r[a >> 5], R = (n_max + 31) / 32;
        for (int k = 0; k < m; ++k) {
            const int sa = b.h_act_of[h_src[s + k]];
            r_src[r_rowptr[a] + k] = sa;
            r_w[r_rowptr[a] + k] = h_w[s + k];
            if (n_max <= VREC_CANON_SEG) {
                const int l = k & 31, c2 = k >> 5;
                const long long step = c2 < R - 1 ? (long long)l * (R - 1) + c2 : 32LL * (R - 1) + l;
                const long long p = chunk_ptr[a >> 5] + step * 32 + (a & 31);
                sell_src[p] = (unsigned short)sa;
                sell_w[p] = h_w[s + k];
            }
        }
        for (int k = s + m; k < e; ++k) {
            const int p = fill[h_src[k]]++;
            z_row[p] = a;
            z_pos[p] = k - s;
            z_w[p] = h_w[k];
        }
    }
    cudaStream_t st = g->ctx->stream;
    VREC_TRY(b.r_rowptr.upload(r_rowptr.data(), r_rowptr.size(), st));
    VREC_TRY(b.r_src.upload(r_src.data(), r_src.size(), st));
    VREC_TRY(b.r_w.upload(r_w.data(), r_w.size(), st));
    VREC_TRY(b.full_len.upload(full_len.data(), full_len.size(), st));
    VREC_TRY(b.chunk_r.upload(chunk_r.data(), chunk_r.size(), st));
    VREC_TRY(b.chunk_ptr.upload(chunk_ptr.data(), chunk_ptr.size(), st));
    VREC_TRY(b.sell_src.upload(sell_src.data(), sell_src.size(), st));
    VREC_TRY(b.sell_w.upload(sell_w.data(), sell_w.size(), st));
    VREC_TRY(b.act_id.upload(act_id.data(), act_id.size(), st));
    VREC_TRY(b.z_rowptr.upload(z_rowptr.data(), z_rowptr.size(), st));
    VREC_TRY(b.z_row.upload(z_row.data(), z_row.size(), st));
    VREC_TRY(b.z_pos.upload(z_pos.data(), z_pos.size(), st));
    VREC_TRY(b.z_w.upload(z_w.data(), z_w.size(), st));
    VREC_TRY(b.x1a.alloc((size_t)n_a));
    VREC_CUDA(cudaStreamSynchronize(st));
    b.n_a = n_a;
    b.n_fast_chunks = n_fast_chunks;
    b.n_slow = n_slow;
    b.r_nnz = r_nnz;
    b.sell_nnz = chunk_ptr[n_chunks];
    b.x1_ready = false;
    b.ok = true;
    return VREC_OK;
}

// Iteration 0 without a start vertex (once per graph): x1 over the active vertices and its residual.
int sg_batch_prepare(vrec_sg *g) {
    SgBatch &b = g->batch;
    if (b.x1_ready) return VREC_OK;
    vrec_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int n_a = b.n_a;
    std::vector<int> act_vertex((size_t)n_a);
    for (int64_t v = 0; v < g->N; ++v)
        if (b.h_act_of[v] >= 0) act_vertex[b.h_act_of[v]] = (int)v;
    DevBuf<int> d_av;
    VREC_TRY(d_av.upload(act_vertex.data(), act_vertex.size(), st));
    VREC_TRY(sg_run_device(g, -1, 0.0, 1, true));
    SgState hs;
    int buf = 0;
    VREC_TRY(sg_fetch_state(g, 1, &hs, &buf));
    sg_gather_active_kernel<<<(n_a + 255) / 256, 256, 0, st>>>(g->d_x[1], d_av.p, n_a, b.x1a.p);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaStreamSynchronize(st));
    b.r1_base = hs.residual;
    b.x1_ready = true;
    return VREC_OK;
}

// Serves the queries qidx[] (positions in the caller's arrays; qvertex[] = their vertex indices, all
// without in-edges).  Requires max_it >= 1 and that iteration 0 did not converge (the caller checks).
int sg_batch_query(vrec_sg *g, const std::vector<int> &qidx, const std::vector<int> &qvertex,
                   double epsilon, int max_it, const int64_t *place_filter, int64_t n_filter, int max_recs,
                   int64_t *out_id, double *out_prob, int32_t *out_count, int32_t *out_iterations,
                   int32_t *out_converged) {
    SgBatch &b = g->batch;
    vrec_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int n_q = (int)qidx.size();
    const int n_a = b.n_a;
    const int n_chunks = (n_a + 31) / 32;
    VREC_TRY(sg_batch_prepare(g));
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
        return (long long)std::chrono::duration_cast<std::chrono::microseconds>(y - x).count();
    };
    const auto t_begin = now();
    // candidates of the ranked top-N: the filter's places that are active vertices (the others have
    // probability 0 after the first iteration and fail `probability > 0`, :85-88)
    std::vector<unsigned> mask((size_t)n_chunks, place_filter ? 0u : 0xffffffffu);
    if (place_filter) {
        for (int64_t k = 0; k < n_filter; ++k) {
            int64_t v = sg_lookup(g, place_filter[k]);
            if (v >= 0 && b.h_act_of[v] >= 0) mask[b.h_act_of[v] >> 5] |= 1u << (b.h_act_of[v] & 31);
        }
    }
    const int m = std::max(1, max_recs);
    VREC_TRY(b.cand_mask.upload(mask.data(), mask.size(), st));
    VREC_TRY(b.q_vertex.upload(qvertex.data(), qvertex.size(), st));
    VREC_TRY(b.out_id.ensure((size_t)n_q * m));
    VREC_TRY(b.out_prob.ensure((size_t)n_q * m));
    VREC_TRY(b.out_count.ensure((size_t)n_q));
    VREC_TRY(b.out_it.ensure((size_t)n_q));
    VREC_TRY(b.out_conv.ensure((size_t)n_q));
    VREC_TRY(b.counter.ensure(1));
    VREC_CUDA(cudaMemsetAsync(b.counter.p, 0, sizeof(int), st));
    // targets per CTA: as many as fit in shared memory, but keep every SM busy
    int T = 1;
    if (batch_smem(2, n_a, b.n_slow) <= SMEM_LIMIT && (n_q + 1) / 2 >= ctx->sm_count) T = 2;
    if ((b.force_t == 1 || b.force_t == 2) && batch_smem(b.force_t, n_a, b.n_slow) <= SMEM_LIMIT) T = b.force_t;
    const int n_groups = (n_q + T - 1) / T;
    const int grid = std::max(1, std::min(n_groups, ctx->sm_count));
    VREC_TRY(b.scratch.ensure((size_t)grid * T * n_a));
    SgBatchArgs a;
    a.n_a = n_a;
    a.n_chunks = n_chunks;
    a.n_fast_chunks = b.n_fast_chunks;
    a.n_slow = b.n_slow;
    a.chunk_r = b.chunk_r.p;
    a.chunk_ptr = b.chunk_ptr.p;
    a.sell_src = b.sell_src.p;
    a.sell_w = b.sell_w.p;
    a.r_rowptr = b.r_rowptr.p;
    a.r_src = b.r_src.p;
    a.full_len = b.full_len.p;
    a.r_w = b.r_w.p;
    a.z_rowptr = b.z_rowptr.p;
    a.z_row = b.z_row.p;
    a.z_pos = b.z_pos.p;
    a.z_w = b.z_w.p;
    a.x1a = b.x1a.p;
    a.act_id = b.act_id.p;
    a.q_vertex = b.q_vertex.p;
    a.n_q = n_q;
    a.max_it = max_it;
    a.eps2 = epsilon * epsilon;                                    // :40
    a.cand_mask = b.cand_mask.p;
    a.max_recs = max_recs;
    a.out_id = b.out_id.p;
    a.out_prob = b.out_prob.p;
    a.out_count = b.out_count.p;
    a.out_it = b.out_it.p;
    a.out_conv = b.out_conv.p;
    a.scratch = b.scratch.p;
    a.counter = b.counter.p;
    const size_t smem = batch_smem(T, n_a, b.n_slow);
    VREC_CUDA(cudaStreamSynchronize(st));
    const auto t_launch = now();
    if (T == 2) VREC_TRY(launch_batch<2>(ctx, a, grid, smem));
    else VREC_TRY(launch_batch<1>(ctx, a, grid, smem));
    VREC_CUDA(cudaStreamSynchronize(st));
    const auto t_kernel = now();
    std::vector<long long> h_id((size_t)n_q * m);
    std::vector<double> h_prob((size_t)n_q * m);
    std::vector<int> h_count((size_t)n_q), h_it((size_t)n_q), h_conv((size_t)n_q);
    VREC_CUDA(cudaMemcpyAsync(h_count.data(), b.out_count.p, sizeof(int) * n_q, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(h_it.data(), b.out_it.p, sizeof(int) * n_q, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(h_conv.data(), b.out_conv.p, sizeof(int) * n_q, cudaMemcpyDeviceToHost, st));
    if (max_recs > 0) {
        VREC_CUDA(cudaMemcpyAsync(h_id.data(), b.out_id.p, sizeof(long long) * (size_t)n_q * m, cudaMemcpyDeviceToHost, st));
        VREC_CUDA(cudaMemcpyAsync(h_prob.data(), b.out_prob.p, sizeof(double) * (size_t)n_q * m, cudaMemcpyDeviceToHost, st));
    }
    VREC_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n_q; ++i) {
        const int q = qidx[i];
        out_count[q] = h_count[i];
        if (out_iterations) out_iterations[q] = h_it[i];
        if (out_converged) out_converged[q] = h_conv[i];
        for (int k = 0; k < h_count[i]; ++k) {
            out_id[(size_t)q * max_recs + k] = h_id[(size_t)i * m + k];
            out_prob[(size_t)q * max_recs + k] = h_prob[(size_t)i * m + k];
        }
    }
    b.last_batched = n_q;
    b.us_prepare = us(t_begin, t_launch);
    b.us_kernel = us(t_launch, t_kernel);
    b.us_results = us(t_kernel, now());
    return VREC_OK;
}
