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
// reduced graph (12 B per edge, L2-resident) is streamed once per iteration for all T, rows are
// summed two per warp in the canonical order, and the step() loop (:92-106), the convergence test
// (:130-141) and the ranked top-N all run inside the kernel -- no launch per iteration or per query.
#include <algorithm>
#include <climits>

#include "vrec_sg.cuh"

namespace {

constexpr int BT = 1024;            // threads per CTA
constexpr int BW = BT / 32;
constexpr size_t SMEM_LIMIT = 227 * 1024 - 3 * 1024;   // dynamic part; statics + reserve stay below 3 KB

struct SgBatchArgs {
    int n_a, n_chunks;
    const int *r_rowptr, *r_src, *full_len;
    const double *r_w;
    const int *z_rowptr, *z_row, *z_pos;
    const double *z_w;
    const double *x1a;
    const int *q_vertex;
    int n_q, max_it;
    double eps2;
    const int *cand_act;
    const long long *cand_id;
    int n_cand, max_recs;
    long long *out_id;
    double *out_prob;
    int *out_count, *out_it, *out_conv;
    double *scratch;
    int *counter;
};

__device__ __forceinline__ int ld_nc_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Canonical sums of one reduced row (n <= VREC_CANON_SEG terms starting at edge s) for T x vectors
// in shared memory, by a group of G lanes; same lane rules as canon_row_sum in vrec_sg.cu.
template <int G, int T>
__device__ __forceinline__ void canon_row_sum_s(const int *__restrict__ src, const double *__restrict__ w,
                                                const double *xs, int n_a, int s, int n, int sublane,
                                                double (&out)[T]) {
    constexpr int V = 32 / G;
    constexpr int U = 4 / V > 0 ? 4 / V : 1;
    double acc[T][V];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[t][j] = 0.0;
    for (int kb = 0; kb < n; kb += 32 * U) {
        int c[U * V];
        double ww[U * V];
        bool ok[U * V];
#pragma unroll
        for (int q = 0; q < U * V; ++q) {
            int k = kb + sublane + G * q;
            ok[q] = k < n;
            c[q] = ok[q] ? ld_nc_i32(src + s + k) : 0;
            ww[q] = ok[q] ? ld_nc_f64(w + s + k) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < U * V; ++q)
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (ok[q]) acc[t][q % V] = xadd(acc[t][q % V], xmul(xs[t * n_a + c[q]], ww[q]));
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[t][j] = xadd(acc[t][j], __shfl_xor_sync(0xffffffffu, acc[t][j], off));
        }
#pragma unroll
        for (int off = 1; off < V; off <<= 1) {
            double tmp[V];
#pragma unroll
            for (int j = 0; j < V; ++j) tmp[j] = xadd(acc[t][j], acc[t][j ^ off]);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[t][j] = tmp[j];
        }
        out[t] = acc[t][0];
    }
}

// General row sum by a whole warp: the reduced prefix of `row` (m terms at positions 0..m-1 of the
// full row) plus the start vertex's own terms z[e..ee) (positions z_pos >= m, ascending, value
// 0.15 * w).  Rows whose FULL length exceeds VREC_CANON_SEG are summed per 1024-term segment and the
// segment sums by the same lane rule; all-zero segments are skipped.
__device__ double warp_row_sigma(const SgBatchArgs &a, const double *xs, int row, int e, int ee, int lane) {
    const int s = a.r_rowptr[row], m = a.r_rowptr[row + 1] - s;
    const int fl = a.full_len[row];
    if (fl <= VREC_CANON_SEG) {
        double acc = 0.0;
        for (int k = lane; k < m; k += 32) acc = xadd(acc, xmul(xs[a.r_src[s + k]], a.r_w[s + k]));
        for (int q = e; q < ee; ++q)
            if ((a.z_pos[q] & 31) == lane) acc = xadd(acc, xmul(kAlpha, a.z_w[q]));
        return canon_butterfly(acc);
    }
    const int nsp = (m + VREC_CANON_SEG - 1) / VREC_CANON_SEG;
    double acc2 = 0.0;
    int q = e;
    for (int j = 0; j < nsp; ++j) {
        const int len = min(VREC_CANON_SEG, m - j * VREC_CANON_SEG);
        const int s2 = s + j * VREC_CANON_SEG;
        double acc = 0.0;
        for (int k = lane; k < len; k += 32) acc = xadd(acc, xmul(xs[a.r_src[s2 + k]], a.r_w[s2 + k]));
        while (q < ee && a.z_pos[q] / VREC_CANON_SEG == j) {
            if ((a.z_pos[q] & 31) == lane) acc = xadd(acc, xmul(kAlpha, a.z_w[q]));
            ++q;
        }
        double part = canon_butterfly(acc);
        if ((j & 31) == lane) acc2 = xadd(acc2, part);
    }
    while (q < ee) {
        const int j = a.z_pos[q] / VREC_CANON_SEG;
        double acc = 0.0;
        while (q < ee && a.z_pos[q] / VREC_CANON_SEG == j) {
            if ((a.z_pos[q] & 31) == lane) acc = xadd(acc, xmul(kAlpha, a.z_w[q]));
            ++q;
        }
        double part = canon_butterfly(acc);
        if ((j & 31) == lane) acc2 = xadd(acc2, part);
    }
    return canon_butterfly(acc2);
}

template <int T>
__global__ void __launch_bounds__(BT, 1) sg_batch_kernel(const SgBatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_a = a.n_a, n_chunks = a.n_chunks;
    double *xs = reinterpret_cast<double *>(smem_raw);                      // [T][n_a]
    double *chunk_res = xs + (size_t)T * n_a;                               // [T][n_chunks]
    unsigned *flag = reinterpret_cast<unsigned *>(chunk_res + (size_t)T * n_chunks);   // [T][n_chunks] row bitmaps
    __shared__ int s_grp, s_next_chunk;
    __shared__ int s_done[T], s_iter[T], s_conv[T], s_copy[T];
    __shared__ double s_bv[BW];
    __shared__ long long s_bk[BW];
    __shared__ double s_sel_v;
    __shared__ long long s_sel_k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, sub = lane & 15;
    const double one_minus = 1 - kAlpha;                                    // :121
    double *nx = a.scratch + (size_t)blockIdx.x * T * n_a;                  // x' of this CTA's start vertices
    const int n_groups = (a.n_q + T - 1) / T;

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
            for (int t = 0; t < T; ++t) xs[t * n_a + i] = v;
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
                    double sigma = warp_row_sigma(a, xs + t * n_a, row, e, ee, lane);
                    if (lane == 0) {
                        nx[t * n_a + row] = xadd(xmul(0.0, kAlpha), xmul(sigma, one_minus));   // :120-122
                        atomicOr(&flag[t * n_chunks + (row >> 5)], 1u << (row & 31));
                    }
                }
            }
            __syncthreads();
            // every other row: calcNextX (:108-128) over the reduced graph, 32 rows per warp-chunk
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_next_chunk, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= n_chunks) break;
                const int r = c * 32 + lane;
                int s = 0, n = 0;
                if (r < n_a) {
                    s = a.r_rowptr[r];
                    n = a.r_rowptr[r + 1] - s;
                }
                double sigma[T];
#pragma unroll
                for (int t = 0; t < T; ++t) sigma[t] = 0.0;
                const unsigned shortm = __ballot_sync(0xffffffffu, n > 0 && n <= VREC_CANON_SEG);
                unsigned longm = __ballot_sync(0xffffffffu, n > VREC_CANON_SEG);
                unsigned pairs = (shortm | (shortm >> 1)) & 0x55555555u;
                while (pairs) {
                    const int l0 = __ffs(pairs) - 1;
                    pairs &= pairs - 1;
                    const int l = l0 + half;
                    const int rs = __shfl_sync(0xffffffffu, s, l);
                    int rn = __shfl_sync(0xffffffffu, n, l);
                    if (rn > VREC_CANON_SEG) rn = 0;
                    double acc[T];
                    canon_row_sum_s<16, T>(a.r_src, a.r_w, xs, n_a, rs, rn, sub, acc);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        double other = __shfl_xor_sync(0xffffffffu, acc[t], 16);
                        if (lane == l0) sigma[t] = half == 0 ? acc[t] : other;
                        if (lane == l0 + 1) sigma[t] = half == 1 ? acc[t] : other;
                    }
                }
                while (longm) {
                    const int l = __ffs(longm) - 1;
                    longm &= longm - 1;
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        double v = warp_row_sigma(a, xs + t * n_a, c * 32 + l, 0, 0, lane);
                        if (lane == l) sigma[t] = v;
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
                        double d = xsub(v, xs[t * n_a + r]);                // isConverged, :131-139
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
                for (int k = lane; k < n_chunks; k += 32) acc = xadd(acc, chunk_res[t * n_chunks + k]);
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
                for (int i = tid; i < n_a; i += BT) xs[t * n_a + i] = nx[t * n_a + i];
            }
            __syncthreads();
            bool all = true;
#pragma unroll
            for (int t = 0; t < T; ++t) all = all && s_done[t];
            if (all) break;
        }
        // makeRecommendations0's filter (:85-88) + printRecommendations' ranked top-N (Main :69-73)
        for (int t = 0; t < nt; ++t) {
            const int q = grp * T + t;
            double pv = __longlong_as_double(0x7ff0000000000000LL);
            long long pk = LLONG_MIN;
            int count = 0;
            for (int r = 0; r < a.max_recs; ++r) {
                double bv = -1.0;
                long long bk = LLONG_MAX;
                for (int i = tid; i < a.n_cand; i += BT) {
                    double v = xs[t * n_a + a.cand_act[i]];
                    long long k = a.cand_id[i];
                    if (v > 0 && ranks_before(pv, pk, v, k) && ranks_before(v, k, bv, bk)) {
                        bv = v;
                        bk = k;
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    long long ok = __shfl_xor_sync(0xffffffffu, bk, off);
                    if (ranks_before(ov, ok, bv, bk)) {
                        bv = ov;
                        bk = ok;
                    }
                }
                if (lane == 0) {
                    s_bv[warp] = bv;
                    s_bk[warp] = bk;
                }
                __syncthreads();
                if (warp == 0) {
                    bv = s_bv[lane];
                    bk = s_bk[lane];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                        long long ok = __shfl_xor_sync(0xffffffffu, bk, off);
                        if (ranks_before(ov, ok, bv, bk)) {
                            bv = ov;
                            bk = ok;
                        }
                    }
                    if (lane == 0) {
                        s_sel_v = bv;
                        s_sel_k = bk;
                    }
                }
                __syncthreads();
                pv = s_sel_v;
                pk = s_sel_k;
                if (!(pv > 0)) break;
                if (tid == 0) {
                    a.out_id[(size_t)q * a.max_recs + r] = pk;
                    a.out_prob[(size_t)q * a.max_recs + r] = pv;
                }
                ++count;
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

size_t batch_smem(int T, int n_a) {
    size_t n_chunks = (size_t)(n_a + 31) / 32;
    return (size_t)T * n_a * 8 + (size_t)T * n_chunks * 8 + (size_t)T * n_chunks * 4 + 16;
}

template <int T>
int launch_batch(vrec_ctx *ctx, const SgBatchArgs &a, int grid, size_t smem) {
    VREC_CUDA(cudaFuncSetAttribute(sg_batch_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sg_batch_kernel<T><<<grid, BT, smem, ctx->stream>>>(a);
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}

}  // namespace

// Load-time analysis of a host-built graph: active set, reduced CSR, out-edges of the in-degree-0
// vertices with their positions in the full rows.  Leaves batch.ok = false when the graph does not
// have the required shape.
int sg_batch_analyse(vrec_sg *g, const std::vector<int> &rowptr, const int *h_src, const double *h_w) {
    SgBatch &b = g->batch;
    b.analysed = true;
    b.ok = false;
    const int64_t N = g->N;
    if (N <= 0 || g->partitioned || g->nblocks != 1 || g->row_lo != 0 || g->row_hi != N) return VREC_OK;
    b.h_act_of.assign((size_t)N, -1);
    std::vector<int> act_vertex;
    for (int64_t v = 0; v < N; ++v)
        if (rowptr[v + 1] > rowptr[v]) {
            b.h_act_of[v] = (int)act_vertex.size();
            act_vertex.push_back((int)v);
        }
    const int n_a = (int)act_vertex.size();
    if (n_a == 0 || n_a == N || batch_smem(1, n_a) > SMEM_LIMIT) return VREC_OK;
    std::vector<int> r_rowptr((size_t)n_a + 1, 0), full_len((size_t)n_a), zcnt((size_t)N + 1, 0);
    for (int a = 0; a < n_a; ++a) {
        const int v = act_vertex[a], s = rowptr[v], e = rowptr[v + 1];
        int m = 0;
        while (s + m < e && b.h_act_of[h_src[s + m]] >= 0) ++m;
        for (int k = s + m; k < e; ++k) {
            if (b.h_act_of[h_src[k]] >= 0) return VREC_OK;     // an active source after a Z source: not this shape
            zcnt[h_src[k] + 1]++;
        }
        r_rowptr[a + 1] = r_rowptr[a] + m;
        full_len[a] = e - s;
    }
    const int r_nnz = r_rowptr[n_a];
    std::vector<int> r_src((size_t)std::max(1, r_nnz));
    std::vector<double> r_w((size_t)std::max(1, r_nnz));
    for (int64_t i = 0; i < N; ++i) zcnt[i + 1] += zcnt[i];
    const int z_nnz = zcnt[N];
    std::vector<int> z_rowptr(zcnt), z_row((size_t)std::max(1, z_nnz)), z_pos((size_t)std::max(1, z_nnz));
    std::vector<double> z_w((size_t)std::max(1, z_nnz));
    std::vector<int> fill(zcnt.begin(), zcnt.end() - 1);
    for (int a = 0; a < n_a; ++a) {                            // ascending rows, ascending positions: lists end up sorted
        const int v = act_vertex[a], s = rowptr[v], e = rowptr[v + 1];
        const int m = r_rowptr[a + 1] - r_rowptr[a];
        for (int k = 0; k < m; ++k) {
            r_src[r_rowptr[a] + k] = b.h_act_of[h_src[s + k]];
            r_w[r_rowptr[a] + k] = h_w[s + k];
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
    VREC_TRY(b.z_rowptr.upload(z_rowptr.data(), z_rowptr.size(), st));
    VREC_TRY(b.z_row.upload(z_row.data(), z_row.size(), st));
    VREC_TRY(b.z_pos.upload(z_pos.data(), z_pos.size(), st));
    VREC_TRY(b.z_w.upload(z_w.data(), z_w.size(), st));
    VREC_TRY(b.x1a.alloc((size_t)n_a));
    VREC_CUDA(cudaStreamSynchronize(st));
    b.n_a = n_a;
    b.r_nnz = r_nnz;
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
    sg_gather_active_kernel<<<(n_a + 255) / 256, 256, 0, st>>>(g->d_x[1].p, d_av.p, n_a, b.x1a.p);
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
    VREC_TRY(sg_batch_prepare(g));
    // candidates of the ranked top-N: the filter's places that are active vertices (the others have
    // probability 0 after the first iteration and fail `probability > 0`, :85-88)
    std::vector<int> cand_act;
    std::vector<long long> cand_id;
    if (place_filter) {
        for (int64_t k = 0; k < n_filter; ++k) {
            int64_t v = sg_lookup(g, place_filter[k]);
            if (v >= 0 && b.h_act_of[v] >= 0) {
                cand_act.push_back(b.h_act_of[v]);
                cand_id.push_back((long long)place_filter[k]);
            }
        }
    } else {
        for (int64_t v = 0; v < g->N; ++v)
            if (b.h_act_of[v] >= 0) {
                cand_act.push_back(b.h_act_of[v]);
                cand_id.push_back((long long)g->h_ids[v]);
            }
    }
    const int n_cand = (int)cand_act.size();
    const int m = std::max(1, max_recs);
    VREC_TRY(b.cand_act.upload(cand_act.data(), cand_act.size(), st));
    VREC_TRY(b.cand_id.upload(cand_id.data(), cand_id.size(), st));
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
    for (int cand : {2, 4})
        if (batch_smem(cand, n_a) <= SMEM_LIMIT && (n_q + cand - 1) / cand >= ctx->sm_count) T = cand;
    if (b.force_t == 1 || b.force_t == 2 || b.force_t == 4) {
        if (batch_smem(b.force_t, n_a) <= SMEM_LIMIT) T = b.force_t;
    }
    const int n_groups = (n_q + T - 1) / T;
    const int grid = std::max(1, std::min(n_groups, ctx->sm_count));
    VREC_TRY(b.scratch.ensure((size_t)grid * T * n_a));
    SgBatchArgs a;
    a.n_a = n_a;
    a.n_chunks = (n_a + 31) / 32;
    a.r_rowptr = b.r_rowptr.p;
    a.r_src = b.r_src.p;
    a.full_len = b.full_len.p;
    a.r_w = b.r_w.p;
    a.z_rowptr = b.z_rowptr.p;
    a.z_row = b.z_row.p;
    a.z_pos = b.z_pos.p;
    a.z_w = b.z_w.p;
    a.x1a = b.x1a.p;
    a.q_vertex = b.q_vertex.p;
    a.n_q = n_q;
    a.max_it = max_it;
    a.eps2 = epsilon * epsilon;                                    // :40
    a.cand_act = b.cand_act.p;
    a.cand_id = b.cand_id.p;
    a.n_cand = n_cand;
    a.max_recs = max_recs;
    a.out_id = b.out_id.p;
    a.out_prob = b.out_prob.p;
    a.out_count = b.out_count.p;
    a.out_it = b.out_it.p;
    a.out_conv = b.out_conv.p;
    a.scratch = b.scratch.p;
    a.counter = b.counter.p;
    const size_t smem = batch_smem(T, n_a);
    if (T == 4) VREC_TRY(launch_batch<4>(ctx, a, grid, smem));
    else if (T == 2) VREC_TRY(launch_batch<2>(ctx, a, grid, smem));
    else VREC_TRY(launch_batch<1>(ctx, a, grid, smem));
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
    return VREC_OK;
}
