// SG path: damped power iteration x' = alpha*u + (1-alpha) * P^T x over the CSR of P^T with a
// fused residual / convergence check, then ranked top-N place extraction.
// Reference: stochastic/StochasticRecommender.scala:38-141, stochastic/StochasticRecommenderMain.scala:64-76.
#include <algorithm>
#include <memory>
#include <numeric>

#include <string.h>

#include "vrec_sg.cuh"

static int sg_check_params(double epsilon, int32_t max_iterations);

namespace {

__global__ void sg_fill_kernel(double *x, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] = v;
}

__global__ void sg_reset_state_kernel(SgState *st, int max_it, long long uidx, double eps2, int check,
                                      unsigned long long flag_base) {
    st->done = max_it <= 0 ? 1 : 0;          // maxIterations == 0 returns x0 (:93-95)
    st->iterations = 0;
    st->converged = 0;
    st->ticket = 0;
    st->residual = -1.0;
    st->cur = 0;
    st->max_it = max_it;
    st->check = check;
    st->uidx = uidx;
    st->eps2 = eps2;
    st->flag_base = flag_base;
}

// Canonical sum of terms x[src[k]] * w[k], k in [s, s+n), n <= VREC_CANON_SEG, computed by a
// group of G physical lanes (G = 32 or 16).  Term k belongs to canonical lane k % 32; physical
// lane p = k % G keeps the 32/G canonical lanes p, p+G, ... in separate accumulators, so the
// per-lane order and the xor-butterfly 1,2,4,8,16 are exactly those of the oracle's warp_sum.
// With G = 16 a warp sums two rows at once: twice the gathers in flight per warp.
template <int G>
__device__ __forceinline__ double canon_row_sum(const int *__restrict__ src, const double *__restrict__ w,
                                                const double *__restrict__ x, int s, int n, int sublane,
                                                unsigned long long pol_stream, unsigned long long pol_keep) {
    constexpr int V = 32 / G;           // canonical lanes per physical lane
    constexpr int U = 4 / V > 0 ? 4 / V : 1;   // unroll so that 4 gathers are in flight per lane
    double acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0;
    for (int kb = 0; kb < n; kb += 32 * U) {    // uniform trip count within the group
        int c[U * V];
        double ww[U * V], xx[U * V];
        bool ok[U * V];
#pragma unroll
        for (int q = 0; q < U * V; ++q) {
            int k = kb + sublane + G * q;
            ok[q] = k < n;
            c[q] = ok[q] ? ld_stream_i32(src + s + k, pol_stream) : 0;
            ww[q] = ok[q] ? ld_stream_f64(w + s + k, pol_stream) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < U * V; ++q) xx[q] = ld_keep_f64(x + c[q], pol_keep);
#pragma unroll
        for (int q = 0; q < U * V; ++q)
            if (ok[q]) acc[q % V] = xadd(acc[q % V], xmul(xx[q], ww[q]));
    }
    // butterfly over canonical lanes: offsets < G are shuffles, offsets >= G stay in the thread
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = xadd(acc[j], __shfl_xor_sync(0xffffffffu, acc[j], off));
    }
#pragma unroll
    for (int off = 1; off < V; off <<= 1) {
        double t[V];
#pragma unroll
        for (int j = 0; j < V; ++j) t[j] = xadd(acc[j], acc[j ^ off]);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = t[j];
    }
    return acc[0];
}


// ---------------------------------------------------------------------------------------
// Flat-window summation (round 2).  A warp owns consecutive rows whose in-edges are one contiguous
// range [S, E) of src / w.  It walks that range in 128-byte aligned windows of 32 edges -- lane l
// always holds position p = l (mod 32), so every src load of the warp is one 128-byte line and
// every w load two: the L1TEX pipe, which retires one wavefront per clock and is what bounds
// this kernel (one wavefront per scattered gather of x, tools/gather_probe.cu), carries no
// wasted streaming wavefronts.  Term k of a row that starts at rs sits at position rs + k,
// i.e. in physical lane (rs + k) mod 32: canonical lane c = k mod 32 is physical lane
// (c + rs) mod 32, a rotation.  Each physical lane adds its terms in ascending position
// (= ascending k, step 32), which is exactly the canonical per-lane order, and when a row ends
// the 32 lane sums are combined by the canonical xor-butterfly over the *rotated* lane numbers.
// Rows may end anywhere inside a window; the loads and gathers of the next windows are
// already in flight while a row is being closed (no per-row drain of the memory pipeline).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double rotated_butterfly(double v, int rot, int lane) {
    const int c = (lane - rot) & 31;                    // canonical lane of this physical lane
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int srcl = ((c ^ off) + rot) & 31;
        v = xadd(v, __shfl_sync(0xffffffffu, v, srcl));
    }
    return v;                                           // fp add commutes: every lane holds lane 0's sum
}


struct FlatRows {                                       // the rows still open in a run (uniform across the warp)
    unsigned rows;
    int cur, rs_cur, re_cur;
    double acc, sigma;
};

// One batch (FLAT_U windows from wb) of gathered values and weights: adds every lane's products to
// its open row and closes the rows that end inside a window.
template <int FLAT_U>
__device__ __forceinline__ void flat_process(FlatRows &f, const double (&xx)[FLAT_U], const double (&pw)[FLAT_U],
                                             int wb, int S, int E, int e, int lane) {
#pragma unroll
    for (int q = 0; q < FLAT_U; ++q) {
        const int w0 = wb + 32 * q;
        if (w0 < E) {                                   // uniform
            const int P = w0 + lane, wend = w0 + 32;
            const bool ok = P >= S && P < E;
            const double prod = xmul(xx[q], pw[q]);
            while (f.re_cur <= wend) {                  // uniform: row `cur` ends inside this window
                const bool mine = ok && P >= f.rs_cur && P < f.re_cur;
                double v = mine ? xadd(f.acc, prod) : f.acc;
                v = rotated_butterfly(v, f.rs_cur & 31, lane);
                if (lane == f.cur) f.sigma = v;
                f.acc = 0.0;
                f.rs_cur = f.re_cur;                    // rows tile [S, E): the next non-empty row starts here
                if (f.rows) {
                    f.cur = __ffs(f.rows) - 1;
                    f.rows &= f.rows - 1;
                    f.re_cur = __shfl_sync(0xffffffffu, e, f.cur);
                } else {
                    f.cur = -1;
                    f.rs_cur = f.re_cur = 0x7fffffff;
                }
            }
            if (ok && P >= f.rs_cur) f.acc = xadd(f.acc, prod);   // the row that goes on past this window
        }
    }
}

template <int FLAT_U>
__device__ __forceinline__ void flat_load_src(int (&c)[FLAT_U], const int *__restrict__ src, int wb, int S, int E,
                                              int lane, unsigned long long pol) {
#pragma unroll
    for (int q = 0; q < FLAT_U; ++q) {
        const int P = wb + 32 * q + lane;
        c[q] = (P >= S && P < E) ? ld_stream_i32(src + P, pol) : -1;
    }
}
// gathers of x and the weights of the same batch (both are first needed when the batch is processed)
template <int FLAT_U>
__device__ __forceinline__ void flat_gather(double (&xx)[FLAT_U], double (&pw)[FLAT_U], const int (&c)[FLAT_U],
                                            const double *__restrict__ x, const double *__restrict__ w, int wb,
                                            int lane, unsigned long long pol_stream, unsigned long long pol_keep) {
#pragma unroll
    for (int q = 0; q < FLAT_U; ++q) {
        const bool ok = c[q] >= 0;
        xx[q] = ok ? ld_keep_f64(x + c[q], pol_keep) : 0.0;
        pw[q] = ok ? ld_stream_f64(w + wb + 32 * q + lane, pol_stream) : 0.0;
    }
}

// Sums the non-empty rows `rows` (bit i = row of lane i; e = end of that lane's row; rows ascending and
// contiguous: together they tile [S, E)) and returns this lane's row sum (0.0 for the other lanes).
// Software pipeline, three batches deep: the columns of batch n+2 and the gathers / weights of batch n+1
// are in flight while batch n is processed.
template <int FLAT_U>
__device__ __forceinline__ double flat_rows_sum(const int *__restrict__ src, const double *__restrict__ w,
                                                const double *__restrict__ x, int e, unsigned rows,
                                                int S, int E, int lane, unsigned long long pol_stream,
                                                unsigned long long pol_keep) {
    FlatRows f;
    f.sigma = 0.0;
    if (rows == 0u || S >= E) return f.sigma;
    f.cur = __ffs(rows) - 1;
    f.rows = rows & (rows - 1);
    f.rs_cur = S;
    f.re_cur = __shfl_sync(0xffffffffu, e, f.cur);
    f.acc = 0.0;
    constexpr int B = 32 * FLAT_U;
    int cA[FLAT_U], cB[FLAT_U];
    double xA[FLAT_U], wA[FLAT_U], xB[FLAT_U], wB[FLAT_U];
    int wb = S & ~31;
    flat_load_src<FLAT_U>(cA, src, wb, S, E, lane, pol_stream);
    flat_load_src<FLAT_U>(cB, src, wb + B, S, E, lane, pol_stream);
    flat_gather<FLAT_U>(xA, wA, cA, x, w, wb, lane, pol_stream, pol_keep);
    flat_load_src<FLAT_U>(cA, src, wb + 2 * B, S, E, lane, pol_stream);
    for (; wb < E; wb += 2 * B) {
        // batch n in (xA, wA); columns of n+1 in cB, of n+2 in cA
        flat_gather<FLAT_U>(xB, wB, cB, x, w, wb + B, lane, pol_stream, pol_keep);
        flat_load_src<FLAT_U>(cB, src, wb + 3 * B, S, E, lane, pol_stream);
        flat_process<FLAT_U>(f, xA, wA, wb, S, E, e, lane);
        if (wb + B >= E) break;                         // uniform
        flat_gather<FLAT_U>(xA, wA, cA, x, w, wb + 2 * B, lane, pol_stream, pol_keep);
        flat_load_src<FLAT_U>(cA, src, wb + 4 * B, S, E, lane, pol_stream);
        flat_process<FLAT_U>(f, xB, wB, wb + B, S, E, e, lane);
    }
    return f.sigma;
}

// One warp per 1024-term segment of a long row.
__global__ void __launch_bounds__(SPMV_THREADS)
sg_long_partials_kernel(const int *__restrict__ row_start, const int *__restrict__ row_end,
                        const int *__restrict__ src,
                        const double *__restrict__ w, const double *x0, const double *x1,
                        const int *__restrict__ long_rows, const int *__restrict__ long_segptr,
                        const int *__restrict__ seg_row, int n_seg, double *__restrict__ partials,
                        const SgState *st, float keep_frac) {
    if (st->done) return;
    const double *__restrict__ x = (st->cur & 1) ? x1 : x0;
    const int lane = threadIdx.x & 31;
    int seg = blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    int slot = seg_row[seg];
    int row = long_rows[slot];
    int j = seg - long_segptr[slot];
    int s = row_start[row] + j * VREC_CANON_SEG;
    int n = min(VREC_CANON_SEG, row_end[row] - s);
    // one row [s, s+n) held by lane 0
    double v = flat_rows_sum<4>(src, w, x, s + n, 1u, s, s + n, lane, policy_evict_first(), policy_evict_last(keep_frac));
    if (lane == 0) partials[seg] = v;
}

// step() control, stochastic/StochasticRecommender.scala:92-106,130-141
__device__ __forceinline__ void sg_step_decide(SgState *st, double acc, int iteration) {
    const int max_it = st->max_it;
    if (st->check) st->residual = acc;
    if (st->check && acc <= st->eps2) {  // :140 `diffSquared <= epsilonSquared`
        st->converged = 1;
        st->iterations = iteration;      // :100 "Converged in $iteration iterations"
        st->done = 1;
    } else if (iteration + 1 >= max_it) { // :93-95
        st->converged = 0;
        st->iterations = max_it;
        st->done = 1;
    }
    st->cur = iteration + 1;             // every CTA of this sweep has read `cur` (it took its ticket at its end)
}

// Main pass: sigma per row in the canonical order, x' = u*alpha + sigma*(1-alpha)
// (calcNextX, :108-128), squared-difference residual (isConverged, :130-141) reduced in a
// fixed order, and the step() control (:92-106) updated by the last block.
template <bool flat, int MINB, int U>
__global__ void __launch_bounds__(SPMV_THREADS, MINB)
sg_spmv_kernel(int n_rows, long long row_lo, const int *__restrict__ row_start, const int *__restrict__ row_end,
               const int *__restrict__ src,
               const double *__restrict__ w, double *x0, double *x1,
               const int *__restrict__ long_rows, const int *__restrict__ long_segptr,
               int n_long, const double *__restrict__ partials, SgState *st,
               double *__restrict__ block_partials, float keep_frac, int acc_in, int finalize,
               const SgPeers *__restrict__ peers, cudaGraphConditionalHandle loop_cond) {
    if (st->done) {
        // replayed by the CUDA-graph `while` node (loop_cond != 0): make sure the loop ends
        if (loop_cond && finalize && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(loop_cond, 0u);
        return;
    }
    // the sweep's parameters (device-resident loop state): iteration `cur` reads x[cur & 1] and writes the other
    const int iteration = st->cur, check_convergence = st->check;
    const long long uidx = st->uidx;
    const double *__restrict__ x = (iteration & 1) ? x1 : x0;
    double *__restrict__ nx = (iteration & 1) ? x0 : x1;
    const int out_buf = (iteration + 1) & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total_warps = (long long)gridDim.x * SPMV_WARPS;
    const double one_minus = 1 - kAlpha;                     // :121, evaluates to 0.85
    const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last(keep_frac);
    double dsum = 0.0;
    for (long long base = ((long long)blockIdx.x * SPMV_WARPS + warp) * 32; base < n_rows;
         base += total_warps * 32) {
        long long r = base + lane;
        int s, e;
        if (r < n_rows) {
            s = row_start[r];
            e = row_end[r];
        } else {
            s = e = row_end[n_rows - 1];                 // keeps the chunk's flat range well-formed
        }
        int n = e - s;
        double sigma = 0.0;
        const unsigned nonempty = __ballot_sync(0xffffffffu, n > 0 && n <= VREC_CANON_SEG);
        unsigned longrows = __ballot_sync(0xffffffffu, n > VREC_CANON_SEG);
        if (flat) {
            // rows with contiguous in-edges: one flat sweep per run of rows between two long rows
            unsigned lm = longrows;
            int i0 = 0;
            while (i0 < 32) {
                const int j = lm ? __ffs(lm) - 1 : 32;
                if (j > i0) {
                    const unsigned run = (j == 32 ? 0xffffffffu : ((1u << j) - 1u)) & ~((1u << i0) - 1u);
                    const int S = __shfl_sync(0xffffffffu, s, i0), E = __shfl_sync(0xffffffffu, e, j - 1);
                    const double v = flat_rows_sum<U>(src, w, x, e, nonempty & run, S, E, lane, pol_stream, pol_keep);
                    if ((run >> lane) & 1u) sigma = v;
                }
                if (j == 32) break;
                lm &= lm - 1;
                i0 = j + 1;
            }
        }
        // (old layout: sub-ranges of a row are not adjacent) two rows at a time: half-warp h sums row 2j+h
        unsigned pairs = flat ? 0u : (nonempty | (nonempty >> 1)) & 0x55555555u;
        const int half = lane >> 4, sub = lane & 15;
        while (pairs) {
            int l0 = __ffs(pairs) - 1;
            pairs &= pairs - 1;
            int l = l0 + half;
            int rs = __shfl_sync(0xffffffffu, s, l);
            int rn = __shfl_sync(0xffffffffu, n, l);
            if (rn > VREC_CANON_SEG) rn = 0;               // long rows are handled below
            double acc = canon_row_sum<16>(src, w, x, rs, rn, sub, pol_stream, pol_keep);
            double other = __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane == l0) sigma = half == 0 ? acc : other;
            if (lane == l0 + 1) sigma = half == 1 ? acc : other;
        }
        while (longrows) {
            // second level: lane-strided sum of the segment partials of this row
            int l = __ffs(longrows) - 1;
            longrows &= longrows - 1;
            int row = (int)(base + l);
            int lo = 0, hi = n_long;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (long_rows[mid] < row) lo = mid + 1; else hi = mid;
            }
            int ps = long_segptr[lo], pn = long_segptr[lo + 1] - ps;
            double acc = 0.0;
            for (int k = lane; k < pn; k += 32) acc = xadd(acc, partials[ps + k]);
            acc = canon_butterfly(acc);
            if (lane == l) sigma = acc;
        }
        if (r < n_rows && !finalize) {
            // not the last source block: keep the running sum of the block sums in nx
            long long gi = row_lo + r;
            nx[gi] = acc_in ? xadd(nx[gi], sigma) : sigma;
        }
        if (r < n_rows && finalize) {
            long long gi = row_lo + r;
            if (acc_in) sigma = xadd(nx[gi], sigma);     // block sums are added left to right
            double u = (gi == uidx) ? 1.0 : 0.0;
            double v = xadd(xmul(u, kAlpha), xmul(sigma, one_minus));
            st_stream_f64(nx + gi, v, pol_stream);         // the written buffer is the old x: demote it
            if (peers) {
                // row-partitioned graph: the exchange step of the path, fused into the sweep -- every rank
                // needs the whole x' for its gathers, so the owner stores its rows straight into the peers'
                // buffers over NVLink (32 consecutive rows per warp = 256-byte peer stores)
                const int world = peers->world, me = peers->rank;
                for (int p = 0; p < world; ++p)
                    if (p != me) peers->x[out_buf][p][gi] = v;
            }
            double d = xsub(v, ld_keep_f64(x + gi, pol_keep));
            dsum = xadd(dsum, xmul(d, d));
        }
    }
    if (!finalize) return;
    // fixed-order reduction of the residual; the last CTA also advances the loop state
    __shared__ double s_part[SPMV_WARPS];
    __shared__ int s_last;
    dsum = canon_butterfly(dsum);
    if (lane == 0) s_part[warp] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < SPMV_WARPS; ++i) t = xadd(t, s_part[i]);
        block_partials[blockIdx.x] = t;
    }
    // this CTA's stores (local and peer) are ordered before its ticket
    if (peers) __threadfence_system(); else __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tk = atomicInc(&st->ticket, gridDim.x - 1);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        double acc = 0.0;
        for (int k = lane; k < (int)gridDim.x; k += 32) acc = xadd(acc, __ldcg(block_partials + k));
        acc = canon_butterfly(acc);
        if (peers) {
            // every CTA's ticket has been seen: publish this rank's residual partial and then its step flag to
            // every rank (itself included); sg_exchange_wait_kernel takes the step() decision on each rank
            const int world = peers->world, me = peers->rank;
            if (lane < world) peers->xchg[lane]->res[iteration & 1][me] = acc;
            __threadfence_system();
            __syncwarp();
            if (lane < world) {
                volatile unsigned long long *f = &peers->xchg[lane]->flag[me];
                *f = st->flag_base + (unsigned long long)(iteration + 1);
            }
        } else if (lane == 0) {
            sg_step_decide(st, acc, iteration);
            // body of the CUDA-graph `while` node: go on while step() has not returned
            if (loop_cond) cudaGraphSetConditional(loop_cond, st->done ? 0u : 1u);
        }
    }
}

// Row-partitioned graphs.  Waits until every rank has published `flag_value` (its x' rows and residual
// partial have then landed in this rank's memory), adds the partials in rank order -- the same sum on
// every rank -- and takes the step() decision.  `spin` = 0 when the ranks are driven from one stream
// (single-process group): stream order already guarantees the flags.
__global__ void sg_exchange_wait_kernel(SgState *st, SgExchange *xchg, int world, int step, int decide, int spin) {
    if (st->done) return;
    const int lane = threadIdx.x;
    const int iteration = st->cur;
    const unsigned long long flag_value = st->flag_base + (unsigned long long)step;
    int late = 0;
    if (spin && lane < world) {
        volatile unsigned long long *f = &xchg->flag[lane];
        const long long t0 = clock64();
        while (*f < flag_value) {
            if (clock64() - t0 > (1LL << 33)) {     // ~4 s: a peer is gone
                late = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    late = __any_sync(0xffffffffu, late);
    __threadfence_system();
    if (lane == 0) {
        if (late) {
            st->converged = -1;                     // reported as VREC_ENCCL by the host
            st->done = 1;
        } else if (decide) {
            volatile double *r = xchg->res[iteration & 1];
            double acc = 0.0;
            for (int q = 0; q < world; ++q) acc = xadd(acc, r[q]);
            sg_step_decide(st, acc, iteration);
        }
    }
}

// publishes a step flag without a sweep (start-of-query barrier: nobody may write a peer's buffers while that
// peer still reads the result of its previous query)
__global__ void sg_exchange_signal_kernel(const SgPeers *__restrict__ peers, unsigned long long flag_value) {
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < peers->world) {
        volatile unsigned long long *f = &peers->xchg[lane]->flag[peers->rank];
        *f = flag_value;
    }
}

// candidate values for the ranked top-N: filter ids -> vertex index by binary search
__global__ void sg_candidates_kernel(const long long *__restrict__ ids, long long N,
                                     const long long *__restrict__ filter, long long n_filter,
                                     const double *x0, const double *x1, const SgState *st, long long target_id,
                                     double *__restrict__ cand_val, long long *__restrict__ cand_key) {
    // which buffer step() returned: nextX when it converged (:101), x when the limit was reached (:95)
    const int buf = st->max_it <= 0 ? 0 : st->converged ? ((st->iterations + 1) & 1) : (st->max_it & 1);
    const double *__restrict__ x = buf ? x1 : x0;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = filter ? n_filter : N;
    if (i >= n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    long long id;
    long long idx;
    if (filter) {
        id = filter[i];
        if (ids) {
            long long lo = 0, hi = N;
            while (lo < hi) {
                long long mid = (lo + hi) >> 1;
                if (ids[mid] < id) lo = mid + 1; else hi = mid;
            }
            idx = (lo < N && ids[lo] == id) ? lo : -1;
        } else {
            idx = (id >= 0 && id < N) ? id : -1;
        }
    } else {
        idx = i;
        id = ids ? ids[i] : i;                  // generated graphs: ids are 0..N-1
    }
    double v = nan;
    if (idx >= 0 && id != target_id) {
        double p = x[idx];
        if (p > 0) v = p;                       // :85-88 id != vertex and probability > 0
    }
    cand_val[i] = v;
    cand_key[i] = id;
}

// builds the long-(sub-)range tables of one block from a host copy of its rowptr
int sg_fill_block(vrec_sg *g, vrec_sg::Block &blk, const int *h_rowptr, int64_t rows) {
    std::vector<int> long_rows, long_segptr(1, 0), seg_row;
    for (int64_t r = 0; r < rows; ++r) {
        int n = h_rowptr[r + 1] - h_rowptr[r];
        if (n > VREC_CANON_SEG) {
            int m = (n + VREC_CANON_SEG - 1) / VREC_CANON_SEG;
            for (int j = 0; j < m; ++j) seg_row.push_back((int)long_rows.size());
            long_rows.push_back((int)r);
            long_segptr.push_back(long_segptr.back() + m);
        }
    }
    blk.n_long = (int)long_rows.size();
    blk.n_seg = (int)seg_row.size();
    cudaStream_t st = g->ctx->stream;
    VREC_TRY(blk.long_rows.upload(long_rows.data(), long_rows.size(), st));
    VREC_TRY(blk.long_segptr.upload(long_segptr.data(), long_segptr.size(), st));
    VREC_TRY(blk.seg_row.upload(seg_row.data(), seg_row.size(), st));
    VREC_TRY(blk.partials.alloc(std::max(1, blk.n_seg)));
    return VREC_OK;
}

int sg_alloc_state(vrec_sg *g) {
    vrec_ctx *ctx = g->ctx;
    const int64_t rows = g->row_hi - g->row_lo;
    // x[0] | x[1] | exchange block in ONE allocation: a peer maps the whole graph state with one IPC handle
    g->xstride = (g->N + 31) / 32 * 32;
    const size_t xchg_doubles = (sizeof(SgExchange) + 7) / 8;
    VREC_TRY(g->d_xbuf.alloc((size_t)(2 * g->xstride) + xchg_doubles));
    g->d_x[0] = g->d_xbuf.p;
    g->d_x[1] = g->d_xbuf.p + g->xstride;
    g->d_xchg = (SgExchange *)(g->d_xbuf.p + 2 * g->xstride);
    VREC_CUDA(cudaMemsetAsync(g->d_xbuf.p, 0, g->d_xbuf.bytes(), ctx->stream));
    // grid: a fixed function of the row count only, so the residual order is reproducible
    int64_t want = (rows + 32 * SPMV_WARPS - 1) / (32 * SPMV_WARPS);
    g->grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, 148 * 8));
    VREC_TRY(g->d_block_partials.alloc(g->grid));
    VREC_TRY(g->d_state.alloc(1));
    return VREC_OK;
}

inline int sg_count_blocks(int64_t N) { return N > SRC_BLOCK ? (int)((N + SRC_BLOCK - 1) / SRC_BLOCK) : 1; }

// host-built graphs: rowptr / src / w are host copies of this process's rows (sources ascending in each
// row); splits them into the per-source-block CSRs and uploads those
int sg_setup_device(vrec_sg *g, const std::vector<int> &rowptr, const int *h_src, const double *h_w) {
    vrec_ctx *ctx = g->ctx;
    const int64_t rows = g->row_hi - g->row_lo;
    g->nblocks = sg_count_blocks(g->N);
    g->blocks.clear();
    if (g->nblocks == 1) {
        g->blocks.emplace_back(new vrec_sg::Block());
        vrec_sg::Block &b = *g->blocks[0];
        b.nnz = g->nnz;
        VREC_TRY(b.rowptr.upload(rowptr.data(), rowptr.size(), ctx->stream));
        VREC_TRY(b.src.upload(h_src, (size_t)g->nnz, ctx->stream));
        VREC_TRY(b.w.upload(h_w, (size_t)g->nnz, ctx->stream));
        VREC_TRY(sg_fill_block(g, b, rowptr.data(), rows));
    } else {
        std::vector<int> cut((size_t)(g->nblocks + 1) * (size_t)rows);   // boundaries of every block inside every row
        for (int64_t r = 0; r < rows; ++r) {
            const int *lo = h_src + rowptr[r], *hi = h_src + rowptr[r + 1];
            cut[r] = rowptr[r];
            for (int b = 1; b < g->nblocks; ++b)
                cut[(size_t)b * rows + r] = (int)(std::lower_bound(lo, hi, (int)((int64_t)b * SRC_BLOCK)) - h_src);
            cut[(size_t)g->nblocks * rows + r] = rowptr[r + 1];
        }
        std::vector<int> brp((size_t)rows + 1), bsrc;
        std::vector<double> bw;
        for (int b = 0; b < g->nblocks; ++b) {
            const int *c0 = cut.data() + (size_t)b * rows, *c1 = cut.data() + (size_t)(b + 1) * rows;
            brp[0] = 0;
            for (int64_t r = 0; r < rows; ++r) brp[r + 1] = brp[r] + (c1[r] - c0[r]);
            bsrc.resize((size_t)brp[rows]);
            bw.resize((size_t)brp[rows]);
            for (int64_t r = 0; r < rows; ++r) {
                std::copy(h_src + c0[r], h_src + c1[r], bsrc.begin() + brp[r]);
                std::copy(h_w + c0[r], h_w + c1[r], bw.begin() + brp[r]);
            }
            g->blocks.emplace_back(new vrec_sg::Block());
            vrec_sg::Block &blk = *g->blocks[b];
            blk.nnz = brp[rows];
            VREC_TRY(blk.rowptr.upload(brp.data(), brp.size(), ctx->stream));
            VREC_TRY(blk.src.upload(bsrc.data(), bsrc.size(), ctx->stream));
            VREC_TRY(blk.w.upload(bw.data(), bw.size(), ctx->stream));
            VREC_TRY(sg_fill_block(g, blk, brp.data(), rows));
            VREC_CUDA(cudaStreamSynchronize(ctx->stream));          // brp / bsrc / bw are reused
        }
    }
    VREC_TRY(sg_alloc_state(g));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    return VREC_OK;
}

// ---- device-generated rows (every row has at most VREC_CANON_SEG in-edges: no long tables) ----
// cut[b * n_rows + r] = first in-edge of row r whose source belongs to block b or later
__global__ void sg_block_cut_kernel(int n_rows, const int *__restrict__ rowptr, const int *__restrict__ src, int nblocks,
                                    int *__restrict__ cut) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_rows * (nblocks + 1)) return;
    int b = (int)(i / n_rows), r = (int)(i % n_rows);
    int lo = rowptr[r], hi = rowptr[r + 1];
    if (b == 0) {
        cut[i] = lo;
    } else if (b == nblocks) {
        cut[i] = hi;
    } else {
        const int key = (int)((long long)b * SRC_BLOCK);
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (src[mid] < key) lo = mid + 1; else hi = mid;
        }
        cut[i] = lo;
    }
}

// one warp per row: copies the row's part [c0, c1) of the row-major arrays to its place in the block's CSR
__global__ void sg_block_copy_kernel(int n_rows, const int *__restrict__ c0, const int *__restrict__ c1,
                                     const int *__restrict__ brp, const int *__restrict__ src,
                                     const double *__restrict__ w, int *__restrict__ bsrc, double *__restrict__ bw) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += (long long)gridDim.x * wpb) {
        const int s = c0[r], n = c1[r] - s, d = brp[r];
        for (int k = lane; k < n; k += 32) {
            bsrc[d + k] = src[s + k];
            bw[d + k] = w[s + k];
        }
    }
}

// rm_* = the generated rows in row-major order; they are released once the blocks are built
int sg_setup_generated(vrec_sg *g, DevBuf<int> &rm_rowptr, DevBuf<int> &rm_src, DevBuf<double> &rm_w) {
    vrec_ctx *ctx = g->ctx;
    const int64_t rows = g->row_hi - g->row_lo;
    g->nblocks = sg_count_blocks(g->N);
    g->blocks.clear();
    if (g->nblocks == 1) {
        g->blocks.emplace_back(new vrec_sg::Block());
        vrec_sg::Block &b = *g->blocks[0];
        b.nnz = g->nnz;
        std::swap(b.rowptr.p, rm_rowptr.p); std::swap(b.rowptr.n, rm_rowptr.n);
        std::swap(b.src.p, rm_src.p); std::swap(b.src.n, rm_src.n);
        std::swap(b.w.p, rm_w.p); std::swap(b.w.n, rm_w.n);
        VREC_TRY(b.long_rows.alloc(1));
        VREC_TRY(b.long_segptr.alloc(2));
        VREC_TRY(b.seg_row.alloc(1));
        VREC_TRY(b.partials.alloc(1));
    } else {
        DevBuf<int> d_cut, d_brp;
        VREC_TRY(d_cut.alloc((size_t)(g->nblocks + 1) * (size_t)rows));
        VREC_TRY(d_brp.alloc((size_t)rows + 1));
        long long n = (long long)rows * (g->nblocks + 1);
        sg_block_cut_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int)rows, rm_rowptr.p, rm_src.p,
                                                                               g->nblocks, d_cut.p);
        VREC_LAUNCHED(ctx);
        std::vector<int> cut((size_t)n), brp((size_t)rows + 1);
        VREC_CUDA(cudaMemcpyAsync(cut.data(), d_cut.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        VREC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int b = 0; b < g->nblocks; ++b) {
            const int *c0 = cut.data() + (size_t)b * rows, *c1 = cut.data() + (size_t)(b + 1) * rows;
            brp[0] = 0;
            for (int64_t r = 0; r < rows; ++r) brp[r + 1] = brp[r] + (c1[r] - c0[r]);
            g->blocks.emplace_back(new vrec_sg::Block());
            vrec_sg::Block &blk = *g->blocks[b];
            blk.nnz = brp[rows];
            VREC_TRY(blk.rowptr.upload(brp.data(), brp.size(), ctx->stream));
            VREC_TRY(blk.src.alloc((size_t)blk.nnz));
            VREC_TRY(blk.w.alloc((size_t)blk.nnz));
            sg_block_copy_kernel<<<148 * 8, 256, 0, ctx->stream>>>((int)rows, d_cut.p + (size_t)b * rows,
                                                                   d_cut.p + (size_t)(b + 1) * rows, blk.rowptr.p,
                                                                   rm_src.p, rm_w.p, blk.src.p, blk.w.p);
            VREC_LAUNCHED(ctx);
            VREC_TRY(blk.long_rows.alloc(1));
            VREC_TRY(blk.long_segptr.alloc(2));
            VREC_TRY(blk.seg_row.alloc(1));
            VREC_TRY(blk.partials.alloc(1));
            VREC_CUDA(cudaStreamSynchronize(ctx->stream));          // brp is reused
        }
        rm_rowptr.release();
        rm_src.release();
        rm_w.release();
    }
    VREC_TRY(sg_alloc_state(g));
    return VREC_OK;
}

// ---- peers of a row-partitioned graph ----
// Offset of a device pointer inside its allocation (an IPC handle always names the allocation's base).
int sg_alloc_offset(const void *p, size_t *off) {
    typedef int (*get_range_t)(unsigned long long *, size_t *, unsigned long long);
    static get_range_t fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &qr);
        if (e != cudaSuccess || !sym) {
            vrec_set_error("cuMemGetAddressRange is not available: %s", cudaGetErrorString(e));
            return VREC_ECUDA;
        }
        fn = (get_range_t)sym;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) {
        vrec_set_error("cuMemGetAddressRange failed");
        return VREC_ECUDA;
    }
    *off = (size_t)((unsigned long long)(uintptr_t)p - base);
    return VREC_OK;
}

int sg_publish_peers(vrec_sg *g) {
    VREC_TRY(g->d_peers.upload(&g->peers, 1, g->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(g->ctx->stream));
    g->peers_ready = true;
    return VREC_OK;
}

// one process per GPU: every rank maps every other rank's state buffer through CUDA IPC; the handles travel
// over the NCCL communicator the context already owns (vrec_comm_init)
int sg_setup_peers_ipc(vrec_sg *g) {
    vrec_ctx *ctx = g->ctx;
    const int W = ctx->world, me = ctx->rank;
    if (W > VREC_MAX_WORLD) {
        vrec_set_error("row-partitioned graphs support at most %d ranks", VREC_MAX_WORLD);
        return VREC_EINVAL;
    }
    struct Wire {
        cudaIpcMemHandle_t handle;
        unsigned long long offset;
    };
    Wire mine;
    memset(&mine, 0, sizeof mine);
    size_t off = 0;
    VREC_TRY(sg_alloc_offset(g->d_xbuf.p, &off));
    mine.offset = off;
    VREC_CUDA(cudaIpcGetMemHandle(&mine.handle, g->d_xbuf.p));
    DevBuf<unsigned char> d_send, d_recv;
    VREC_TRY(d_send.upload((const unsigned char *)&mine, sizeof mine, ctx->stream));
    VREC_TRY(d_recv.alloc(sizeof(Wire) * (size_t)W));
    VREC_TRY(vrec_comm_allgather_bytes(ctx, d_send.p, d_recv.p, sizeof(Wire)));
    std::vector<Wire> all((size_t)W);
    VREC_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, sizeof(Wire) * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    memset(&g->peers, 0, sizeof g->peers);
    g->peers.world = W;
    g->peers.rank = me;
    for (int r = 0; r < W; ++r) {
        double *base;
        if (r == me) {
            base = g->d_xbuf.p;
        } else {
            void *m = nullptr;
            VREC_CUDA(cudaIpcOpenMemHandle(&m, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
            g->ipc_opened.push_back(m);
            base = (double *)((char *)m + all[r].offset);
        }
        g->peers.x[0][r] = base;
        g->peers.x[1][r] = base + g->xstride;
        g->peers.xchg[r] = (SgExchange *)(base + 2 * g->xstride);
    }
    return sg_publish_peers(g);
}

}  // namespace

vrec_sg::~vrec_sg() {
    for (void *m : ipc_opened) cudaIpcCloseMemHandle(m);
    if (loop_exec) cudaGraphExecDestroy((cudaGraphExec_t)loop_exec);
    if (loop_graph) cudaGraphDestroy((cudaGraph_t)loop_graph);
}

// ---------------------------------------------------------------------------------------
// The step() loop on the device.  Split in three so that a single-process group of ranks can be
// driven from one stream (vrec_sg_group_stationary): begin / one sweep / the exchange barrier.
// ---------------------------------------------------------------------------------------
static float sg_keep_frac(const vrec_sg *g) {
    // pin at most ~60 MiB of x in L2 (measured: 64 MB of evict_last data stays resident on B200
    // next to the streamed matrix, 80 MB does not)
    const double keep_bytes = 60.0 * 1024 * 1024;
    const double gathered = 8.0 * (double)std::min<int64_t>(std::max<int64_t>(1, g->N), SRC_BLOCK);
    return (float)std::min(1.0, keep_bytes / gathered);
}

static unsigned long long sg_flag(const vrec_sg *g, int step) { return (g->epoch << 32) | (unsigned)step; }

int sg_run_begin(vrec_sg *g, long long uidx, double eps2, int max_it, bool check_convergence, bool spin) {
    vrec_ctx *ctx = g->ctx;
    const double x0 = 1.0 / (double)g->N;                   // :53-54
    if (g->partitioned) {
        if (!g->peers_ready) {
            vrec_set_error("row-partitioned graph: peers are not connected");
            return VREC_EINVAL;
        }
        g->epoch++;
    }
    sg_reset_state_kernel<<<1, 1, 0, ctx->stream>>>(g->d_state.p, max_it, uidx, eps2, check_convergence ? 1 : 0,
                                                    sg_flag(g, 0));
    VREC_LAUNCHED(ctx);
    int fill_grid = (int)std::min<int64_t>((g->N + 255) / 256, 148 * 16);
    sg_fill_kernel<<<std::max(1, fill_grid), 256, 0, ctx->stream>>>(g->d_x[0], g->N, x0);
    VREC_LAUNCHED(ctx);
    if (g->partitioned && max_it > 0) {
        // nobody writes x[1] of a peer that may still be reading the result of its previous query
        sg_exchange_signal_kernel<<<1, 32, 0, ctx->stream>>>(g->d_peers.p, sg_flag(g, 0));
        VREC_LAUNCHED(ctx);
        if (spin) {
            sg_exchange_wait_kernel<<<1, 32, 0, ctx->stream>>>(g->d_state.p, g->d_xchg, g->peers.world, 0, 0, 1);
            VREC_LAUNCHED(ctx);
        }
    }
    return VREC_OK;
}

// one sweep x -> x' over all source blocks; which iteration it is, the start vertex, epsilon and the buffers'
// roles come from the device-resident SgState, so the same launches can be replayed
int sg_run_sweep(vrec_sg *g, cudaGraphConditionalHandle loop_cond = 0) {
    vrec_ctx *ctx = g->ctx;
    const int rows = (int)(g->row_hi - g->row_lo);
    const float keep_frac = sg_keep_frac(g);
    for (int b = 0; b < g->nblocks; ++b) {
        vrec_sg::Block &blk = *g->blocks[b];
        if (blk.n_seg > 0) {
            int pg = (blk.n_seg + SPMV_WARPS - 1) / SPMV_WARPS;
            sg_long_partials_kernel<<<pg, SPMV_THREADS, 0, ctx->stream>>>(
                blk.rowptr.p, blk.rowptr.p + 1, blk.src.p, blk.w.p, g->d_x[0], g->d_x[1], blk.long_rows.p,
                blk.long_segptr.p, blk.seg_row.p, blk.n_seg, blk.partials.p, g->d_state.p, keep_frac);
            VREC_LAUNCHED(ctx);
        }
        // measured on 6 M vertices x 100 in-edges (tools/sg_bench.py): (4 CTAs/SM, 2 windows/batch) 2388 us,
        // (3, 4) 2509 us, (3, 3) 2535 us, (5, 2) 2571 us, (2, 4) 2665 us per iteration
        auto kern = g->rows_kernel ? sg_spmv_kernel<false, 4, 4>
                    : g->flat_variant == 1 ? sg_spmv_kernel<true, 3, 4>
                    : g->flat_variant == 2 ? sg_spmv_kernel<true, 3, 3>
                                           : sg_spmv_kernel<true, 4, 2>;
        kern<<<g->grid, SPMV_THREADS, 0, ctx->stream>>>(
            rows, g->row_lo, blk.rowptr.p, blk.rowptr.p + 1, blk.src.p, blk.w.p, g->d_x[0], g->d_x[1], blk.long_rows.p,
            blk.long_segptr.p, blk.n_long, blk.partials.p, g->d_state.p, g->d_block_partials.p, keep_frac,
            b > 0 ? 1 : 0, b == g->nblocks - 1 ? 1 : 0, g->partitioned ? g->d_peers.p : nullptr, loop_cond);
        VREC_LAUNCHED(ctx);
    }
    return VREC_OK;
}

int sg_run_barrier(vrec_sg *g, int it, bool check_convergence, bool spin) {
    if (!g->partitioned) return VREC_OK;
    vrec_ctx *ctx = g->ctx;
    sg_exchange_wait_kernel<<<1, 32, 0, ctx->stream>>>(g->d_state.p, g->d_xchg, g->peers.world, it + 1, 1,
                                                       spin ? 1 : 0);
    (void)check_convergence;      // the decision reads SgState::check
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}

// The step() loop of one graph as a CUDA graph: a conditional `while` node whose body is one sweep (its last CTA
// sets the loop condition from SgState::done).  A query then costs one graph launch whatever maxIterations is, and nothing runs after the
// iteration that converged (the reference pays one Spark action per iteration, :130-141; the round-1 engine
// queued all maxIterations sweeps and let the late ones exit on `done`).
static int sg_build_loop_graph(vrec_sg *g) {
    vrec_ctx *ctx = g->ctx;
    cudaGraph_t graph = nullptr;
    VREC_CUDA(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle handle;
    VREC_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    VREC_CUDA(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    const int64_t l0 = ctx->launches;
    VREC_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    int rc = sg_run_sweep(g, handle);            // the last CTA of the sweep sets the loop condition
    cudaGraph_t captured = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &captured);
    g->loop_launches = (int)(ctx->launches - l0);
    ctx->launches = l0;                                     // captured, not launched
    if (rc != VREC_OK || e != cudaSuccess) {
        if (e != cudaSuccess) vrec_set_error("cudaStreamEndCapture -> %s", cudaGetErrorString(e));
        cudaGraphDestroy(graph);
        return rc != VREC_OK ? rc : VREC_ECUDA;
    }
    cudaGraphExec_t exec = nullptr;
    VREC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    g->loop_graph = graph;
    g->loop_exec = exec;
    return VREC_OK;
}

// launches the whole step() loop for one vertex index; results stay on the device
int sg_run_device(vrec_sg *g, long long uidx, double epsilon, int max_it, bool check_convergence) {
    const double eps2 = epsilon * epsilon;                  // :40
    if (g->partitioned && g->peers_local) {
        vrec_set_error("this graph belongs to a single-process group: use vrec_sg_group_stationary");
        return VREC_EINVAL;
    }
    VREC_TRY(sg_run_begin(g, uidx, eps2, max_it, check_convergence, true));
    if (max_it <= 0) return VREC_OK;
    if (!g->partitioned && g->use_graph) {
        if (!g->loop_exec) VREC_TRY(sg_build_loop_graph(g));
        VREC_CUDA(cudaGraphLaunch((cudaGraphExec_t)g->loop_exec, g->ctx->stream));
        g->ctx->launches += g->loop_launches;               // at least one pass; sg_fetch_state adds the rest
        g->graph_pending = true;
        return VREC_OK;
    }
    for (int it = 0; it < max_it; ++it) {
        VREC_TRY(sg_run_sweep(g));
        VREC_TRY(sg_run_barrier(g, it, check_convergence, true));
    }
    return VREC_OK;
}

int sg_fetch_state(vrec_sg *g, int max_it, SgState *h, int *result_buf) {
    VREC_CUDA(cudaMemcpyAsync(h, g->d_state.p, sizeof(SgState), cudaMemcpyDeviceToHost, g->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (h->converged < 0) {
        vrec_set_error("row-partitioned graph: a peer did not reach the exchange barrier in time");
        return VREC_ENCCL;
    }
    if (g->graph_pending) {                                 // passes of the graph's body beyond the first
        g->ctx->launches += (int64_t)g->loop_launches * std::max(0, h->cur - 1);
        g->graph_pending = false;
    }
    if (max_it <= 0) {
        h->iterations = 0;
        h->converged = 0;
        *result_buf = 0;
    } else if (h->converged) {
        *result_buf = (h->iterations + 1) & 1;   // converged: step() returns nextX (:101)
    } else {
        *result_buf = max_it & 1;                // limit reached: returns x (:95)
    }
    return VREC_OK;
}

int64_t sg_lookup(const vrec_sg *g, int64_t id) {
    auto it = std::lower_bound(g->h_ids.begin(), g->h_ids.end(), id);
    if (it == g->h_ids.end() || *it != id) return -1;
    return it - g->h_ids.begin();
}

// ---------------------------------------------------------------------------------------
// Host-side construction of the vertex table and the CSR of P^T (exposed for CPU tests).
//   ids      distinct(source ∪ target) ascending            (:42-49)
//   rowptr   in-edge ranges per target vertex
//   src/w    in-edges in ascending source index, stable for duplicate (s,t) pairs
// ---------------------------------------------------------------------------------------
int vrec_host_build_sg(int64_t nnz, const int64_t *source, const int64_t *target, const double *weight,
                       std::vector<int64_t> &ids, std::vector<int> &rowptr, std::vector<int> &src,
                       std::vector<double> &w) {
    if (nnz < 0 || nnz >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_sg_load: nnz %lld out of range [0, 2^31-1)", (long long)nnz);
        return VREC_EINVAL;
    }
    ids.resize((size_t)(2 * nnz));
    for (int64_t e = 0; e < nnz; ++e) {
        ids[2 * e] = source[e];
        ids[2 * e + 1] = target[e];
    }
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    const int64_t N = (int64_t)ids.size();
    std::vector<int> si((size_t)nnz), ti((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) {
        si[e] = (int)(std::lower_bound(ids.begin(), ids.end(), source[e]) - ids.begin());
        ti[e] = (int)(std::lower_bound(ids.begin(), ids.end(), target[e]) - ids.begin());
    }
    // stable counting sort by source, then by target
    std::vector<int64_t> cnt((size_t)N + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) cnt[si[e] + 1]++;
    for (int64_t i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    std::vector<int> ord((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) ord[cnt[si[e]]++] = (int)e;
    rowptr.assign((size_t)N + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) rowptr[ti[e] + 1]++;
    for (int64_t i = 0; i < N; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int> pos(rowptr.begin(), rowptr.end() - 1);
    src.resize((size_t)nnz);
    w.resize((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) {
        int e = ord[k];
        int p = pos[ti[e]]++;
        src[p] = si[e];
        w[p] = weight[e];
    }
    return VREC_OK;
}

extern "C" int vrec_host_sg_csr(int64_t nnz, const int64_t *source, const int64_t *target,
                                const double *weight, int64_t *out_n, int64_t *out_ids /*[2*nnz]*/,
                                int32_t *out_rowptr /*[2*nnz+1]*/, int32_t *out_src, double *out_w) {
    std::vector<int64_t> ids;
    std::vector<int> rowptr, src;
    std::vector<double> w;
    VREC_TRY(vrec_host_build_sg(nnz, source, target, weight, ids, rowptr, src, w));
    *out_n = (int64_t)ids.size();
    std::copy(ids.begin(), ids.end(), out_ids);
    std::copy(rowptr.begin(), rowptr.end(), out_rowptr);
    std::copy(src.begin(), src.end(), out_src);
    std::copy(w.begin(), w.end(), out_w);
    return VREC_OK;
}

extern "C" int vrec_sg_load(vrec_ctx *ctx, int64_t nnz, const int64_t *source_id, const int64_t *target_id,
                            const double *balanced_weight, vrec_sg **out) {
    if (!ctx || !out || (nnz > 0 && (!source_id || !target_id || !balanced_weight))) {
        vrec_set_error("vrec_sg_load: NULL argument");
        return VREC_EINVAL;
    }
    *out = nullptr;
    VREC_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> rowptr, src;
    std::vector<double> w;
    vrec_sg *g = new vrec_sg();
    g->ctx = ctx;
    int rc = vrec_host_build_sg(nnz, source_id, target_id, balanced_weight, g->h_ids, rowptr, src, w);
    if (rc == VREC_OK) {
        g->N = (int64_t)g->h_ids.size();
        g->nnz = nnz;
        g->row_lo = 0;
        g->row_hi = g->N;
        rc = g->d_ids.upload((const long long *)g->h_ids.data(), g->h_ids.size(), ctx->stream);
    }
    if (rc == VREC_OK) rc = sg_setup_device(g, rowptr, src.data(), w.data());
    if (rc == VREC_OK) rc = sg_batch_analyse(g, rowptr, src.data(), w.data());
    if (rc != VREC_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return VREC_OK;
}

// Row ranges of a row-partitioned graph: contiguous, balanced by in-edges (rank k starts at the first row whose
// rowptr reaches k * nnz / world), the same on every rank.
void vrec_sg_partition_bounds(const std::vector<int> &rowptr, int world, std::vector<int64_t> &bounds) {
    const int64_t N = (int64_t)rowptr.size() - 1, nnz = rowptr[N];
    bounds.assign((size_t)world + 1, N);
    bounds[0] = 0;
    for (int k = 1; k < world; ++k) {
        const int64_t want = nnz / world * k + (nnz % world) * k / world;
        int64_t r = std::lower_bound(rowptr.begin(), rowptr.end(), (int)std::min<int64_t>(want, 0x7fffffff)) - rowptr.begin();
        bounds[k] = std::max(bounds[k - 1], std::min<int64_t>(r, N));
    }
}

extern "C" int vrec_host_sg_partition(int64_t n_rows, const int32_t *rowptr, int32_t world, int64_t *out_bounds) {
    if (!rowptr || !out_bounds || world < 1 || n_rows < 0) return VREC_EINVAL;
    std::vector<int> rp(rowptr, rowptr + n_rows + 1);
    std::vector<int64_t> b;
    vrec_sg_partition_bounds(rp, world, b);
    std::copy(b.begin(), b.end(), out_bounds);
    return VREC_OK;
}

static int sg_load_rows(vrec_ctx *ctx, int rank, int world, int64_t nnz, const int64_t *source_id,
                        const int64_t *target_id, const double *balanced_weight, vrec_sg **out) {
    std::vector<int> rowptr, src;
    std::vector<double> w;
    vrec_sg *g = new vrec_sg();
    g->ctx = ctx;
    int rc = vrec_host_build_sg(nnz, source_id, target_id, balanced_weight, g->h_ids, rowptr, src, w);
    if (rc == VREC_OK) {
        g->N = (int64_t)g->h_ids.size();
        std::vector<int64_t> bounds;
        vrec_sg_partition_bounds(rowptr, world, bounds);
        g->row_lo = bounds[rank];
        g->row_hi = bounds[rank + 1];
        g->partitioned = world > 1;
        const int e0 = rowptr[g->row_lo], e1 = rowptr[g->row_hi];
        g->nnz = e1 - e0;
        std::vector<int> lrp(rowptr.begin() + g->row_lo, rowptr.begin() + g->row_hi + 1);
        for (int &v : lrp) v -= e0;
        rc = g->d_ids.upload((const long long *)g->h_ids.data(), g->h_ids.size(), ctx->stream);
        if (rc == VREC_OK) rc = sg_setup_device(g, lrp, src.data() + e0, w.data() + e0);
    }
    if (rc != VREC_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return VREC_OK;
}

// Row-partitioned load of one oversized graph: every rank passes the whole edge list and keeps a contiguous
// range of rows of P^T (balanced by in-edges); during a sweep every rank stores its rows of x' straight into
// the peers' buffers (sg_spmv_kernel), so no separate exchange step exists.
extern "C" int vrec_sg_load_partitioned(vrec_ctx *ctx, int64_t nnz, const int64_t *source_id,
                                        const int64_t *target_id, const double *balanced_weight, vrec_sg **out) {
    if (!ctx || !out || (nnz > 0 && (!source_id || !target_id || !balanced_weight))) {
        vrec_set_error("vrec_sg_load_partitioned: NULL argument");
        return VREC_EINVAL;
    }
    *out = nullptr;
    VREC_CUDA(cudaSetDevice(ctx->device));
    vrec_sg *g = nullptr;
    VREC_TRY(sg_load_rows(ctx, ctx->rank, ctx->world, nnz, source_id, target_id, balanced_weight, &g));
    if (g->partitioned) {
        int rc = sg_setup_peers_ipc(g);
        if (rc != VREC_OK) {
            delete g;
            return rc;
        }
    }
    *out = g;
    return VREC_OK;
}

// ---------------------------------------------------------------------------------------
// Single-process group: the `world` parts of one row-partitioned graph held by ONE process (all on the
// context's device, or on several devices with peer access), driven from the calling thread.  Same kernels,
// same peer stores, same residual slots as the one-process-per-GPU path; the exchange barrier is the stream
// order instead of a spin on the peers' flags.  It is how the partition logic is checked where only one GPU
// is visible (tests), and it serves hosts that drive several GPUs from one process.
// ---------------------------------------------------------------------------------------
extern "C" int vrec_sg_group_load(vrec_ctx *ctx, int32_t world, int64_t nnz, const int64_t *source_id,
                                  const int64_t *target_id, const double *balanced_weight, vrec_sg **out_parts) {
    if (!ctx || !out_parts || world < 1 || world > VREC_MAX_WORLD ||
        (nnz > 0 && (!source_id || !target_id || !balanced_weight))) {
        vrec_set_error("vrec_sg_group_load: bad argument");
        return VREC_EINVAL;
    }
    VREC_CUDA(cudaSetDevice(ctx->device));
    std::vector<vrec_sg *> parts((size_t)world, nullptr);
    int rc = VREC_OK;
    for (int r = 0; r < world && rc == VREC_OK; ++r)
        rc = sg_load_rows(ctx, r, world, nnz, source_id, target_id, balanced_weight, &parts[r]);
    for (int r = 0; r < world && rc == VREC_OK; ++r) {
        vrec_sg *g = parts[r];
        g->partitioned = true;                      // also for world == 1: exercises the exchange code
        g->peers_local = true;
        memset(&g->peers, 0, sizeof g->peers);
        g->peers.world = world;
        g->peers.rank = r;
        for (int q = 0; q < world; ++q) {
            g->peers.x[0][q] = parts[q]->d_x[0];
            g->peers.x[1][q] = parts[q]->d_x[1];
            g->peers.xchg[q] = parts[q]->d_xchg;
        }
        rc = sg_publish_peers(g);
    }
    if (rc != VREC_OK) {
        for (vrec_sg *g : parts) delete g;
        return rc;
    }
    std::copy(parts.begin(), parts.end(), out_parts);
    return VREC_OK;
}

extern "C" int vrec_sg_group_stationary(vrec_sg **parts, int32_t world, int64_t vertex, double epsilon,
                                        int32_t max_iterations, double *out_x /* [world x N] */,
                                        int32_t *out_iterations, int32_t *out_converged, double *out_residual) {
    if (!parts || world < 1 || !out_x) return VREC_EINVAL;
    VREC_TRY(sg_check_params(epsilon, max_iterations));
    vrec_sg *g0 = parts[0];
    VREC_CUDA(cudaSetDevice(g0->ctx->device));
    int64_t v = sg_lookup(g0, vertex);
    if (v < 0) {
        vrec_set_error("No such vertex in the graph: %lld", (long long)vertex);
        return VREC_ENOENT;
    }
    const double eps2 = epsilon * epsilon;
    for (int r = 0; r < world; ++r) VREC_TRY(sg_run_begin(parts[r], v, eps2, max_iterations, true, false));
    for (int it = 0; it < max_iterations; ++it) {
        for (int r = 0; r < world; ++r) VREC_TRY(sg_run_sweep(parts[r]));
        for (int r = 0; r < world; ++r) VREC_TRY(sg_run_barrier(parts[r], it, true, false));
    }
    for (int r = 0; r < world; ++r) {
        vrec_sg *g = parts[r];
        SgState st;
        int buf = 0;
        VREC_TRY(sg_fetch_state(g, max_iterations, &st, &buf));
        VREC_CUDA(cudaMemcpyAsync(out_x + (size_t)r * (size_t)g->N, g->d_x[buf], sizeof(double) * (size_t)g->N,
                                  cudaMemcpyDeviceToHost, g->ctx->stream));
        VREC_CUDA(cudaStreamSynchronize(g->ctx->stream));
        if (out_iterations) out_iterations[r] = st.iterations;
        if (out_converged) out_converged[r] = st.converged;
        if (out_residual) out_residual[r] = st.residual;
    }
    return VREC_OK;
}

extern "C" int vrec_sg_row_range(vrec_sg *sg, int64_t *out_lo, int64_t *out_hi) {
    if (!sg || !out_lo || !out_hi) return VREC_EINVAL;
    *out_lo = sg->row_lo;
    *out_hi = sg->row_hi;
    return VREC_OK;
}

extern "C" void vrec_sg_free(vrec_sg *sg) {
    if (!sg) return;
    cudaSetDevice(sg->ctx->device);
    cudaStreamSynchronize(sg->ctx->stream);
    delete sg;
}

extern "C" int64_t vrec_sg_vertex_count(vrec_sg *sg) { return sg ? sg->N : 0; }
extern "C" int64_t vrec_sg_edge_count(vrec_sg *sg) { return sg ? sg->nnz : 0; }

extern "C" int vrec_sg_vertex_ids(vrec_sg *sg, int64_t *out_ids) {
    if (!sg || !out_ids) return VREC_EINVAL;
    if (!sg->h_ids.empty()) {
        std::copy(sg->h_ids.begin(), sg->h_ids.end(), out_ids);
    } else {
        for (int64_t i = 0; i < sg->N; ++i) out_ids[i] = i;   // generated graphs: ids are 0..N-1
    }
    return VREC_OK;
}

extern "C" int64_t vrec_sg_resident_bytes(vrec_sg *sg) {
    if (!sg) return 0;
    size_t b = sg->d_xbuf.bytes() + sg->d_ids.bytes();
    for (auto &blk : sg->blocks) b += blk->rowptr.bytes() + blk->src.bytes() + blk->w.bytes();
    return (int64_t)b;
}

static int sg_check_params(double epsilon, int32_t max_iterations) {
    if (!(epsilon >= 0)) {          // :33
        vrec_set_error("requirement failed: epsilon must be non-negative");
        return VREC_EINVAL;
    }
    if (max_iterations < 0) {       // :34
        vrec_set_error("requirement failed: max iterations number must be non-negative");
        return VREC_EINVAL;
    }
    return VREC_OK;
}

extern "C" int vrec_sg_stationary(vrec_sg *sg, int64_t vertex, double epsilon, int32_t max_iterations,
                                  double *out_x, int32_t *out_iterations, int32_t *out_converged,
                                  double *out_residual) {
    if (!sg || !out_x) return VREC_EINVAL;
    VREC_TRY(sg_check_params(epsilon, max_iterations));
    VREC_CUDA(cudaSetDevice(sg->ctx->device));
    int64_t v = sg->h_ids.empty() ? ((vertex >= 0 && vertex < sg->N) ? vertex : -1) : sg_lookup(sg, vertex);
    if (v < 0) {
        vrec_set_error("No such vertex in the graph: %lld", (long long)vertex);   // :70
        return VREC_ENOENT;
    }
    VREC_TRY(sg_run_device(sg, v, epsilon, max_iterations, true));
    SgState st;
    int buf = 0;
    VREC_TRY(sg_fetch_state(sg, max_iterations, &st, &buf));
    VREC_CUDA(cudaMemcpyAsync(out_x, sg->d_x[buf], sizeof(double) * (size_t)sg->N, cudaMemcpyDeviceToHost,
                              sg->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(sg->ctx->stream));
    if (out_iterations) *out_iterations = st.iterations;
    if (out_converged) *out_converged = st.converged;
    if (out_residual) *out_residual = st.residual;
    return VREC_OK;
}

extern "C" int vrec_sg_query(vrec_sg *sg, const int64_t *vertices, int32_t n, double epsilon,
                             int32_t max_iterations, const int64_t *place_filter, int64_t n_filter,
                             int32_t max_recs, int64_t *out_id, double *out_prob, int32_t *out_count,
                             int32_t *out_iterations, int32_t *out_converged, int32_t *out_status) {
    if (!sg || n < 0 || (n > 0 && (!vertices || !out_count || !out_status))) {
        vrec_set_error("vrec_sg_query: NULL argument");
        return VREC_EINVAL;
    }
    VREC_TRY(sg_check_params(epsilon, max_iterations));
    if (max_recs < 0 || n_filter < 0) {
        vrec_set_error("Maximum recommendations number must be non-negative");
        return VREC_EINVAL;
    }
    vrec_ctx *ctx = sg->ctx;
    VREC_CUDA(cudaSetDevice(ctx->device));
    const bool use_filter = place_filter != nullptr;
    const long long n_cand = use_filter ? n_filter : sg->N;
    if (use_filter)
        VREC_TRY(sg->d_filter_ids.upload((const long long *)place_filter, (size_t)n_filter, ctx->stream));
    VREC_TRY(sg->d_cand_val.ensure((size_t)std::max<long long>(1, n_cand)));
    VREC_TRY(sg->d_cand_key.ensure((size_t)std::max<long long>(1, n_cand)));
    const int m = std::max(1, (int)max_recs);
    VREC_TRY(sg->d_out_key.ensure(m));
    VREC_TRY(sg->d_out_val.ensure(m));
    VREC_TRY(sg->d_out_count.ensure(1));
    std::vector<long long> h_key(m);
    std::vector<double> h_val(m);
    // Start vertices without in-edges (every person) go to the batch kernel when there are enough of
    // them and the shared first iteration is not already the answer.
    std::vector<char> batched((size_t)n, 0);
    SgBatch &bt = sg->batch;
    bt.last_batched = 0;
    if (bt.ok && bt.mode > 0 && max_iterations >= 1 && n > 0) {
        std::vector<int> qidx, qvertex;
        for (int q = 0; q < n; ++q) {
            int64_t v = sg_lookup(sg, vertices[q]);
            if (v >= 0 && bt.h_act_of[v] < 0) {
                qidx.push_back(q);
                qvertex.push_back((int)v);
            }
        }
        if (!qidx.empty() && (bt.mode == 2 || qidx.size() >= 4)) {
            VREC_TRY(sg_batch_prepare(sg));
            // residual of iteration 0 with the start vertex's own term (0.15 - x0)^2 in place of (0 - x0)^2
            const double x0 = 1.0 / (double)sg->N;
            const double r1 = bt.r1_base - x0 * x0 + (kAlpha - x0) * (kAlpha - x0);
            if (!(r1 <= epsilon * epsilon)) {
                VREC_TRY(sg_batch_query(sg, qidx, qvertex, epsilon, max_iterations, place_filter, n_filter, max_recs,
                                        out_id, out_prob, out_count, out_iterations, out_converged));
                for (int q : qidx) {
                    batched[q] = 1;
                    out_status[q] = VREC_OK;
                }
            }
        }
    }
    for (int q = 0; q < n; ++q) {
        if (batched[q]) continue;
        out_count[q] = 0;
        if (out_iterations) out_iterations[q] = 0;
        if (out_converged) out_converged[q] = 0;
        int64_t v = sg->h_ids.empty() ? ((vertices[q] >= 0 && vertices[q] < sg->N) ? vertices[q] : -1)
                                      : sg_lookup(sg, vertices[q]);
        if (v < 0) {
            vrec_set_error("No such vertex in the graph: %lld", (long long)vertices[q]);
            out_status[q] = VREC_ENOENT;
            continue;
        }
        out_status[q] = VREC_OK;
        VREC_TRY(sg_run_device(sg, v, epsilon, max_iterations, true));
        // everything below is queued behind the loop: one synchronisation per query
        const bool want_recs = max_recs > 0 && n_cand > 0;
        int cnt = 0;
        if (want_recs) {
            int cg = (int)((n_cand + 255) / 256);
            sg_candidates_kernel<<<cg, 256, 0, ctx->stream>>>(
                sg->h_ids.empty() ? nullptr : sg->d_ids.p, sg->N, use_filter ? sg->d_filter_ids.p : nullptr, n_filter,
                sg->d_x[0], sg->d_x[1], sg->d_state.p, (long long)vertices[q], sg->d_cand_val.p, sg->d_cand_key.p);
            VREC_LAUNCHED(ctx);
            VREC_TRY(vrec_launch_select_topn(ctx, sg->d_cand_val.p, sg->d_cand_key.p, nullptr, n_cand, 0, 1,
                                             max_recs, sg->d_out_key.p, sg->d_out_val.p, sg->d_out_count.p));
            VREC_CUDA(cudaMemcpyAsync(&cnt, sg->d_out_count.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            VREC_CUDA(cudaMemcpyAsync(h_key.data(), sg->d_out_key.p, sizeof(long long) * m, cudaMemcpyDeviceToHost,
                                      ctx->stream));
            VREC_CUDA(cudaMemcpyAsync(h_val.data(), sg->d_out_val.p, sizeof(double) * m, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        }
        SgState st;
        int buf = 0;
        VREC_TRY(sg_fetch_state(sg, max_iterations, &st, &buf));          // synchronises
        if (out_iterations) out_iterations[q] = st.iterations;
        if (out_converged) out_converged[q] = st.converged;
        if (!want_recs) continue;
        out_count[q] = cnt;
        for (int k = 0; k < cnt; ++k) {
            out_id[(size_t)q * max_recs + k] = h_key[k];
            out_prob[(size_t)q * max_recs + k] = h_val[k];
        }
    }
    return VREC_OK;
}

// "batch": 0 = per-query kernels only, 1 = batch kernel for >= 4 eligible start vertices (default),
// 2 = batch kernel for every eligible start vertex; "batch_targets_per_cta": 0 = auto, 1 or 2.
extern "C" int vrec_sg_set_option(vrec_sg *sg, const char *name, int32_t value) {
    if (!sg || !name) return VREC_EINVAL;
    std::string k(name);
    if (k == "batch" && value >= 0 && value <= 2) {
        sg->batch.mode = value;
    } else if (k == "batch_targets_per_cta" && (value >= 0 && value <= 2)) {
        sg->batch.force_t = value;
    } else if (k == "flat_variant" && value >= 0 && value <= 2) {
        sg->flat_variant = value;
    } else if (k == "graph" && (value == 0 || value == 1)) {
        // 0 = queue all maxIterations sweeps (late ones exit on `done`) instead of the CUDA-graph while loop
        sg->use_graph = value;
    } else if (k == "rows_kernel" && (value == 0 || value == 1)) {
        // A/B: 1 = the round-1 half-warp-per-row kernel instead of the flat-window kernel
        sg->rows_kernel = value != 0;
    } else {
        vrec_set_error("vrec_sg_set_option: unknown option or bad value: %s = %d", name, (int)value);
        return VREC_EINVAL;
    }
    return VREC_OK;
}

// what: 0 = start vertices the last vrec_sg_query served with the batch kernel, 1 = batch path
// available for this graph (0/1), 2 = active vertices, 3 = edges between active vertices.
extern "C" int64_t vrec_sg_batch_info(vrec_sg *sg, int32_t what) {
    if (!sg) return -1;
    switch (what) {
        case 0: return sg->batch.last_batched;
        case 1: return sg->batch.ok ? 1 : 0;
        case 2: return sg->batch.n_a;
        case 3: return sg->batch.r_nnz;
        case 4: return sg->batch.sell_nnz;
        case 5: return sg->batch.us_prepare;
        case 6: return sg->batch.us_kernel;
        case 7: return sg->batch.us_results;
        default: return -1;
    }
}

// Copies the device CSR of P^T (rows owned by this process) back to the host in row-major order, for tests:
// a row is the concatenation of its parts in the source blocks.
extern "C" int vrec_sg_export_csr(vrec_sg *sg, int32_t *out_rowptr, int32_t *out_src, double *out_w) {
    if (!sg || !out_rowptr || !out_src || !out_w) return VREC_EINVAL;
    cudaStream_t s = sg->ctx->stream;
    VREC_CUDA(cudaSetDevice(sg->ctx->device));
    const size_t rows = (size_t)(sg->row_hi - sg->row_lo);
    if (sg->nblocks == 1) {
        vrec_sg::Block &b = *sg->blocks[0];
        VREC_CUDA(cudaMemcpyAsync(out_rowptr, b.rowptr.p, sizeof(int) * (rows + 1), cudaMemcpyDeviceToHost, s));
        VREC_CUDA(cudaMemcpyAsync(out_src, b.src.p, sizeof(int) * (size_t)sg->nnz, cudaMemcpyDeviceToHost, s));
        VREC_CUDA(cudaMemcpyAsync(out_w, b.w.p, sizeof(double) * (size_t)sg->nnz, cudaMemcpyDeviceToHost, s));
        VREC_CUDA(cudaStreamSynchronize(s));
        return VREC_OK;
    }
    std::vector<std::vector<int>> rp((size_t)sg->nblocks), bs((size_t)sg->nblocks);
    std::vector<std::vector<double>> bw((size_t)sg->nblocks);
    for (int b = 0; b < sg->nblocks; ++b) {
        vrec_sg::Block &blk = *sg->blocks[b];
        rp[b].resize(rows + 1);
        bs[b].resize((size_t)blk.nnz);
        bw[b].resize((size_t)blk.nnz);
        VREC_CUDA(cudaMemcpyAsync(rp[b].data(), blk.rowptr.p, sizeof(int) * (rows + 1), cudaMemcpyDeviceToHost, s));
        VREC_CUDA(cudaMemcpyAsync(bs[b].data(), blk.src.p, sizeof(int) * (size_t)blk.nnz, cudaMemcpyDeviceToHost, s));
        VREC_CUDA(cudaMemcpyAsync(bw[b].data(), blk.w.p, sizeof(double) * (size_t)blk.nnz, cudaMemcpyDeviceToHost, s));
    }
    VREC_CUDA(cudaStreamSynchronize(s));
    size_t pos = 0;
    for (size_t r = 0; r < rows; ++r) {
        out_rowptr[r] = (int32_t)pos;
        for (int b = 0; b < sg->nblocks; ++b) {
            const int a = rp[b][r], e = rp[b][r + 1];
            std::copy(bs[b].begin() + a, bs[b].begin() + e, out_src + pos);
            std::copy(bw[b].begin() + a, bw[b].begin() + e, out_w + pos);
            pos += (size_t)(e - a);
        }
    }
    out_rowptr[rows] = (int32_t)pos;
    return VREC_OK;
}

extern "C" int vrec_sg_iterate_device(vrec_sg *sg, int32_t iterations) {
    if (!sg || iterations < 0) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(sg->ctx->device));
    return sg_run_device(sg, 0, 0.0, iterations, false);
}

// ---------------------------------------------------------------------------------------
// Synthetic graph generated on the device (bench config "oversized graph").
// Row t of P^T gets its in-edges directly: in-degree = out_degree for every vertex, sources
// 50 % uniform / 50 % skewed (u^3 * N) over a multiplicative permutation, ascending per row,
// weight = 1/out_degree so that P is row-stochastic in expectation.
// ---------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__global__ void sg_gen_rows_kernel(long long N, int deg, unsigned long long seed, long long row_lo,
                                   int n_rows, int *__restrict__ rowptr, int *__restrict__ src,
                                   double *__restrict__ w) {
    // one warp per row; deg <= 128: each lane draws deg/32 sources, then a bitonic-free
    // insertion by rank keeps ascending order (rank = number of smaller (value, slot) pairs)
    extern __shared__ int s_buf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int *buf = s_buf + warp * deg;
    for (long long r = (long long)blockIdx.x * wpb + warp; r < n_rows; r += (long long)gridDim.x * wpb) {
        long long gr = row_lo + r;
        for (int k = lane; k < deg; k += 32) {
            unsigned long long h = splitmix64(seed ^ (unsigned long long)(gr * 1024 + k) * 0x2545F4914F6CDD1DULL);
            double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            long long v;
            if (k & 1) {
                v = (long long)(u * (double)N);
            } else {
                long long z = (long long)(u * u * u * (double)N);
                v = (long long)(((unsigned long long)z * 2654435761ULL + 12345ULL) % (unsigned long long)N);
            }
            if (v >= N) v = N - 1;
            buf[k] = (int)v;
        }
        __syncwarp();
        for (int k = lane; k < deg; k += 32) {
            int v = buf[k], rank = 0;
            for (int j = 0; j < deg; ++j) {
                int o = buf[j];
                rank += (o < v) || (o == v && j < k);
            }
            src[(size_t)r * deg + rank] = v;
            w[(size_t)r * deg + rank] = 1.0 / (double)deg;
        }
        __syncwarp();
        if (lane == 0) rowptr[r] = (int)(r * deg);
        if (lane == 0 && r == n_rows - 1) rowptr[n_rows] = (int)((long long)n_rows * deg);
    }
}

}  // namespace

extern "C" int vrec_sg_generate(vrec_ctx *ctx, int64_t n_vertices, int32_t out_degree, uint64_t seed,
                                int32_t rank, int32_t world, vrec_sg **out) {
    if (!ctx || !out || n_vertices <= 0 || out_degree <= 0 || out_degree > 1024 || world <= 0 || rank < 0 ||
        rank >= world) {
        vrec_set_error("vrec_sg_generate: bad argument");
        return VREC_EINVAL;
    }
    *out = nullptr;
    VREC_CUDA(cudaSetDevice(ctx->device));
    if (world > 1 && (ctx->world != world || ctx->rank != rank)) {
        vrec_set_error("vrec_sg_generate: rank/world (%d/%d) do not match vrec_comm_init (%d/%d)", rank, world,
                       ctx->rank, ctx->world);
        return VREC_EINVAL;
    }
    const int64_t slice = (n_vertices + world - 1) / world;      // every row has out_degree in-edges: equal rows = equal nnz
    int64_t lo = std::min<int64_t>(n_vertices, slice * rank), hi = std::min<int64_t>(n_vertices, lo + slice);
    int64_t rows = hi - lo;
    if (rows * out_degree >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_sg_generate: %lld edges per process exceed 2^31-1", (long long)(rows * out_degree));
        return VREC_EINVAL;
    }
    vrec_sg *g = new vrec_sg();
    g->ctx = ctx;
    g->N = n_vertices;
    g->nnz = rows * out_degree;
    g->row_lo = lo;
    g->row_hi = hi;
    g->partitioned = world > 1;
    DevBuf<int> rm_rowptr, rm_src;
    DevBuf<double> rm_w;
    int rc = rm_rowptr.alloc((size_t)rows + 1);
    if (rc == VREC_OK) rc = rm_src.alloc((size_t)g->nnz);
    if (rc == VREC_OK) rc = rm_w.alloc((size_t)g->nnz);
    if (rc == VREC_OK) rc = g->d_ids.alloc(1);
    if (rc == VREC_OK) {
        int wpb = 8;
        size_t smem = (size_t)wpb * out_degree * sizeof(int);
        int grid = (int)std::min<int64_t>((rows + wpb - 1) / wpb, 148 * 16);
        sg_gen_rows_kernel<<<grid, wpb * 32, smem, ctx->stream>>>(n_vertices, out_degree, seed, lo, (int)rows,
                                                                  rm_rowptr.p, rm_src.p, rm_w.p);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            vrec_set_error("sg_gen_rows_kernel -> %s", cudaGetErrorString(e));
            rc = VREC_ECUDA;
        }
    }
    if (rc == VREC_OK) {
        // every row has out_degree <= 1024 terms: no long (sub-)ranges; the source blocks are cut on the device
        rc = sg_setup_generated(g, rm_rowptr, rm_src, rm_w);
    }
    if (rc == VREC_OK && g->partitioned) rc = sg_setup_peers_ipc(g);
    if (rc == VREC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        vrec_set_error("vrec_sg_generate: %s", cudaGetErrorString(cudaGetLastError()));
        rc = VREC_ECUDA;
    }
    if (rc != VREC_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return VREC_OK;
}
