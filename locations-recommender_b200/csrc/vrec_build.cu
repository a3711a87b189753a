// Rating vectors builder on the device: the step in front of the KNN path.
//   RatingsBuilder.calcRatings            knn/RatingsBuilder.scala:32-48
//       count(*) per (person_id, entity); rank() over (partition by person_id order by count desc) <= topN
//       -- rank() keeps ties, so a person can keep more than topN entities
//   RatingVectorsBuilder.calcRatingVectors knn/RatingVectorsBuilder.scala:12-83
//       vector size = max(entity id) + 1 (:26-34, ids must fit an Int :36-41), one sparse vector per person,
//       indices ascending (:45-50), values = the counts as doubles (:69)
// All integer work: results are bit-identical to the oracle.  Output = the CSR layout vrec_knn_load takes
// (persons ascending, columns ascending), so the Parquet round trip between builder and recommender can go.
//
// Pipeline: two stable LSD radix sorts (entity, then person; cub::DeviceRadixSort -- the one library call, off
// the hot path) -> run heads of equal (person, entity) -> counts -> person segments -> rank filter (one warp
// per person) -> compaction into CSR.
#include <algorithm>
#include <cub/cub.cuh>

#include "vrec_internal.cuh"

namespace {

// Scratch of one builder call, from the stream-ordered allocator: the driver's memory pool keeps the
// blocks between calls (release threshold raised once), so a call does not pay ~25 cudaMalloc / cudaFree
// round trips with their device-wide synchronisations.
thread_local cudaStream_t g_pool_stream = nullptr;

template <typename T>
struct PoolBuf {
    T *p = nullptr;
    size_t n = 0;
    PoolBuf() = default;
    PoolBuf(const PoolBuf &) = delete;
    PoolBuf &operator=(const PoolBuf &) = delete;
    ~PoolBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, g_pool_stream);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), g_pool_stream);
        if (e != cudaSuccess) {
            p = nullptr;
            vrec_set_error("cudaMallocAsync(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
            return VREC_ENOMEM;
        }
        n = count;
        return VREC_OK;
    }
    int ensure(size_t count) { return count <= n ? VREC_OK : alloc(count); }
    int upload(const T *host, size_t count, cudaStream_t s) {
        VREC_TRY(alloc(count));
        if (count) VREC_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return VREC_OK;
    }
};

int pool_setup(vrec_ctx *ctx) {
    static thread_local int done_for = -1;
    g_pool_stream = ctx->stream;
    if (done_for != ctx->device) {
        cudaMemPool_t pool;
        VREC_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        unsigned long long keep = ~0ULL;
        VREC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        done_for = ctx->device;
    }
    return VREC_OK;
}

__global__ void bld_iota_kernel(long long n, long long *p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void bld_gather_kernel(long long n, const long long *__restrict__ idx, const long long *__restrict__ src,
                                  long long *__restrict__ dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// head[i] = 1 if row i starts a new (person, entity) run; phead[i] = 1 if it starts a new person
__global__ void bld_heads_kernel(long long n, const long long *__restrict__ person, const long long *__restrict__ entity,
                                 int *__restrict__ head, int *__restrict__ phead) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool np_ = i == 0 || person[i] != person[i - 1];
    head[i] = (np_ || entity[i] != entity[i - 1]) ? 1 : 0;
    phead[i] = np_ ? 1 : 0;
}

// one thread per sorted row: rows of a run add their weight to the run's count; run heads record the keys
__global__ void bld_runs_kernel(long long n, const long long *__restrict__ person, const long long *__restrict__ entity,
                                const long long *__restrict__ weight_sorted, const int *__restrict__ head,
                                const int *__restrict__ run_of, const int *__restrict__ prow_of,
                                long long *__restrict__ run_person_row, long long *__restrict__ run_entity,
                                unsigned long long *__restrict__ run_count, long long *__restrict__ person_ids,
                                int *__restrict__ person_first_run) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = run_of[i] - 1;                 // inclusive scan of the heads
    atomicAdd(run_count + r, (unsigned long long)(weight_sorted ? weight_sorted[i] : 1LL));
    if (head[i]) {
        const int p = prow_of[i] - 1;
        run_person_row[r] = p;
        run_entity[r] = entity[i];
        if (i == 0 || person[i] != person[i - 1]) {
            person_ids[p] = person[i];
            person_first_run[p] = r;
        }
    }
}

// rank() <= topN inside every person's runs, one warp per person: rank = 1 + #{runs with a larger count}
__global__ void bld_rank_kernel(int n_persons, int n_runs, const int *__restrict__ person_first_run,
                                const unsigned long long *__restrict__ run_count, int top_n, int *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const int p = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (p >= n_persons) return;
    const int s = person_first_run[p], e = p + 1 < n_persons ? person_first_run[p + 1] : n_runs;
    for (int r = s + lane; r < e; r += 32) {
        const unsigned long long c = run_count[r];
        int larger = 0;
        for (int q = s; q < e; ++q) larger += run_count[q] > c ? 1 : 0;
        keep[r] = larger + 1 <= top_n ? 1 : 0;
    }
}

__global__ void bld_scatter_kernel(int n_runs, const int *__restrict__ keep, const int *__restrict__ kept_incl,
                                   const long long *__restrict__ run_person_row, const long long *__restrict__ run_entity,
                                   const unsigned long long *__restrict__ run_count, int *__restrict__ out_col,
                                   double *__restrict__ out_val, unsigned long long *__restrict__ row_cnt,
                                   long long *__restrict__ max_entity, int *__restrict__ bad) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs || !keep[r]) return;
    const int at = kept_incl[r] - 1;
    const long long ent = run_entity[r];
    if (ent < 0 || ent > 0x7fffffffLL) {         // checkedCast, knn/RatingVectorsBuilder.scala:36-41
        *bad = 1;
        return;
    }
    out_col[at] = (int)ent;
    out_val[at] = (double)run_count[r];          // :69 rating.toDouble
    atomicAdd(row_cnt + run_person_row[r], 1ULL);
    atomicMax(max_entity, ent);
}

template <typename T>
int scan_inclusive(vrec_ctx *ctx, const int *in, T *out, long long n, PoolBuf<unsigned char> &tmp) {
    size_t bytes = 0;
    VREC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
    VREC_TRY(tmp.ensure(bytes));
    VREC_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, in, out, (int)n, ctx->stream));
    ctx->launches++;
    return VREC_OK;
}


// Everything the two entry points share: (src, dst[, weight]) rows on the device -> runs of equal (src, dst)
// with their counts, the distinct sources, and the rank() <= top_n filter.
struct FamilyCore {
    int n_runs = 0, n_src = 0, nnz = 0;
    PoolBuf<long long> run_src_row, run_dst, src_id;
    PoolBuf<unsigned long long> run_count;
    PoolBuf<int> first_run, keep, kept_incl;
};

int family_core(vrec_ctx *ctx, long long n, PoolBuf<long long> &d_src, PoolBuf<long long> &d_dst,
                PoolBuf<long long> *d_weight, int top_n, FamilyCore &c) {
    cudaStream_t st = ctx->stream;
    const int grid = (int)((n + 255) / 256);
    PoolBuf<long long> d_idx, k_a, k_b, v_a, v_b, d_wsorted;
    PoolBuf<unsigned char> tmp;
    VREC_TRY(k_a.alloc((size_t)n));
    VREC_TRY(k_b.alloc((size_t)n));
    VREC_TRY(v_a.alloc((size_t)n));
    VREC_TRY(v_b.alloc((size_t)n));
    // stable LSD: by dst, then by src; with weights the row index travels along and gathers them afterwards
    size_t bytes = 0;
    VREC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_dst.p, k_a.p, d_src.p, v_a.p, (int)n, 0, 64, st));
    VREC_TRY(tmp.ensure(bytes));
    if (d_weight) {
        VREC_TRY(d_idx.alloc((size_t)n));
        VREC_TRY(d_wsorted.alloc((size_t)n));
        bld_iota_kernel<<<grid, 256, 0, st>>>(n, d_idx.p);
        VREC_LAUNCHED(ctx);
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_dst.p, k_a.p, d_idx.p, v_a.p, (int)n, 0, 64, st));
        PoolBuf<long long> pg;
        VREC_TRY(pg.alloc((size_t)n));
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_a.p, d_src.p, pg.p);
        VREC_LAUNCHED(ctx);
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, pg.p, k_b.p, v_a.p, v_b.p, (int)n, 0, 64, st));
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_b.p, d_dst.p, k_a.p);
        VREC_LAUNCHED(ctx);
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_b.p, d_weight->p, d_wsorted.p);
        VREC_LAUNCHED(ctx);
        ctx->launches += 2;
    } else {
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_dst.p, k_a.p, d_src.p, v_a.p, (int)n, 0, 64, st));
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, v_a.p, k_b.p, k_a.p, v_b.p, (int)n, 0, 64, st));
        VREC_CUDA(cudaMemcpyAsync(k_a.p, v_b.p, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        ctx->launches += 2;
    }
    const long long *s_src = k_b.p, *s_dst = k_a.p;            // rows in (src, dst) order
    PoolBuf<int> head, phead, run_of, prow_of;
    VREC_TRY(head.alloc((size_t)n));
    VREC_TRY(phead.alloc((size_t)n));
    VREC_TRY(run_of.alloc((size_t)n));
    VREC_TRY(prow_of.alloc((size_t)n));
    bld_heads_kernel<<<grid, 256, 0, st>>>(n, s_src, s_dst, head.p, phead.p);
    VREC_LAUNCHED(ctx);
    VREC_TRY(scan_inclusive<int>(ctx, head.p, run_of.p, n, tmp));
    VREC_TRY(scan_inclusive<int>(ctx, phead.p, prow_of.p, n, tmp));
    VREC_CUDA(cudaMemcpyAsync(&c.n_runs, run_of.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&c.n_src, prow_of.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    VREC_TRY(c.run_src_row.alloc((size_t)c.n_runs));
    VREC_TRY(c.run_dst.alloc((size_t)c.n_runs));
    VREC_TRY(c.run_count.alloc((size_t)c.n_runs));
    VREC_TRY(c.src_id.alloc((size_t)c.n_src));
    VREC_TRY(c.first_run.alloc((size_t)c.n_src));
    VREC_TRY(c.keep.alloc((size_t)c.n_runs));
    VREC_TRY(c.kept_incl.alloc((size_t)c.n_runs));
    VREC_CUDA(cudaMemsetAsync(c.run_count.p, 0, sizeof(unsigned long long) * (size_t)c.n_runs, st));
    bld_runs_kernel<<<grid, 256, 0, st>>>(n, s_src, s_dst, d_weight ? d_wsorted.p : nullptr, head.p, run_of.p, prow_of.p,
                                          c.run_src_row.p, c.run_dst.p, c.run_count.p, c.src_id.p, c.first_run.p);
    VREC_LAUNCHED(ctx);
    bld_rank_kernel<<<(int)(((long long)c.n_src * 32 + 255) / 256), 256, 0, st>>>(c.n_src, c.n_runs, c.first_run.p,
                                                                                 c.run_count.p, top_n, c.keep.p);
    VREC_LAUNCHED(ctx);
    VREC_TRY(scan_inclusive<int>(ctx, c.keep.p, c.kept_incl.p, c.n_runs, tmp));
    VREC_CUDA(cudaMemcpyAsync(&c.nnz, c.kept_incl.p + (c.n_runs - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    return VREC_OK;
}

// edge family: per source the sum of the kept counts (one warp per source), then weight = count / total * beta
// for every kept run (PersonLikesPlace.scala:26-36 and the other three families; StochasticGraphBuilder.scala:8-28)
__global__ void bld_totals_kernel(int n_src, int n_runs, const int *__restrict__ first_run,
                                  const unsigned long long *__restrict__ run_count, const int *__restrict__ keep,
                                  unsigned long long *__restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int p = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (p >= n_src) return;
    const int s = first_run[p], e = p + 1 < n_src ? first_run[p + 1] : n_runs;
    unsigned long long t = 0;
    for (int r = s + lane; r < e; r += 32) t += keep[r] ? run_count[r] : 0ULL;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) total[p] = t;
}
__global__ void bld_edges_kernel(int n_runs, const int *__restrict__ keep, const int *__restrict__ kept_incl,
                                 const long long *__restrict__ run_src_row, const long long *__restrict__ run_dst,
                                 const unsigned long long *__restrict__ run_count, const long long *__restrict__ src_id,
                                 const unsigned long long *__restrict__ total, double beta, long long *__restrict__ out_src,
                                 long long *__restrict__ out_dst, double *__restrict__ out_w) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs || !keep[r]) return;
    const int at = kept_incl[r] - 1;
    const long long p = run_src_row[r];
    out_src[at] = src_id[p];
    out_dst[at] = run_dst[r];
    const double w = xdiv((double)run_count[r], (double)total[p]);   // visit_count / total_visit_count
    out_w[at] = xmul(w, beta);                                        // weight * beta
}

// device rows -> one balanced edge family on the device (COO sorted by (source, target)); *out_n edges
int family_edges(vrec_ctx *ctx, long long n, PoolBuf<long long> &d_src, PoolBuf<long long> &d_dst,
                 PoolBuf<long long> *d_weight, int top_n, double beta, PoolBuf<long long> &e_src,
                 PoolBuf<long long> &e_dst, PoolBuf<double> &e_w, int *out_n) {
    *out_n = 0;
    if (n == 0) return VREC_OK;
    cudaStream_t st = ctx->stream;
    FamilyCore c;
    VREC_TRY(family_core(ctx, n, d_src, d_dst, d_weight, top_n, c));
    PoolBuf<unsigned long long> total;
    VREC_TRY(total.alloc((size_t)c.n_src));
    VREC_TRY(e_src.alloc((size_t)std::max(1, c.nnz)));
    VREC_TRY(e_dst.alloc((size_t)std::max(1, c.nnz)));
    VREC_TRY(e_w.alloc((size_t)std::max(1, c.nnz)));
    bld_totals_kernel<<<(int)(((long long)c.n_src * 32 + 255) / 256), 256, 0, st>>>(c.n_src, c.n_runs, c.first_run.p,
                                                                                   c.run_count.p, c.keep.p, total.p);
    VREC_LAUNCHED(ctx);
    bld_edges_kernel<<<(c.n_runs + 255) / 256, 256, 0, st>>>(c.n_runs, c.keep.p, c.kept_incl.p, c.run_src_row.p,
                                                             c.run_dst.p, c.run_count.p, c.src_id.p, total.p, beta,
                                                             e_src.p, e_dst.p, e_w.p);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaStreamSynchronize(st));       // c's buffers are released on return
    *out_n = c.nnz;
    return VREC_OK;
}

// PlaceSimilarPlace.scala:18-42: the self-join of the visits on person_id, kept where the places differ and
// the timestamps are at most `interval_ms` apart; one output row (place, that_place) per pair of visits.
// Visits sorted by person; pair_cnt[i] = matching partners of visit i.
__global__ void bld_pair_count_kernel(long long n, const long long *__restrict__ person, const long long *__restrict__ place,
                                      const long long *__restrict__ ts, const int *__restrict__ seg_start_of,
                                      const int *__restrict__ seg_end_of, long long interval_ms, int *__restrict__ pair_cnt) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for (int j = seg_start_of[i]; j < seg_end_of[i]; ++j) {
        long long dt = ts[i] - ts[j];
        if (dt < 0) dt = -dt;
        c += (place[j] != place[i] && dt <= interval_ms) ? 1 : 0;
    }
    (void)person;
    pair_cnt[i] = c;
}
__global__ void bld_pair_emit_kernel(long long n, const long long *__restrict__ place, const long long *__restrict__ ts,
                                     const int *__restrict__ seg_start_of, const int *__restrict__ seg_end_of,
                                     long long interval_ms, const long long *__restrict__ pair_incl,
                                     const int *__restrict__ pair_cnt, long long *__restrict__ out_a,
                                     long long *__restrict__ out_b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long at = pair_incl[i] - pair_cnt[i];
    for (int j = seg_start_of[i]; j < seg_end_of[i]; ++j) {
        long long dt = ts[i] - ts[j];
        if (dt < 0) dt = -dt;
        if (place[j] != place[i] && dt <= interval_ms) {
            out_a[at] = place[i];
            out_b[at] = place[j];
            ++at;
        }
    }
}
// seg_start_of / seg_end_of for every row of a person-sorted visit list
__global__ void bld_segments_kernel(long long n, const long long *__restrict__ person, int *__restrict__ seg_start_of,
                                    int *__restrict__ seg_end_of) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // binary searches for the first / one-past-last row of person[i] (rows are sorted by person)
    const long long key = person[i];
    long long lo = 0, hi = i;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (person[mid] < key) lo = mid + 1; else hi = mid;
    }
    seg_start_of[i] = (int)lo;
    lo = i;
    hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (person[mid] <= key) lo = mid + 1; else hi = mid;
    }
    seg_end_of[i] = (int)lo;
}

int scan_inclusive_ll(vrec_ctx *ctx, const int *in, long long *out, long long n, PoolBuf<unsigned char> &tmp) {
    size_t bytes = 0;
    VREC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
    VREC_TRY(tmp.ensure(bytes));
    VREC_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, in, out, (int)n, ctx->stream));
    ctx->launches++;
    return VREC_OK;
}

}  // namespace

extern "C" int vrec_build_rating_vectors(vrec_ctx *ctx, int64_t n_rows, const int64_t *person_id,
                                         const int64_t *entity_id, const int64_t *weight, int32_t top_n,
                                         int64_t *out_n_persons, int64_t *out_nnz, int64_t *out_person_id,
                                         int64_t *out_rowptr, int32_t *out_col, double *out_val, int32_t *out_dim) {
    if (!ctx || n_rows < 0 || (n_rows > 0 && (!person_id || !entity_id)) || !out_n_persons || !out_nnz ||
        !out_person_id || !out_rowptr || !out_col || !out_val || !out_dim) {
        vrec_set_error("vrec_build_rating_vectors: NULL argument");
        return VREC_EINVAL;
    }
    if (top_n <= 0 || n_rows >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_build_rating_vectors: top_n must be positive and n_rows < 2^31");
        return VREC_EINVAL;
    }
    *out_n_persons = 0;
    *out_nnz = 0;
    *out_dim = 1;
    out_rowptr[0] = 0;
    if (n_rows == 0) return VREC_OK;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(pool_setup(ctx));
    cudaStream_t st = ctx->stream;
    const long long n = n_rows;
    PoolBuf<long long> d_person, d_entity, d_weight;
    VREC_TRY(d_person.upload((const long long *)person_id, (size_t)n, st));
    VREC_TRY(d_entity.upload((const long long *)entity_id, (size_t)n, st));
    if (weight) VREC_TRY(d_weight.upload((const long long *)weight, (size_t)n, st));
    FamilyCore c;
    VREC_TRY(family_core(ctx, n, d_person, d_entity, weight ? &d_weight : nullptr, top_n, c));
    const int n_persons = c.n_src, n_runs = c.n_runs, nnz = c.nnz;
    PoolBuf<unsigned long long> row_cnt;
    PoolBuf<long long> d_max;
    PoolBuf<int> d_col, d_bad;
    PoolBuf<double> d_val;
    VREC_TRY(row_cnt.alloc((size_t)n_persons));
    VREC_TRY(d_max.alloc(1));
    VREC_TRY(d_bad.alloc(1));
    VREC_TRY(d_col.alloc((size_t)std::max(1, nnz)));
    VREC_TRY(d_val.alloc((size_t)std::max(1, nnz)));
    VREC_CUDA(cudaMemsetAsync(row_cnt.p, 0, sizeof(unsigned long long) * (size_t)n_persons, st));
    VREC_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    const long long minus1 = -1;
    VREC_CUDA(cudaMemcpyAsync(d_max.p, &minus1, sizeof(long long), cudaMemcpyHostToDevice, st));
    bld_scatter_kernel<<<(n_runs + 255) / 256, 256, 0, st>>>(n_runs, c.keep.p, c.kept_incl.p, c.run_src_row.p, c.run_dst.p,
                                                             c.run_count.p, d_col.p, d_val.p, row_cnt.p, d_max.p, d_bad.p);
    VREC_LAUNCHED(ctx);
    std::vector<unsigned long long> h_cnt((size_t)n_persons);
    long long h_max = -1;
    int h_bad = 0;
    VREC_CUDA(cudaMemcpyAsync(h_cnt.data(), row_cnt.p, sizeof(unsigned long long) * (size_t)n_persons, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&h_bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_person_id, c.src_id.p, sizeof(long long) * (size_t)n_persons, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_col, d_col.p, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_val, d_val.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        vrec_set_error("Index out of Int range");     // ArithmeticException of knn/RatingVectorsBuilder.scala:40
        return VREC_EINVAL;
    }
    out_rowptr[0] = 0;
    for (int p = 0; p < n_persons; ++p) out_rowptr[p + 1] = out_rowptr[p] + (int64_t)h_cnt[p];
    *out_n_persons = n_persons;
    *out_nnz = nnz;
    *out_dim = (int32_t)(h_max + 1);               // :26-34 max(entity id) + 1
    return VREC_OK;
}

// One balanced edge family of the stochastic graph.
extern "C" int vrec_build_edge_family(vrec_ctx *ctx, int64_t n_rows, const int64_t *source_id, const int64_t *target_id,
                                      const int64_t *weight, int32_t top_n, double beta, int64_t capacity,
                                      int64_t *out_n, int64_t *out_source, int64_t *out_target, double *out_weight) {
    if (!ctx || n_rows < 0 || (n_rows > 0 && (!source_id || !target_id)) || !out_n || top_n <= 0 ||
        n_rows >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_build_edge_family: bad argument");
        return VREC_EINVAL;
    }
    *out_n = 0;
    if (n_rows == 0) return VREC_OK;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(pool_setup(ctx));
    cudaStream_t st = ctx->stream;
    PoolBuf<long long> d_src, d_dst, d_weight, e_src, e_dst;
    PoolBuf<double> e_w;
    VREC_TRY(d_src.upload((const long long *)source_id, (size_t)n_rows, st));
    VREC_TRY(d_dst.upload((const long long *)target_id, (size_t)n_rows, st));
    if (weight) VREC_TRY(d_weight.upload((const long long *)weight, (size_t)n_rows, st));
    int ne = 0;
    VREC_TRY(family_edges(ctx, n_rows, d_src, d_dst, weight ? &d_weight : nullptr, top_n, beta, e_src, e_dst, e_w, &ne));
    *out_n = ne;
    if (ne > capacity || !out_source || !out_target || !out_weight) {
        vrec_set_error("vrec_build_edge_family: %d edges, capacity %lld", ne, (long long)capacity);
        return VREC_ENOMEM;
    }
    VREC_CUDA(cudaMemcpyAsync(out_source, e_src.p, sizeof(long long) * (size_t)ne, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_target, e_dst.p, sizeof(long long) * (size_t)ne, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_weight, e_w.p, sizeof(double) * (size_t)ne, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    return VREC_OK;
}

// StochasticGraphBuilderMain.generateStochasticGraph (stochastic/StochasticGraphBuilderMain.scala:47-66): the four
// edge families in the reference's union order, each sorted by (source, target).
extern "C" int vrec_build_stochastic_graph(vrec_ctx *ctx, int64_t n_visits, const int64_t *person_id,
                                           const int64_t *place_id, const int64_t *category_id,
                                           const int64_t *timestamp_ms, double beta_person_place,
                                           double beta_person_category, int64_t capacity, int64_t *out_n,
                                           int64_t *out_source, int64_t *out_target, double *out_weight) {
    if (!ctx || n_visits < 0 || (n_visits > 0 && (!person_id || !place_id || !category_id || !timestamp_ms)) || !out_n ||
        n_visits >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_build_stochastic_graph: bad argument");
        return VREC_EINVAL;
    }
    *out_n = 0;
    if (n_visits == 0) return VREC_OK;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(pool_setup(ctx));
    cudaStream_t st = ctx->stream;
    const long long n = n_visits;
    const int grid = (int)((n + 255) / 256);
    PoolBuf<long long> d_person, d_place, d_cat, d_ts;
    VREC_TRY(d_person.upload((const long long *)person_id, (size_t)n, st));
    VREC_TRY(d_place.upload((const long long *)place_id, (size_t)n, st));
    VREC_TRY(d_cat.upload((const long long *)category_id, (size_t)n, st));
    VREC_TRY(d_ts.upload((const long long *)timestamp_ms, (size_t)n, st));
    PoolBuf<long long> fam_src[4], fam_dst[4];
    PoolBuf<double> fam_w[4];
    int fam_n[4] = {0, 0, 0, 0};
    {
        // PlaceSimilarPlace (top 50, beta 1): sort the visits by person, expand the pairs, then the common pipeline
        PoolBuf<long long> idx, s_person, s_idx, s_place, s_ts, pair_incl, pa, pb;
        PoolBuf<int> seg_s, seg_e, pair_cnt;
        PoolBuf<unsigned char> tmp;
        VREC_TRY(idx.alloc((size_t)n));
        VREC_TRY(s_person.alloc((size_t)n));
        VREC_TRY(s_idx.alloc((size_t)n));
        VREC_TRY(s_place.alloc((size_t)n));
        VREC_TRY(s_ts.alloc((size_t)n));
        VREC_TRY(seg_s.alloc((size_t)n));
        VREC_TRY(seg_e.alloc((size_t)n));
        VREC_TRY(pair_cnt.alloc((size_t)n));
        VREC_TRY(pair_incl.alloc((size_t)n));
        bld_iota_kernel<<<grid, 256, 0, st>>>(n, idx.p);
        VREC_LAUNCHED(ctx);
        size_t bytes = 0;
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_person.p, s_person.p, idx.p, s_idx.p, (int)n, 0, 64, st));
        VREC_TRY(tmp.ensure(bytes));
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_person.p, s_person.p, idx.p, s_idx.p, (int)n, 0, 64, st));
        ctx->launches++;
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, s_idx.p, d_place.p, s_place.p);
        VREC_LAUNCHED(ctx);
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, s_idx.p, d_ts.p, s_ts.p);
        VREC_LAUNCHED(ctx);
        bld_segments_kernel<<<grid, 256, 0, st>>>(n, s_person.p, seg_s.p, seg_e.p);
        VREC_LAUNCHED(ctx);
        const long long interval_ms = 7LL * 24 * 3600 * 1000;          // PlaceSimilarPlace.scala:13-14
        bld_pair_count_kernel<<<grid, 256, 0, st>>>(n, s_person.p, s_place.p, s_ts.p, seg_s.p, seg_e.p, interval_ms, pair_cnt.p);
        VREC_LAUNCHED(ctx);
        VREC_TRY(scan_inclusive_ll(ctx, pair_cnt.p, pair_incl.p, n, tmp));
        long long n_pairs = 0;
        VREC_CUDA(cudaMemcpyAsync(&n_pairs, pair_incl.p + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
        VREC_CUDA(cudaStreamSynchronize(st));
        if (n_pairs >= (long long)0x7fffffff) {
            vrec_set_error("vrec_build_stochastic_graph: %lld co-visit pairs exceed 2^31-1", n_pairs);
            return VREC_EINVAL;
        }
        if (n_pairs > 0) {
            VREC_TRY(pa.alloc((size_t)n_pairs));
            VREC_TRY(pb.alloc((size_t)n_pairs));
            bld_pair_emit_kernel<<<grid, 256, 0, st>>>(n, s_place.p, s_ts.p, seg_s.p, seg_e.p, interval_ms, pair_incl.p,
                                                       pair_cnt.p, pa.p, pb.p);
            VREC_LAUNCHED(ctx);
            VREC_TRY(family_edges(ctx, n_pairs, pa, pb, nullptr, 50, 1.0, fam_src[0], fam_dst[0], fam_w[0], &fam_n[0]));
        }
        VREC_CUDA(cudaStreamSynchronize(st));
    }
    // CategorySelectedPlace (top 100, beta 1), PersonLikesPlace (top 100), PersonLikesCategory (top 100)
    VREC_TRY(family_edges(ctx, n, d_cat, d_place, nullptr, 100, 1.0, fam_src[1], fam_dst[1], fam_w[1], &fam_n[1]));
    VREC_TRY(family_edges(ctx, n, d_person, d_place, nullptr, 100, beta_person_place, fam_src[2], fam_dst[2], fam_w[2], &fam_n[2]));
    VREC_TRY(family_edges(ctx, n, d_person, d_cat, nullptr, 100, beta_person_category, fam_src[3], fam_dst[3], fam_w[3], &fam_n[3]));
    const int64_t total = (int64_t)fam_n[0] + fam_n[1] + fam_n[2] + fam_n[3];
    *out_n = total;
    if (total > capacity || !out_source || !out_target || !out_weight) {
        vrec_set_error("vrec_build_stochastic_graph: %lld edges, capacity %lld", (long long)total, (long long)capacity);
        return VREC_ENOMEM;
    }
    int64_t at = 0;
    for (int f = 0; f < 4; ++f) {
        if (fam_n[f] == 0) continue;
        VREC_CUDA(cudaMemcpyAsync(out_source + at, fam_src[f].p, sizeof(long long) * (size_t)fam_n[f], cudaMemcpyDeviceToHost, st));
        VREC_CUDA(cudaMemcpyAsync(out_target + at, fam_dst[f].p, sizeof(long long) * (size_t)fam_n[f], cudaMemcpyDeviceToHost, st));
        VREC_CUDA(cudaMemcpyAsync(out_weight + at, fam_w[f].p, sizeof(double) * (size_t)fam_n[f], cudaMemcpyDeviceToHost, st));
        at += fam_n[f];
    }
    VREC_CUDA(cudaStreamSynchronize(st));
    return VREC_OK;
}

// ---------------------------------------------------------------------------------------
// PlaceVisits.calcPlaceVisits (PlaceVisits.scala:11-48) with the spatial grid its authors ask for (":30 TODO
// Very inefficient almost cross-join. Introduce a grid ..."): a location visit becomes a place visit of every
// place of its region within `accuracy` metres (haversine, Location.scala:30-43), for the visits of the last
// `last_days_count` days (counted from the latest timestamp, PlaceVisits.scala:50-61).
// Places are binned on the host into cells at least `accuracy` metres wide in both directions (per region, the
// longitude step widened by 1 / cos of the region's largest |latitude|), so a visit only looks at the 3 x 3
// cells around its own.  Distances are fp64 with CUDA's sin / cos / asin / sqrt: a pair whose distance is within
// a few ulps of `accuracy` can fall on the other side than with the reference's FastMath.
// ---------------------------------------------------------------------------------------
namespace {

struct PvGrid {
    int n_regions, n_cells;
    const long long *region_ids;          // ascending
    const double *lat0, *lon0, *cell_lon;  // per region
    double cell_lat;
    const unsigned long long *cell_key;    // ascending (region index << 42 | iy << 21 | ix)
    const int *cell_start;                 // [n_cells + 1] into the place arrays (sorted by cell, then place id)
    const long long *p_id, *p_cat;
    const double *p_lat, *p_lon;
};

__device__ __forceinline__ double pv_distance(double lat1d, double lon1d, double lat2d, double lon2d) {
    const double k = 0.017453292519943295;                 // toRadians
    const double lat1 = lat1d * k, lat2 = lat2d * k, lon1 = lon1d * k, lon2 = lon2d * k;
    const double s1 = sin((lat2 - lat1) / 2), s2 = sin((lon2 - lon1) / 2);
    const double hav = s1 * s1 + cos(lat1) * cos(lat2) * (s2 * s2);
    return 6371000.0 * 2 * asin(sqrt(hav));
}

template <bool EMIT>
__global__ void pv_match_kernel(long long n, const long long *__restrict__ person, const double *__restrict__ lat,
                                const double *__restrict__ lon, const long long *__restrict__ ts,
                                const long long *__restrict__ region, long long ts_from, double accuracy, PvGrid g,
                                int *__restrict__ cnt, const long long *__restrict__ incl, long long *__restrict__ o_person,
                                long long *__restrict__ o_ts, long long *__restrict__ o_place,
                                long long *__restrict__ o_region, long long *__restrict__ o_cat) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    long long at = EMIT ? incl[i] - cnt[i] : 0;
    const long long at0 = at;
    if (ts[i] >= ts_from) {
        int lo = 0, hi = g.n_regions;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (g.region_ids[mid] < region[i]) lo = mid + 1; else hi = mid;
        }
        if (lo < g.n_regions && g.region_ids[lo] == region[i]) {
            const int r = lo;
            const long long iy = (long long)floor((lat[i] - g.lat0[r]) / g.cell_lat);
            const long long ix = (long long)floor((lon[i] - g.lon0[r]) / g.cell_lon[r]);
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const long long y = iy + dy, x = ix + dx;
                    if (y < 0 || x < 0 || y >= (1 << 21) || x >= (1 << 21)) continue;
                    const unsigned long long key = ((unsigned long long)r << 42) | ((unsigned long long)y << 21) | (unsigned long long)x;
                    int l2 = 0, h2 = g.n_cells;
                    while (l2 < h2) {
                        int mid = (l2 + h2) >> 1;
                        if (g.cell_key[mid] < key) l2 = mid + 1; else h2 = mid;
                    }
                    if (l2 >= g.n_cells || g.cell_key[l2] != key) continue;
                    for (int p = g.cell_start[l2]; p < g.cell_start[l2 + 1]; ++p) {
                        if (pv_distance(lat[i], lon[i], g.p_lat[p], g.p_lon[p]) <= accuracy) {   // PlaceVisits.scala:22
                            if (EMIT) {
                                // keep this visit's rows in ascending place id (insertion into the short run)
                                long long q = at;
                                while (q > at0 && o_place[q - 1] > g.p_id[p]) {
                                    o_place[q] = o_place[q - 1];
                                    o_cat[q] = o_cat[q - 1];
                                    --q;
                                }
                                o_place[q] = g.p_id[p];
                                o_cat[q] = g.p_cat[p];
                                o_person[at] = person[i];
                                o_ts[at] = ts[i];
                                o_region[at] = region[i];
                                ++at;
                            }
                            ++c;
                        }
                    }
                }
        }
    }
    if (!EMIT) cnt[i] = c;
}

}  // namespace

extern "C" int vrec_build_place_visits(vrec_ctx *ctx, int64_t n_visits, const int64_t *person_id, const double *latitude,
                                       const double *longitude, const int64_t *timestamp_ms, const int64_t *region_id,
                                       int64_t n_places, const int64_t *place_id, const double *place_latitude,
                                       const double *place_longitude, const int64_t *place_category,
                                       const int64_t *place_region, int32_t last_days_count, double accuracy_m,
                                       int64_t capacity, int64_t *out_n, int64_t *out_person, int64_t *out_timestamp_ms,
                                       int64_t *out_place, int64_t *out_region, int64_t *out_category) {
    if (!ctx || !out_n || n_visits < 0 || n_places < 0 || n_visits >= (int64_t)0x7fffffff || n_places >= (int64_t)0x7fffffff ||
        (n_visits > 0 && (!person_id || !latitude || !longitude || !timestamp_ms || !region_id)) ||
        (n_places > 0 && (!place_id || !place_latitude || !place_longitude || !place_category || !place_region)) ||
        !(accuracy_m > 0) || last_days_count < 0) {
        vrec_set_error("vrec_build_place_visits: bad argument");
        return VREC_EINVAL;
    }
    *out_n = 0;
    if (n_visits == 0 || n_places == 0) return VREC_OK;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(pool_setup(ctx));
    cudaStream_t st = ctx->stream;
    // ---- host: the grid over the places
    std::vector<long long> regions(place_region, place_region + n_places);
    std::sort(regions.begin(), regions.end());
    regions.erase(std::unique(regions.begin(), regions.end()), regions.end());
    const int nr = (int)regions.size();
    std::vector<double> lat0((size_t)nr, 1e300), lon0((size_t)nr, 1e300), amax((size_t)nr, 0.0), cell_lon((size_t)nr);
    std::vector<int> preg((size_t)n_places);
    for (int64_t p = 0; p < n_places; ++p) {
        const int r = (int)(std::lower_bound(regions.begin(), regions.end(), (long long)place_region[p]) - regions.begin());
        preg[p] = r;
        lat0[r] = std::min(lat0[r], place_latitude[p]);
        lon0[r] = std::min(lon0[r], place_longitude[p]);
        amax[r] = std::max(amax[r], std::fabs(place_latitude[p]));
    }
    const double deg_per_m = 180.0 / (3.141592653589793 * 6371000.0);
    const double cell_lat = accuracy_m * deg_per_m * 1.001;
    for (int r = 0; r < nr; ++r) {
        // a visit can be `accuracy` metres north / south of the region's outermost place
        const double a = std::min(89.9, amax[r] + 2 * cell_lat) * 3.141592653589793 / 180.0;
        cell_lon[r] = cell_lat / std::max(1e-3, std::cos(a));
        lat0[r] -= cell_lat;                              // one cell of margin: visit cells are never negative by much
        lon0[r] -= cell_lon[r];
    }
    std::vector<unsigned long long> pkey((size_t)n_places);
    std::vector<int> order((size_t)n_places);
    bool grid_ok = true;
    for (int64_t p = 0; p < n_places; ++p) {
        const int r = preg[p];
        const long long iy = (long long)std::floor((place_latitude[p] - lat0[r]) / cell_lat);
        const long long ix = (long long)std::floor((place_longitude[p] - lon0[r]) / cell_lon[r]);
        if (iy < 0 || ix < 0 || iy >= (1 << 21) || ix >= (1 << 21) || nr >= (1 << 21)) grid_ok = false;
        pkey[p] = ((unsigned long long)r << 42) | ((unsigned long long)iy << 21) | (unsigned long long)ix;
        order[p] = (int)p;
    }
    if (!grid_ok) {
        vrec_set_error("vrec_build_place_visits: a region spans more than 2^21 grid cells");
        return VREC_EINVAL;
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return pkey[a] < pkey[b] || (pkey[a] == pkey[b] && place_id[a] < place_id[b]);
    });
    std::vector<unsigned long long> cell_key;
    std::vector<int> cell_start;
    std::vector<long long> s_id((size_t)n_places), s_cat((size_t)n_places);
    std::vector<double> s_lat((size_t)n_places), s_lon((size_t)n_places);
    for (int64_t k = 0; k < n_places; ++k) {
        const int p = order[k];
        if (k == 0 || pkey[p] != cell_key.back()) {
            cell_key.push_back(pkey[p]);
            cell_start.push_back((int)k);
        }
        s_id[k] = place_id[p];
        s_cat[k] = place_category[p];
        s_lat[k] = place_latitude[p];
        s_lon[k] = place_longitude[p];
    }
    cell_start.push_back((int)n_places);
    long long ts_max = timestamp_ms[0];
    for (int64_t i = 1; i < n_visits; ++i) ts_max = std::max<long long>(ts_max, timestamp_ms[i]);
    const long long ts_from = ts_max - (long long)last_days_count * 86400000LL;     // PlaceVisits.scala:50-61 (UTC days)
    // ---- device
    PoolBuf<long long> d_person, d_ts, d_region, d_rid, d_pid, d_pcat, d_incl;
    PoolBuf<double> d_lat, d_lon, d_lat0, d_lon0, d_clon, d_plat, d_plon;
    PoolBuf<unsigned long long> d_ckey;
    PoolBuf<int> d_cstart, d_cnt;
    PoolBuf<unsigned char> tmp;
    const long long n = n_visits;
    VREC_TRY(d_person.upload((const long long *)person_id, (size_t)n, st));
    VREC_TRY(d_ts.upload((const long long *)timestamp_ms, (size_t)n, st));
    VREC_TRY(d_region.upload((const long long *)region_id, (size_t)n, st));
    VREC_TRY(d_lat.upload(latitude, (size_t)n, st));
    VREC_TRY(d_lon.upload(longitude, (size_t)n, st));
    VREC_TRY(d_rid.upload(regions.data(), regions.size(), st));
    VREC_TRY(d_lat0.upload(lat0.data(), lat0.size(), st));
    VREC_TRY(d_lon0.upload(lon0.data(), lon0.size(), st));
    VREC_TRY(d_clon.upload(cell_lon.data(), cell_lon.size(), st));
    VREC_TRY(d_ckey.upload(cell_key.data(), cell_key.size(), st));
    VREC_TRY(d_cstart.upload(cell_start.data(), cell_start.size(), st));
    VREC_TRY(d_pid.upload(s_id.data(), s_id.size(), st));
    VREC_TRY(d_pcat.upload(s_cat.data(), s_cat.size(), st));
    VREC_TRY(d_plat.upload(s_lat.data(), s_lat.size(), st));
    VREC_TRY(d_plon.upload(s_lon.data(), s_lon.size(), st));
    VREC_TRY(d_cnt.alloc((size_t)n));
    VREC_TRY(d_incl.alloc((size_t)n));
    PvGrid g;
    g.n_regions = nr;
    g.n_cells = (int)cell_key.size();
    g.region_ids = d_rid.p;
    g.lat0 = d_lat0.p;
    g.lon0 = d_lon0.p;
    g.cell_lon = d_clon.p;
    g.cell_lat = cell_lat;
    g.cell_key = d_ckey.p;
    g.cell_start = d_cstart.p;
    g.p_id = d_pid.p;
    g.p_cat = d_pcat.p;
    g.p_lat = d_plat.p;
    g.p_lon = d_plon.p;
    const int grid = (int)((n + 127) / 128);
    pv_match_kernel<false><<<grid, 128, 0, st>>>(n, d_person.p, d_lat.p, d_lon.p, d_ts.p, d_region.p, ts_from, accuracy_m, g,
                                                d_cnt.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    VREC_LAUNCHED(ctx);
    VREC_TRY(scan_inclusive_ll(ctx, d_cnt.p, d_incl.p, n, tmp));
    long long total = 0;
    VREC_CUDA(cudaMemcpyAsync(&total, d_incl.p + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    *out_n = total;
    if (total > capacity || !out_person || !out_timestamp_ms || !out_place || !out_region || !out_category) {
        vrec_set_error("vrec_build_place_visits: %lld place visits, capacity %lld", total, (long long)capacity);
        return VREC_ENOMEM;
    }
    if (total == 0) return VREC_OK;
    PoolBuf<long long> o_person, o_ts, o_place, o_region, o_cat;
    VREC_TRY(o_person.alloc((size_t)total));
    VREC_TRY(o_ts.alloc((size_t)total));
    VREC_TRY(o_place.alloc((size_t)total));
    VREC_TRY(o_region.alloc((size_t)total));
    VREC_TRY(o_cat.alloc((size_t)total));
    pv_match_kernel<true><<<grid, 128, 0, st>>>(n, d_person.p, d_lat.p, d_lon.p, d_ts.p, d_region.p, ts_from, accuracy_m, g,
                                               d_cnt.p, d_incl.p, o_person.p, o_ts.p, o_place.p, o_region.p, o_cat.p);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaMemcpyAsync(out_person, o_person.p, sizeof(long long) * (size_t)total, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_timestamp_ms, o_ts.p, sizeof(long long) * (size_t)total, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_place, o_place.p, sizeof(long long) * (size_t)total, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_region, o_region.p, sizeof(long long) * (size_t)total, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_category, o_cat.p, sizeof(long long) * (size_t)total, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    return VREC_OK;
}
