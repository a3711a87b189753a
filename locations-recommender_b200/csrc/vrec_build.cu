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
    const int grid = (int)((n + 255) / 256);
    PoolBuf<long long> d_person, d_entity, d_idx, k_a, k_b, v_a, v_b, d_weight, d_wsorted;
    PoolBuf<unsigned char> tmp;
    VREC_TRY(d_person.upload((const long long *)person_id, (size_t)n, st));
    VREC_TRY(d_entity.upload((const long long *)entity_id, (size_t)n, st));
    VREC_TRY(k_a.alloc((size_t)n));
    VREC_TRY(k_b.alloc((size_t)n));
    VREC_TRY(v_a.alloc((size_t)n));
    VREC_TRY(v_b.alloc((size_t)n));
    // stable LSD: by entity (values = person), then by person (values = entity); a third pass carries the
    // weights through the same permutation when there are any
    size_t bytes = 0;
    VREC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_entity.p, k_a.p, d_person.p, v_a.p, (int)n, 0, 64, st));
    VREC_TRY(tmp.ensure(bytes));
    if (weight) {
        // sort (entity, row index) first so that the weights can follow: values = original row
        VREC_TRY(d_idx.alloc((size_t)n));
        VREC_TRY(d_weight.upload((const long long *)weight, (size_t)n, st));
        VREC_TRY(d_wsorted.alloc((size_t)n));
        bld_iota_kernel<<<grid, 256, 0, st>>>(n, d_idx.p);
        VREC_LAUNCHED(ctx);
        // pass 1: by entity, values = row index
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_entity.p, k_a.p, d_idx.p, v_a.p, (int)n, 0, 64, st));
        // gather person by the permutation, then pass 2: by person, values = row index
        PoolBuf<long long> pg;
        VREC_TRY(pg.alloc((size_t)n));
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_a.p, d_person.p, pg.p);
        VREC_LAUNCHED(ctx);
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, pg.p, k_b.p, v_a.p, v_b.p, (int)n, 0, 64, st));
        // k_b = persons sorted; v_b = original rows in (person, entity) order
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_b.p, d_entity.p, k_a.p);
        VREC_LAUNCHED(ctx);
        bld_gather_kernel<<<grid, 256, 0, st>>>(n, v_b.p, d_weight.p, d_wsorted.p);
        VREC_LAUNCHED(ctx);
        VREC_CUDA(cudaStreamSynchronize(st));      // pg goes out of scope
        ctx->launches += 2;
    } else {
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_entity.p, k_a.p, d_person.p, v_a.p, (int)n, 0, 64, st));
        // k_a = entity sorted, v_a = person in that order; pass 2: by person, values = entity
        VREC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, v_a.p, k_b.p, k_a.p, v_b.p, (int)n, 0, 64, st));
        // k_b = persons sorted, v_b = entities in (person, entity) order
        VREC_CUDA(cudaMemcpyAsync(k_a.p, v_b.p, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        ctx->launches += 2;
    }
    const long long *s_person = k_b.p, *s_entity = k_a.p;
    // run heads, run / person numbering
    PoolBuf<int> head, phead, run_of, prow_of;
    VREC_TRY(head.alloc((size_t)n));
    VREC_TRY(phead.alloc((size_t)n));
    VREC_TRY(run_of.alloc((size_t)n));
    VREC_TRY(prow_of.alloc((size_t)n));
    bld_heads_kernel<<<grid, 256, 0, st>>>(n, s_person, s_entity, head.p, phead.p);
    VREC_LAUNCHED(ctx);
    VREC_TRY(scan_inclusive<int>(ctx, head.p, run_of.p, n, tmp));
    VREC_TRY(scan_inclusive<int>(ctx, phead.p, prow_of.p, n, tmp));
    int n_runs = 0, n_persons = 0;
    VREC_CUDA(cudaMemcpyAsync(&n_runs, run_of.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&n_persons, prow_of.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    PoolBuf<long long> run_person_row, run_entity, d_pid, d_max;
    PoolBuf<unsigned long long> run_count, row_cnt;
    PoolBuf<int> first_run, keep, kept_incl, d_col, d_bad;
    PoolBuf<double> d_val;
    VREC_TRY(run_person_row.alloc((size_t)n_runs));
    VREC_TRY(run_entity.alloc((size_t)n_runs));
    VREC_TRY(run_count.alloc((size_t)n_runs));
    VREC_TRY(d_pid.alloc((size_t)n_persons));
    VREC_TRY(first_run.alloc((size_t)n_persons));
    VREC_TRY(keep.alloc((size_t)n_runs));
    VREC_TRY(kept_incl.alloc((size_t)n_runs));
    VREC_TRY(row_cnt.alloc((size_t)n_persons));
    VREC_TRY(d_max.alloc(1));
    VREC_TRY(d_bad.alloc(1));
    VREC_CUDA(cudaMemsetAsync(run_count.p, 0, sizeof(unsigned long long) * (size_t)n_runs, st));
    VREC_CUDA(cudaMemsetAsync(row_cnt.p, 0, sizeof(unsigned long long) * (size_t)n_persons, st));
    VREC_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    const long long minus1 = -1;
    VREC_CUDA(cudaMemcpyAsync(d_max.p, &minus1, sizeof(long long), cudaMemcpyHostToDevice, st));
    bld_runs_kernel<<<grid, 256, 0, st>>>(n, s_person, s_entity, weight ? d_wsorted.p : nullptr, head.p, run_of.p,
                                          prow_of.p, run_person_row.p, run_entity.p, run_count.p, d_pid.p, first_run.p);
    VREC_LAUNCHED(ctx);
    bld_rank_kernel<<<(int)(((long long)n_persons * 32 + 255) / 256), 256, 0, st>>>(n_persons, n_runs, first_run.p,
                                                                                    run_count.p, top_n, keep.p);
    VREC_LAUNCHED(ctx);
    VREC_TRY(scan_inclusive<int>(ctx, keep.p, kept_incl.p, n_runs, tmp));
    int nnz = 0;
    VREC_CUDA(cudaMemcpyAsync(&nnz, kept_incl.p + (n_runs - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaStreamSynchronize(st));
    VREC_TRY(d_col.alloc((size_t)std::max(1, nnz)));
    VREC_TRY(d_val.alloc((size_t)std::max(1, nnz)));
    bld_scatter_kernel<<<(n_runs + 255) / 256, 256, 0, st>>>(n_runs, keep.p, kept_incl.p, run_person_row.p, run_entity.p,
                                                             run_count.p, d_col.p, d_val.p, row_cnt.p, d_max.p, d_bad.p);
    VREC_LAUNCHED(ctx);
    std::vector<unsigned long long> h_cnt((size_t)n_persons);
    long long h_max = -1;
    int h_bad = 0;
    VREC_CUDA(cudaMemcpyAsync(h_cnt.data(), row_cnt.p, sizeof(unsigned long long) * (size_t)n_persons, cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(&h_bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VREC_CUDA(cudaMemcpyAsync(out_person_id, d_pid.p, sizeof(long long) * (size_t)n_persons, cudaMemcpyDeviceToHost, st));
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
