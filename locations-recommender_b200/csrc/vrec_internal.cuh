// Shared internals of libvrec.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "vrec.h"

// Canonical summation order shared with the oracle (oracle/vrec_oracle.c, `warp_sum`):
// 32 lane-strided partial sums, xor-butterfly 1,2,4,8,16, segmented every 1024 terms.
#define VREC_CANON_SEG 1024

void vrec_set_error(const char *fmt, ...);

struct vrec_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // multi-GPU (vrec_comm.cu): NCCL communicator of the one-process-per-GPU job
    void *comm = nullptr;
    int rank = 0, world = 1;
    // opt-ins to large dynamic shared memory are per device, so they are remembered per context (a process may
    // hold contexts on several devices), not in function-local statics
    size_t attr_knn_post = 0;
    bool attr_knn_tc = false, attr_knn_ws = false, attr_knn_tile = false;
};

void vrec_comm_destroy(vrec_ctx *ctx);
int vrec_comm_allgather_bytes(vrec_ctx *ctx, const void *send, void *recv, size_t bytes);

#define VREC_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            vrec_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return VREC_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

#define VREC_TRY(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != VREC_OK) return rc__; \
    } while (0)

// Counts the launch and checks for launch errors.
#define VREC_LAUNCHED(ctx)                                                                \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            vrec_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return VREC_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

// Owning device buffer.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            vrec_set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
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
    size_t bytes() const { return n * sizeof(T); }
};

#ifdef __CUDACC__
// ---- exact (non-contracted) fp64 helpers: the JVM never fuses a*b+c ----
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double canon_butterfly(double v) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) v = xadd(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// ---- L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint) ----
// Streamed-once data (CSR columns / weights) is marked evict_first so that it does not push the
// gathered vector (marked evict_last) out of the 126 MB L2.
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// `fraction` of the addresses (chosen by an address hash, so the same subset every time) get
// evict_last; the rest keep the default priority.  Lets a vector somewhat larger than the L2
// keep a stable resident subset instead of thrashing as a whole.
__device__ __forceinline__ unsigned long long policy_evict_last(float fraction = 1.0f) {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(p) : "f"(fraction));
    return p;
}
__device__ __forceinline__ int ld_stream_i32(const int *p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_keep_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_f64(double *p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// (value desc, key asc) total order used for every ranked output.
__device__ __forceinline__ bool ranks_before(double va, long long ka, double vb, long long kb) {
    return va > vb || (va == vb && ka < kb);
}
#endif

// generic top-N selection (vrec_select.cu)
int vrec_launch_select_topn(vrec_ctx *ctx, const double *d_val, const long long *d_key,
                            const unsigned char *d_ok, long long n, long long stride_val, int n_queries,
                            int max_recs, long long *d_out_key, double *d_out_val, int *d_out_count);
