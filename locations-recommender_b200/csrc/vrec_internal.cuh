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
};

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

// (value desc, key asc) total order used for every ranked output.
__device__ __forceinline__ bool ranks_before(double va, long long ka, double vb, long long kb) {
    return va > vb || (va == vb && ka < kb);
}
#endif

// generic top-N selection (vrec_select.cu)
int vrec_launch_select_topn(vrec_ctx *ctx, const double *d_val, const long long *d_key,
                            const unsigned char *d_ok, long long n, long long stride_val, int n_queries,
                            int max_recs, long long *d_out_key, double *d_out_val, int *d_out_count);
