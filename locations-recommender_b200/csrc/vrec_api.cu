// Context, error reporting and the generic ranked top-N selection kernel.
#include <stdarg.h>

#include "vrec_internal.cuh"

static thread_local std::string g_last_error;

void vrec_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" const char *vrec_last_error(void) { return g_last_error.c_str(); }
extern "C" int vrec_abi_version(void) { return VREC_ABI_VERSION; }

extern "C" int vrec_init(int device, vrec_ctx **out) {
    if (!out) {
        vrec_set_error("vrec_init: out is NULL");
        return VREC_EINVAL;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        vrec_set_error("vrec_init: no CUDA device (%s); libvrec has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return VREC_ENODEV;
    }
    if (device < 0) VREC_CUDA(cudaGetDevice(&device));
    if (device >= count) {
        vrec_set_error("vrec_init: device %d out of range (%d devices)", device, count);
        return VREC_ENODEV;
    }
    cudaDeviceProp prop;
    VREC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        vrec_set_error("vrec_init: device %d is sm_%d%d; libvrec is built for sm_100a only", device,
                       prop.major, prop.minor);
        return VREC_ENODEV;
    }
    VREC_CUDA(cudaSetDevice(device));
    vrec_ctx *ctx = new vrec_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        vrec_set_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
        delete ctx;
        return VREC_ECUDA;
    }
    *out = ctx;
    return VREC_OK;
}

extern "C" void vrec_shutdown(vrec_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        vrec_comm_destroy(ctx);
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

extern "C" void *vrec_stream(vrec_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t vrec_launch_count(vrec_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int vrec_synchronize(vrec_ctx *ctx) {
    if (!ctx) return VREC_EINVAL;
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    return VREC_OK;
}

// ---------------------------------------------------------------------------------------
// Ranked top-N: out = the first max_recs candidates in the order (value desc, key asc).
// One block per query; round r picks the best candidate strictly after pick r-1, so equal
// (value, key) pairs collapse to one row.  A candidate is i with ok[i] != 0 (if given) and a
// non-NaN value.  Replaces `orderBy(col desc).limit(N)` of knn/KnnRecommenderMain.scala:99-100
// and stochastic/StochasticRecommenderMain.scala:72-73.
// ---------------------------------------------------------------------------------------
constexpr int SELECT_THREADS = 1024;

__global__ void __launch_bounds__(SELECT_THREADS)
select_topn_kernel(const double *__restrict__ val, const long long *__restrict__ key,
                   const unsigned char *__restrict__ ok, long long n, long long stride_val,
                   int max_recs, long long *__restrict__ out_key, double *__restrict__ out_val,
                   int *__restrict__ out_count) {
    const int q = blockIdx.x;
    const double *v = val + (size_t)q * (size_t)stride_val;
    __shared__ double s_val[SELECT_THREADS / 32];
    __shared__ long long s_key[SELECT_THREADS / 32];
    __shared__ int s_has[SELECT_THREADS / 32];
    __shared__ double pick_val;
    __shared__ long long pick_key;
    __shared__ int pick_has;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool have_last = false;
    double last_val = 0.0;
    long long last_key = 0;
    int count = 0;
    for (int r = 0; r < max_recs; ++r) {
        bool has = false;
        double bv = 0.0;
        long long bk = 0;
        for (long long i = threadIdx.x; i < n; i += SELECT_THREADS) {
            if (ok && !ok[i]) continue;
            double x = v[i];
            if (x != x) continue;
            long long k = key ? key[i] : i;
            if (have_last && !ranks_before(last_val, last_key, x, k)) continue;
            if (!has || ranks_before(x, k, bv, bk)) {
                has = true;
                bv = x;
                bk = k;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            int oh = __shfl_xor_sync(0xffffffffu, (int)has, off);
            double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            long long okk = __shfl_xor_sync(0xffffffffu, bk, off);
            if (oh && (!has || ranks_before(ov, okk, bv, bk))) {
                has = true;
                bv = ov;
                bk = okk;
            }
        }
        if (lane == 0) {
            s_has[warp] = has;
            s_val[warp] = bv;
            s_key[warp] = bk;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            bool h = false;
            double fv = 0.0;
            long long fk = 0;
            for (int w = 0; w < SELECT_THREADS / 32; ++w)
                if (s_has[w] && (!h || ranks_before(s_val[w], s_key[w], fv, fk))) {
                    h = true;
                    fv = s_val[w];
                    fk = s_key[w];
                }
            pick_has = h;
            pick_val = fv;
            pick_key = fk;
        }
        __syncthreads();
        if (!pick_has) break;
        have_last = true;
        last_val = pick_val;
        last_key = pick_key;
        if (threadIdx.x == 0) {
            out_key[(size_t)q * max_recs + r] = last_key;
            out_val[(size_t)q * max_recs + r] = last_val;
        }
        count = r + 1;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_count[q] = count;
}

int vrec_launch_select_topn(vrec_ctx *ctx, const double *d_val, const long long *d_key,
                            const unsigned char *d_ok, long long n, long long stride_val, int n_queries,
                            int max_recs, long long *d_out_key, double *d_out_val, int *d_out_count) {
    if (n_queries <= 0) return VREC_OK;
    select_topn_kernel<<<n_queries, SELECT_THREADS, 0, ctx->stream>>>(
        d_val, d_key, d_ok, n, stride_val, max_recs, d_out_key, d_out_val, d_out_count);
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}
