// Self-test of the tcgen05 path: one 128 x 128 x D fp16 tile through tcgen05.mma + TMEM against a
// scalar reference.  Validates the shared-memory / instruction descriptors of vrec_tc.cuh.
#include "vrec_internal.cuh"
#include "vrec_tc.cuh"

namespace {

constexpr int ST_ROWS = 128;
constexpr int ST_D = 128;

__global__ void __launch_bounds__(128) tc_selftest_kernel(const __half *__restrict__ A, const __half *__restrict__ B,
                                                           float *__restrict__ C, int swizzled) {
    extern __shared__ unsigned char smem_raw[];
    // operand tiles must start on a core-matrix (128 B) boundary: the low bits of the descriptor
    // start address are ignored by the hardware.  Align the dynamic buffer by hand (1 KB).
    unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = smem, *sB = smem + ST_ROWS * ST_D * 2;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init_fence();
    }
    __syncthreads();
    if (swizzled == 2) {
        // A operand in tensor memory: row = lane, 32-bit column j = halves (2j, 2j+1) of the row, at columns 128..191
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + 128;
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            for (int j = 0; j < 32; ++j)
                r[j] = *reinterpret_cast<const uint32_t *>(A + (size_t)tid * ST_D + (h * 32 + j) * 2);
            tc::tmem_st32(ta + h * 32, r);
        }
        tc::tmem_wait_st();
    }
    // row-major global (row r: D halves) -> chunk-major smem
    for (int q = tid; q < ST_ROWS * (ST_D / 8); q += blockDim.x) {
        int r = q % ST_ROWS, c = q / ST_ROWS;
        uint4 va = *reinterpret_cast<const uint4 *>(A + (size_t)r * ST_D + c * 8);
        uint4 vb = *reinterpret_cast<const uint4 *>(B + (size_t)r * ST_D + c * 8);
        uint32_t off = swizzled ? tc::sw128_offset(ST_ROWS, r, c) : (uint32_t)(c * (ST_ROWS * 16) + r * 16);
        if (swizzled == 2) va = make_uint4(0, 0, 0, 0);      // the smem copy of A must not be what is used
        *reinterpret_cast<uint4 *>(sA + off) = va;
        *reinterpret_cast<uint4 *>(sB + off) = vb;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(128, 128);
        for (int k = 0; k < ST_D / 16; ++k) {
            uint64_t da, db;
            if (swizzled) {
                da = tc::make_desc_sw128(tc::sw128_kstep_addr(tc::smem_u32(sA), ST_ROWS, k));
                db = tc::make_desc_sw128(tc::sw128_kstep_addr(tc::smem_u32(sB), ST_ROWS, k));
            } else {
                da = tc::make_desc(tc::smem_u32(sA) + k * 2 * (ST_ROWS * 16), ST_ROWS * 16, 128);
                db = tc::make_desc(tc::smem_u32(sB) + k * 2 * (ST_ROWS * 16), ST_ROWS * 16, 128);
            }
            if (swizzled == 2) tc::mma_f16_ts(tbase, tbase + 128 + k * 8, db, idesc, k > 0);
            else tc::mma_f16(tbase, da, db, idesc, k > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = tid;                       // TMEM lane = row of A
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) C[(size_t)row * 128 + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

// Micro-benchmarks behind tools/tc_stream.py (MMA issue rate, L2 -> shared-memory streaming): not part of the
// shipped library; build with -DVREC_WITH_MICROBENCH=1 (make CUDAFLAGS_EXTRA=-DVREC_WITH_MICROBENCH=1) to use them.
#ifndef VREC_WITH_MICROBENCH
#define VREC_WITH_MICROBENCH 0
#endif
#if VREC_WITH_MICROBENCH
// issue `reps` chains of 8 MMAs (128x128x128) back to back and time them with clock64
__global__ void __launch_bounds__(128) tc_rate_kernel(int swizzled, int reps, long long *out_cycles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = smem, *sB = smem + ST_ROWS * ST_D * 2;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init_fence();
    }
    for (int q = tid; q < 2 * ST_ROWS * ST_D * 2 / 16; q += blockDim.x) reinterpret_cast<uint4 *>(smem)[q] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(128, 128);
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            for (int k = 0; k < ST_D / 16; ++k) {
                uint64_t da, db;
                if (swizzled) {
                    da = tc::make_desc_sw128(tc::sw128_kstep_addr(tc::smem_u32(sA), ST_ROWS, k));
                    db = tc::make_desc_sw128(tc::sw128_kstep_addr(tc::smem_u32(sB), ST_ROWS, k));
                } else {
                    da = tc::make_desc(tc::smem_u32(sA) + k * 2 * (ST_ROWS * 16), ST_ROWS * 16, 128);
                    db = tc::make_desc(tc::smem_u32(sB) + k * 2 * (ST_ROWS * 16), ST_ROWS * 16, 128);
                }
                tc::mma_f16(tbase + (rep & 1) * 128, da, db, idesc, k > 0);
            }
        }
        long long t1 = clock64();
        tc::mma_commit(&bar);
        out_cycles[0] = t1 - t0;
        tc::mbar_wait(&bar, 0);
        out_cycles[1] = clock64() - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

// every CTA streams `ntiles` 32 KB tiles (start staggered by `stagger` tiles per CTA) through a ring of
// `stages` bulk copies and does nothing with them: the L2 -> shared-memory bandwidth the KNN filter can get
__global__ void __launch_bounds__(128) tc_stream_kernel(const unsigned char *__restrict__ src, int ntiles, int stages,
                                                         int stagger) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int s_ = 0; s_ < stages; ++s_) tc::mbar_init(&full[s_], 1);
        tc::mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int rot = (int)(((long long)blockIdx.x * stagger) % ntiles);
        auto tile = [&](int i) { int w = i + rot; return w >= ntiles ? w - ntiles : w; };
        for (int i = 0; i < stages && i < ntiles; ++i) {
            tc::mbar_expect_tx(&full[i], 32768);
            tc::bulk_copy_g2s(smem + (size_t)i * 32768, src + (size_t)tile(i) * 32768, 32768, &full[i]);
        }
        for (int i = 0; i < ntiles; ++i) {
            const int st = i % stages;
            tc::mbar_wait(&full[st], (uint32_t)((i / stages) & 1));
            if (i + stages < ntiles) {
                tc::mbar_expect_tx(&full[st], 32768);
                tc::bulk_copy_g2s(smem + (size_t)st * 32768, src + (size_t)tile(i + stages) * 32768, 32768, &full[st]);
            }
        }
    }
    __syncthreads();
}
#endif   // VREC_WITH_MICROBENCH

}  // namespace

#if VREC_WITH_MICROBENCH
// Debug: milliseconds for `ctas` CTAs to stream ntiles x 32 KB each from a common buffer through `stages`
// bulk copies in flight (start points `stagger` tiles apart).
extern "C" int vrec_debug_tc_stream(vrec_ctx *ctx, int ctas, int ntiles, int stages, int stagger, double *out_ms) {
    if (!ctx || !out_ms || ctas <= 0 || ntiles <= 0 || stages <= 0 || stages > 6) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<unsigned char> buf;
    VREC_TRY(buf.alloc((size_t)ntiles * 32768));
    VREC_CUDA(cudaMemsetAsync(buf.p, 0, (size_t)ntiles * 32768, ctx->stream));
    size_t smem = (size_t)stages * 32768 + 1024;
    VREC_CUDA(cudaFuncSetAttribute(tc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    VREC_CUDA(cudaEventCreate(&e0));
    VREC_CUDA(cudaEventCreate(&e1));
    tc_stream_kernel<<<ctas, 128, smem, ctx->stream>>>(buf.p, ntiles, stages, stagger);      // warm-up
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaEventRecord(e0, ctx->stream));
    tc_stream_kernel<<<ctas, 128, smem, ctx->stream>>>(buf.p, ntiles, stages, stagger);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaEventRecord(e1, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.0f;
    VREC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_ms = ms;
    return VREC_OK;
}

// Debug: cycles to ISSUE and to COMPLETE `reps` chains of eight 128x128x16 fp16 MMAs on one SM.
extern "C" int vrec_debug_tc_mma_rate(vrec_ctx *ctx, int swizzled, int reps, int64_t *out2) {
    if (!ctx || !out2 || reps <= 0) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<long long> d;
    VREC_TRY(d.alloc(2));
    size_t smem = 2 * ST_ROWS * ST_D * 2 + 1024;
    VREC_CUDA(cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_rate_kernel<<<1, 128, smem, ctx->stream>>>(swizzled, reps, d.p);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaMemcpyAsync(out2, d.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    return VREC_OK;
}
#endif   // VREC_WITH_MICROBENCH

// Debug: C[128x128] = A[128x128] * B[128x128]^T for caller-supplied fp16 bit patterns (row-major).
extern "C" int vrec_debug_tc_matmul(vrec_ctx *ctx, const uint16_t *A, const uint16_t *B, float *C, int swizzled) {
    if (!ctx || !A || !B || !C) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<__half> dA, dB;
    DevBuf<float> dC;
    VREC_TRY(dA.upload((const __half *)A, ST_ROWS * ST_D, ctx->stream));
    VREC_TRY(dB.upload((const __half *)B, ST_ROWS * ST_D, ctx->stream));
    VREC_TRY(dC.alloc(ST_ROWS * 128));
    size_t smem = 2 * ST_ROWS * ST_D * 2 + 1024;
    VREC_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, ctx->stream>>>(dA.p, dB.p, dC.p, swizzled);
    VREC_LAUNCHED(ctx);
    VREC_CUDA(cudaMemcpyAsync(C, dC.p, ST_ROWS * 128 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    return VREC_OK;
}

// Debug: max |C_tc - C_ref| over a random 128x128x128 fp16 product (C_ref accumulated in double).
extern "C" int vrec_debug_tc_selftest(vrec_ctx *ctx, double *out_max_abs_err) {
    if (!ctx || !out_max_abs_err) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(ctx->device));
    std::vector<__half> hA(ST_ROWS * ST_D), hB(ST_ROWS * ST_D);
    uint64_t s = 88172645463325252ULL;
    auto rnd = [&]() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (float)((s >> 11) % 2001) / 2000.0f;       // [0, 1]
    };
    for (auto &v : hA) v = __float2half(rnd());
    for (auto &v : hB) v = __float2half(rnd() * 0.5f);
    DevBuf<__half> dA, dB;
    DevBuf<float> dC;
    VREC_TRY(dA.upload(hA.data(), hA.size(), ctx->stream));
    VREC_TRY(dB.upload(hB.data(), hB.size(), ctx->stream));
    VREC_TRY(dC.alloc(ST_ROWS * 128));
    size_t smem = 2 * ST_ROWS * ST_D * 2 + 1024;
    VREC_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double worst = 0.0;
    for (int swizzled = 0; swizzled < 2; ++swizzled) {      // both operand layouts
        tc_selftest_kernel<<<1, 128, smem, ctx->stream>>>(dA.p, dB.p, dC.p, swizzled);
        VREC_LAUNCHED(ctx);
        std::vector<float> hC(ST_ROWS * 128);
        VREC_CUDA(cudaMemcpyAsync(hC.data(), dC.p, hC.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        VREC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < ST_ROWS; ++i)
            for (int j = 0; j < 128; ++j) {
                double ref = 0.0;
                for (int k = 0; k < ST_D; ++k)
                    ref += (double)__half2float(hA[i * ST_D + k]) * (double)__half2float(hB[j * ST_D + k]);
                double e = fabs(ref - (double)hC[i * 128 + j]);
                if (e > worst) worst = e;
            }
    }
    *out_max_abs_err = worst;
    return VREC_OK;
}
