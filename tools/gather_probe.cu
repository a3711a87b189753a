// Micro-benchmark behind the design of sg_spmv_kernel: what bounds a stream of
// (src, w) pairs with one scattered 8-byte gather of x per pair on a B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gather_probe tools/gather_probe.cu
//   tools/gather_probe [nnz_millions]
// Prints one line per variant: time, gathers/s, gathers per clock per SM (at the sampled SM clock),
// and the algorithmic GB/s at 12 B per edge.  Not part of libvrec.so.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ unsigned long long pol_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long pol_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_s_i32(const int *p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_s_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_k_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_k_f64_na(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

__global__ void fill_src(int *src, double *w, long long nnz, unsigned n_x, unsigned long long seed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long st = (long long)gridDim.x * blockDim.x;
    for (; i < nnz; i += st) {
        unsigned long long z = seed + (unsigned long long)i * 0x9e3779b97f4a7c15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        z ^= z >> 31;
        src[i] = (int)(z % n_x);
        w[i] = 0.01;
    }
}
__global__ void fill_x(double *x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) x[i] = 1.0 / (double)(i + 1);
}

// MODE 0: stream src,w + gather.  MODE 1: gather only (hashed index).  MODE 2: stream only (no gather).
// EXTRA_LDS: extra conflict-free LDS.64+STS.64 pairs per gather; EXTRA_SHFL: extra shuffles per gather.
template <int U, int MODE, int EXTRA_LDS, int EXTRA_SHFL, int NOALLOC>
__global__ void __launch_bounds__(256) flat_kernel(const int *__restrict__ src, const double *__restrict__ w,
                                                   const double *__restrict__ x, unsigned n_x, long long nnz,
                                                   double *__restrict__ out) {
    __shared__ double s_buf[256 * 2];
    const unsigned long long pf = pol_first(), pl = pol_last();
    const long long T = (long long)gridDim.x * blockDim.x;
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    s_buf[threadIdx.x] = 0.0;
    s_buf[256 + threadIdx.x] = 0.0;
    for (; k + (long long)(U - 1) * T < nnz; k += (long long)U * T) {
        int c[U];
        double ww[U], xx[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (MODE == 1) {
                unsigned long long z = (unsigned long long)(k + q * T) * 0x9e3779b97f4a7c15ULL;
                z ^= z >> 29;
                c[q] = (int)((unsigned)(z >> 11) % n_x);
                ww[q] = 0.01;
            } else {
                c[q] = ld_s_i32(src + k + q * T, pf);
                ww[q] = ld_s_f64(w + k + q * T, pf);
            }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (MODE == 2) xx[q] = (double)c[q];
            else xx[q] = NOALLOC ? ld_k_f64_na(x + c[q], pl) : ld_k_f64(x + c[q], pl);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            double p = __dmul_rn(xx[q], ww[q]);
#pragma unroll
            for (int e = 0; e < EXTRA_LDS; ++e) {
                s_buf[(e & 1) * 256 + threadIdx.x] = p;
                p = __dadd_rn(p, s_buf[(e & 1) * 256 + (threadIdx.x ^ 1)]);
            }
#pragma unroll
            for (int e = 0; e < EXTRA_SHFL; ++e) {
                int lo = __double2loint(p), hi = __double2hiint(p);
                lo = __shfl_xor_sync(0xffffffffu, lo, 1 << (e % 5));
                p = __hiloint2double(hi, lo);
            }
            acc = __dadd_rn(acc, p);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// Half-warp per row of fixed degree (the shape of sg_spmv_kernel's inner loop), software-pipelined:
// the next row's (src, w) are fetched while the current row's gathers are in flight.
template <int DEG_CHUNKS /* ceil(deg/16) */, int PIPE>
__global__ void __launch_bounds__(256) rows_kernel(const int *__restrict__ src, const double *__restrict__ w,
                                                   const double *__restrict__ x, int deg, long long n_rows,
                                                   double *__restrict__ out) {
    const unsigned long long pf = pol_first(), pl = pol_last();
    const int sub = threadIdx.x & 15;
    long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long G = ((long long)gridDim.x * blockDim.x) >> 4;
    double tot = 0.0;
    int c[DEG_CHUNKS];
    double ww[DEG_CHUNKS];
    if (PIPE && g < n_rows) {
#pragma unroll
        for (int q = 0; q < DEG_CHUNKS; ++q) {
            int k = sub + 16 * q;
            bool ok = k < deg;
            c[q] = ok ? ld_s_i32(src + g * deg + k, pf) : 0;
            ww[q] = ok ? ld_s_f64(w + g * deg + k, pf) : 0.0;
        }
    }
    for (; g < n_rows; g += G) {
        if (!PIPE) {
#pragma unroll
            for (int q = 0; q < DEG_CHUNKS; ++q) {
                int k = sub + 16 * q;
                bool ok = k < deg;
                c[q] = ok ? ld_s_i32(src + g * deg + k, pf) : 0;
                ww[q] = ok ? ld_s_f64(w + g * deg + k, pf) : 0.0;
            }
        }
        double xx[DEG_CHUNKS];
#pragma unroll
        for (int q = 0; q < DEG_CHUNKS; ++q) xx[q] = ld_k_f64(x + c[q], pl);
        double a0 = 0.0, a1 = 0.0;
        double pw[DEG_CHUNKS];
#pragma unroll
        for (int q = 0; q < DEG_CHUNKS; ++q) pw[q] = ww[q];
        if (PIPE && g + G < n_rows) {
#pragma unroll
            for (int q = 0; q < DEG_CHUNKS; ++q) {
                int k = sub + 16 * q;
                bool ok = k < deg;
                c[q] = ok ? ld_s_i32(src + (g + G) * deg + k, pf) : 0;
                ww[q] = ok ? ld_s_f64(w + (g + G) * deg + k, pf) : 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < DEG_CHUNKS; ++q) {
            double p = __dmul_rn(xx[q], pw[q]);
            if (q & 1) a1 = __dadd_rn(a1, p); else a0 = __dadd_rn(a0, p);
        }
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            a0 = __dadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, off));
            a1 = __dadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, off));
        }
        tot = __dadd_rn(tot, __dadd_rn(a0, a1));
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

static float time_it(void (*launch)(void *), void *arg, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch(arg);
    launch(arg);
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch(arg);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

struct Args {
    const int *src;
    const double *w;
    const double *x;
    unsigned n_x;
    long long nnz;
    double *out;
    int grid;
    int deg;
};


static void report(const char *name, float ms, long long nnz, double mhz, int mb, int bps) {
    double gps = nnz / (ms * 1e-3);
    printf("%-34s x=%3d MB  blocks/SM=%d  %7.3f ms  %6.1f G/s  %.3f per clk per SM  %5.0f GB/s algorithmic (%.3f of 6449)\n",
           name, mb, bps, ms, gps * 1e-9, gps / (148.0 * mhz * 1e6), 12.0 * gps * 1e-9, 12.0 * gps * 1e-9 / 6449.1);
    fflush(stdout);
}

// ---- TMA-fed two-phase prototype: a tile of rows (contiguous edge range) lands in shared memory by two bulk
// copies; phase A: every thread takes edges of the tile flat (src from smem, gather x, product back into the
// w slot); phase B: one half-warp per row sums its products from shared memory (lane-strided + butterfly).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar, unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

template <int THREADS, int ET, int STAGES, int U>
__global__ void __launch_bounds__(THREADS) tma_rows_kernel(const int *__restrict__ src, const double *__restrict__ w,
                                                           const double *__restrict__ x, int deg, long long n_rows,
                                                           int rows_per_tile, double *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_w = (double *)smem_raw;                                    // [STAGES][ET]
    int *s_src = (int *)(smem_raw + (size_t)STAGES * ET * 8);            // [STAGES][ET]
    uint64_t *full = (uint64_t *)(smem_raw + (size_t)STAGES * ET * 12);  // [STAGES]
    const unsigned long long pf = pol_first(), pl = pol_last();
    const long long n_tiles = (n_rows + rows_per_tile - 1) / rows_per_tile;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long it) {
        long long tile = blockIdx.x + it * gridDim.x;
        if (tile >= n_tiles) return;
        int st = (int)(it % STAGES);
        long long r0 = tile * rows_per_tile;
        long long r1 = r0 + rows_per_tile < n_rows ? r0 + rows_per_tile : n_rows;
        long long e0 = r0 * deg, e1 = r1 * deg;
        long long e0a = e0 & ~3LL, e1a = (e1 + 3) & ~3LL;
        unsigned n = (unsigned)(e1a - e0a);
        mbar_expect_tx(full + st, n * 12);
        bulk_g2s(s_src + (size_t)st * ET, src + e0a, n * 4, full + st, pf);
        bulk_g2s(s_w + (size_t)st * ET, w + e0a, n * 8, full + st, pf);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (long long it = 0;; ++it) {
        long long tile = blockIdx.x + it * gridDim.x;
        if (tile >= n_tiles) break;
        if (threadIdx.x == 0) issue(it + STAGES - 1);
        int st = (int)(it % STAGES);
        mbar_wait(full + st, (unsigned)((it / STAGES) & 1));
        long long r0 = tile * rows_per_tile;
        long long r1 = r0 + rows_per_tile < n_rows ? r0 + rows_per_tile : n_rows;
        long long e0 = r0 * deg, e1 = r1 * deg;
        int off = (int)(e0 - (e0 & ~3LL)), ne = (int)(e1 - e0);
        const int *ts = s_src + (size_t)st * ET + off;
        double *tw = s_w + (size_t)st * ET + off;
        // phase A
        for (int kb = 0; kb < ne; kb += THREADS * U) {
            int c[U];
            double xx[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                int k = kb + threadIdx.x + q * THREADS;
                c[q] = k < ne ? ts[k] : -1;
            }
#pragma unroll
            for (int q = 0; q < U; ++q) xx[q] = c[q] >= 0 ? ld_k_f64(x + c[q], pl) : 0.0;
#pragma unroll
            for (int q = 0; q < U; ++q) {
                int k = kb + threadIdx.x + q * THREADS;
                if (c[q] >= 0) tw[k] = __dmul_rn(xx[q], tw[k]);
            }
        }
        __syncthreads();
        // phase B
        const int sub = threadIdx.x & 15;
        for (int r = threadIdx.x >> 4; r < (int)(r1 - r0); r += THREADS / 16) {
            const double *p = tw + r * deg;
            double a0 = 0.0, a1 = 0.0;
            for (int k = sub; k < deg; k += 32) {
                a0 = __dadd_rn(a0, p[k]);
                if (k + 16 < deg) a1 = __dadd_rn(a1, p[k + 16]);
            }
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                a0 = __dadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, o));
                a1 = __dadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            }
            if (sub == 0) out[r0 + r] = __dadd_rn(a0, a1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
}

template <int THREADS, int ET, int STAGES, int U>
static void launch_tma(void *p) {
    Args *a = (Args *)p;
    size_t smem = (size_t)STAGES * ET * 12 + 64;
    static bool done = false;
    if (!done) {
        CK(cudaFuncSetAttribute(tma_rows_kernel<THREADS, ET, STAGES, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
        done = true;
    }
    tma_rows_kernel<THREADS, ET, STAGES, U><<<a->grid, THREADS, smem>>>(a->src, a->w, a->x, a->deg, a->nnz / a->deg,
                                                                        (ET - 8) / a->deg, a->out);
}

// ---- warp-specialised persistent prototype: producer thread (bulk copies of src / w / row starts into a ring of
// stages), consumer warps (half-warp per row, row pairs claimed from a per-stage counter, canonical 32-lane sum),
// epilogue warp (x' for the tile's rows, coalesced; residual in fixed order).  No block-wide barrier in the loop.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
struct TileDesc { int row0, row1, e0, e1; };

template <int CW, int ET, int RT, int STAGES, int NA>
__global__ void __launch_bounds__((CW + 2) * 32, 1)
ws_kernel(const int *__restrict__ src, const double *__restrict__ w, const int *__restrict__ rowstart,
          const TileDesc *__restrict__ tiles, int n_tiles, const double *__restrict__ x, double *__restrict__ nx,
          double *__restrict__ block_partials) {
    constexpr int W_BYTES = ET * 8 + 32, S_BYTES = ET * 4 + 16, R_BYTES = (RT + 8) * 4, O_BYTES = RT * 8;
    constexpr int STAGE_BYTES = W_BYTES + S_BYTES + R_BYTES + O_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = (uint64_t *)(smem_raw + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES, *done = empty + STAGES;
    int *claim = (int *)(done + STAGES);
    TileDesc *s_tile = (TileDesc *)(claim + STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
            mbar_init(done + s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long pf = pol_first(), pl = pol_last();
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0;; ++it) {
                int tile = blockIdx.x + it * gridDim.x;
                if (tile >= n_tiles) break;
                int st = it % STAGES;
                if (it >= STAGES) mbar_wait(empty + st, (unsigned)((it / STAGES - 1) & 1));
                TileDesc t = tiles[tile];
                unsigned char *base = smem_raw + (size_t)st * STAGE_BYTES;
                int e0a = t.e0 & ~3, e1a = (t.e1 + 3) & ~3, r0a = t.row0 & ~3, r1a = (t.row1 + 1 + 3) & ~3;
                claim[st] = 0;
                s_tile[st] = t;
                unsigned ne = (unsigned)(e1a - e0a), nr = (unsigned)(r1a - r0a);
                mbar_expect_tx(full + st, ne * 12 + nr * 4);
                bulk_g2s(base, w + e0a, ne * 8, full + st, pf);
                bulk_g2s(base + W_BYTES, src + e0a, ne * 4, full + st, pf);
                bulk_g2s(base + W_BYTES + S_BYTES, rowstart + r0a, nr * 4, full + st, pf);
            }
        }
    } else if (warp == 1) {
        double dsum = 0.0;
        for (int it = 0;; ++it) {
            int tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            int st = it % STAGES;
            mbar_wait(done + st, (unsigned)((it / STAGES) & 1));
            unsigned char *base = smem_raw + (size_t)st * STAGE_BYTES;
            const TileDesc t = s_tile[st];
            const int *s_rs = (const int *)(base + W_BYTES + S_BYTES) + (t.row0 & 3);
            const double *s_out = (const double *)(base + W_BYTES + S_BYTES + R_BYTES);
            const int nrows = t.row1 - t.row0;
            for (int r = lane; r < nrows; r += 32) {
                int n = s_rs[r + 1] - s_rs[r];
                double sigma = n > 0 ? s_out[r] : 0.0;
                long long gi = t.row0 + r;
                double v = __dadd_rn(__dmul_rn(gi == 0 ? 1.0 : 0.0, 0.15), __dmul_rn(sigma, 0.85));
                double d = __dsub_rn(v, x[gi]);
                nx[gi] = v;
                dsum = __dadd_rn(dsum, __dmul_rn(d, d));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) dsum = __dadd_rn(dsum, __shfl_xor_sync(0xffffffffu, dsum, o));
        if (lane == 0) block_partials[blockIdx.x] = dsum;
    } else {
        const int half = lane >> 4, sub = lane & 15;
        for (int it = 0;; ++it) {
            int tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            int st = it % STAGES;
            mbar_wait(full + st, (unsigned)((it / STAGES) & 1));
            unsigned char *base = smem_raw + (size_t)st * STAGE_BYTES;
            const TileDesc t = s_tile[st];
            const double *s_w = (const double *)base;
            const int *s_src = (const int *)(base + W_BYTES);
            const int *s_rs = (const int *)(base + W_BYTES + S_BYTES) + (t.row0 & 3);
            double *s_out = (double *)(base + W_BYTES + S_BYTES + R_BYTES);
            const int nrows = t.row1 - t.row0, npairs = (nrows + 1) >> 1, e0a = t.e0 & ~3;
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(claim + st, 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= npairs) break;
                int r = 2 * j + half;
                int s = 0, n = 0;
                if (r < nrows) {
                    s = s_rs[r] - e0a;
                    n = s_rs[r + 1] - s_rs[r];
                }
                int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
                double a0 = 0.0, a1 = 0.0;
                for (int kb = 0; kb < nmax; kb += 64) {
                    int c[4];
                    double xx[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int k = kb + sub + 16 * q;
                        c[q] = k < n ? s_src[s + k] : -1;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) xx[q] = c[q] >= 0 ? (NA ? ld_k_f64_na(x + c[q], pl) : ld_k_f64(x + c[q], pl)) : 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int k = kb + sub + 16 * q;
                        if (c[q] >= 0) {
                            double p = __dmul_rn(xx[q], s_w[s + k]);
                            if (q & 1) a1 = __dadd_rn(a1, p); else a0 = __dadd_rn(a0, p);
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    a0 = __dadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, o));
                    a1 = __dadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, o));
                }
                if (sub == 0 && r < nrows) s_out[r] = __dadd_rn(a0, a1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(done + st);
        }
    }
}

struct WsArgs {
    Args *a;
    const int *rowstart;
    const TileDesc *tiles;
    int n_tiles;
    double *nx, *bp;
};
template <int CW, int ET, int RT, int STAGES, int NA>
static void launch_ws(void *p) {
    WsArgs *q = (WsArgs *)p;
    constexpr int STAGE_BYTES = ET * 8 + 32 + ET * 4 + 16 + (RT + 8) * 4 + RT * 8;
    size_t smem = (size_t)STAGES * STAGE_BYTES + STAGES * (3 * 8 + 4 + 16) + 64;
    static bool done = false;
    if (!done) {
        CK(cudaFuncSetAttribute(ws_kernel<CW, ET, RT, STAGES, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    ws_kernel<CW, ET, RT, STAGES, NA><<<148, (CW + 2) * 32, smem>>>(q->a->src, q->a->w, q->rowstart, q->tiles, q->n_tiles,
                                                               q->a->x, q->nx, q->bp);
}
template <int CW, int ET, int RT, int STAGES, int NA>
static void run_ws(Args &a, int deg, double mhz, int mb) {
    long long n_rows = a.nnz / deg;
    std::vector<int> rs((size_t)n_rows + 16);
    for (long long r = 0; r <= n_rows; ++r) rs[r] = (int)(r * deg);
    int rpt = std::min(RT, ET / deg);
    std::vector<TileDesc> tiles;
    for (long long r = 0; r < n_rows; r += rpt) {
        long long r1 = std::min<long long>(n_rows, r + rpt);
        tiles.push_back(TileDesc{(int)r, (int)r1, (int)(r * deg), (int)(r1 * deg)});
    }
    int *d_rs;
    TileDesc *d_tiles;
    double *nx, *bp;
    CK(cudaMalloc(&d_rs, rs.size() * 4));
    CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(TileDesc)));
    CK(cudaMalloc(&nx, (n_rows + 16) * 8));
    CK(cudaMalloc(&bp, 148 * 8));
    CK(cudaMemcpy(d_rs, rs.data(), rs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    WsArgs q{&a, d_rs, d_tiles, (int)tiles.size(), nx, bp};
    char name[96];
    snprintf(name, sizeof name, "ws CW=%d ET=%d S=%d%s deg=%d", CW, ET, STAGES, NA ? " na" : "", deg);
    report(name, time_it(launch_ws<CW, ET, RT, STAGES, NA>, &q, 5), a.nnz, mhz, mb, 1);
    CK(cudaFree(d_rs));
    CK(cudaFree(d_tiles));
    CK(cudaFree(nx));
    CK(cudaFree(bp));
}

template <int U, int MODE, int EL, int ES, int NA>
static void launch_flat(void *p) {
    Args *a = (Args *)p;
    flat_kernel<U, MODE, EL, ES, NA><<<a->grid, 256>>>(a->src, a->w, a->x, a->n_x, a->nnz, a->out);
}
template <int DC, int PIPE>
static void launch_rows(void *p) {
    Args *a = (Args *)p;
    rows_kernel<DC, PIPE><<<a->grid, 256>>>(a->src, a->w, a->x, a->deg, a->nnz / a->deg, a->out);
}


int main(int argc, char **argv) {
    long long nnz = (argc > 1 ? atoll(argv[1]) : 256) * 1000000LL;
    nnz -= nnz % 100;
    const bool only_ws = argc > 2;
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int mhz_khz = 0;
    CK(cudaDeviceGetAttribute(&mhz_khz, cudaDevAttrClockRate, dev));
    double mhz = mhz_khz / 1000.0;
    printf("device %s, %d SMs, max SM clock %.0f MHz, L2 %d MB, nnz %lld\n", prop.name, prop.multiProcessorCount, mhz,
           prop.l2CacheSize >> 20, nnz);
    int *src;
    double *w, *x, *out;
    const long long max_x = 80LL * 1024 * 1024 / 8;
    CK(cudaMalloc(&src, nnz * 4));
    CK(cudaMalloc(&w, nnz * 8));
    CK(cudaMalloc(&x, max_x * 8));
    CK(cudaMalloc(&out, std::max<long long>(148LL * 16 * 256, nnz / 50 + 16) * 8));
    fill_x<<<148 * 8, 256>>>(x, max_x);
    Args a{src, w, x, 0, nnz, out, 0, 100};
    const int sizes_mb[] = {16, 32, 48, 64, 80};
    for (int mb : sizes_mb) {
        a.n_x = (unsigned)((long long)mb * 1024 * 1024 / 8);
        fill_src<<<148 * 8, 256>>>(src, w, nnz, a.n_x, 12345);
        CK(cudaDeviceSynchronize());
        for (int bps : {4, 8}) {
            if (only_ws) break;
            a.grid = 148 * bps;
            report("flat U=4", time_it(launch_flat<4, 0, 0, 0, 0>, &a, 5), nnz, mhz, mb, bps);
            report("flat U=8", time_it(launch_flat<8, 0, 0, 0, 0>, &a, 5), nnz, mhz, mb, bps);
            if (mb == 48 || mb == 80) {
                report("flat U=8 gather L1::no_allocate", time_it(launch_flat<8, 0, 0, 0, 1>, &a, 5), nnz, mhz, mb, bps);
                report("flat U=16", time_it(launch_flat<16, 0, 0, 0, 0>, &a, 5), nnz, mhz, mb, bps);
                report("gather only U=8", time_it(launch_flat<8, 1, 0, 0, 0>, &a, 5), nnz, mhz, mb, bps);
                report("stream only U=8", time_it(launch_flat<8, 2, 0, 0, 0>, &a, 5), nnz, mhz, mb, bps);
            }
            if (mb == 48) {
                report("flat U=8 +1 LDS/STS.64 per gather", time_it(launch_flat<8, 0, 1, 0, 0>, &a, 5), nnz, mhz, mb, bps);
                report("flat U=8 +2 LDS/STS.64 per gather", time_it(launch_flat<8, 0, 2, 0, 0>, &a, 5), nnz, mhz, mb, bps);
                report("flat U=8 +1 SHFL per gather", time_it(launch_flat<8, 0, 0, 1, 0>, &a, 5), nnz, mhz, mb, bps);
                report("flat U=8 +2 SHFL per gather", time_it(launch_flat<8, 0, 0, 2, 0>, &a, 5), nnz, mhz, mb, bps);
                report("gather only +2 LDS/STS.64", time_it(launch_flat<8, 1, 2, 0, 0>, &a, 5), nnz, mhz, mb, bps);
                report("gather only +2 SHFL", time_it(launch_flat<8, 1, 0, 2, 0>, &a, 5), nnz, mhz, mb, bps);
            }
        }
        if (mb == 48 || mb == 64) {
            for (int d : {100, 50}) {
                run_ws<30, 4096, 512, 4, 1>(a, d, mhz, mb);
                run_ws<30, 6144, 512, 2, 1>(a, d, mhz, mb);
                run_ws<14, 1024, 128, 4, 0>(a, d, mhz, mb);
                run_ws<30, 1024, 128, 4, 0>(a, d, mhz, mb);
                run_ws<14, 1024, 128, 2, 0>(a, d, mhz, mb);
                run_ws<14, 2048, 256, 3, 0>(a, d, mhz, mb);
                run_ws<30, 2048, 256, 3, 0>(a, d, mhz, mb);
                run_ws<30, 2048, 256, 2, 0>(a, d, mhz, mb);
                run_ws<30, 2048, 256, 2, 1>(a, d, mhz, mb);
                run_ws<22, 1536, 256, 3, 0>(a, d, mhz, mb);
            }
            for (int d : {100, 50}) {
                if (only_ws) break;
                a.deg = d;
                a.grid = 148 * 3;
                report(d == 100 ? "tma T=256 ET=3072 S=2 U=8 deg=100" : "tma T=256 ET=3072 S=2 U=8 deg=50",
                       time_it(launch_tma<256, 3072, 2, 8>, &a, 5), nnz, mhz, mb, 3);
                a.grid = 148 * 2;
                report(d == 100 ? "tma T=256 ET=3072 S=3 U=8 deg=100" : "tma T=256 ET=3072 S=3 U=8 deg=50",
                       time_it(launch_tma<256, 3072, 3, 8>, &a, 5), nnz, mhz, mb, 2);
                report(d == 100 ? "tma T=512 ET=4096 S=2 U=8 deg=100" : "tma T=512 ET=4096 S=2 U=8 deg=50",
                       time_it(launch_tma<512, 4096, 2, 8>, &a, 5), nnz, mhz, mb, 2);
                report(d == 100 ? "tma T=512 ET=4096 S=2 U=4 deg=100" : "tma T=512 ET=4096 S=2 U=4 deg=50",
                       time_it(launch_tma<512, 4096, 2, 4>, &a, 5), nnz, mhz, mb, 2);
                a.grid = 148 * 4;
                report(d == 100 ? "tma T=256 ET=2048 S=2 U=8 deg=100" : "tma T=256 ET=2048 S=2 U=8 deg=50",
                       time_it(launch_tma<256, 2048, 2, 8>, &a, 5), nnz, mhz, mb, 4);
                report(d == 100 ? "tma T=256 ET=2048 S=2 U=4 deg=100" : "tma T=256 ET=2048 S=2 U=4 deg=50",
                       time_it(launch_tma<256, 2048, 2, 4>, &a, 5), nnz, mhz, mb, 4);
            }
        }
        if (!only_ws && (mb == 48 || mb == 32)) {
            for (int bps : {4, 8}) {
                a.grid = 148 * bps;
                a.deg = 100;
                report("rows deg=100 half-warp", time_it(launch_rows<7, 0>, &a, 5), nnz, mhz, mb, bps);
                report("rows deg=100 half-warp pipelined", time_it(launch_rows<7, 1>, &a, 5), nnz, mhz, mb, bps);
                a.deg = 50;
                report("rows deg=50 half-warp", time_it(launch_rows<4, 0>, &a, 5), nnz, mhz, mb, bps);
                report("rows deg=50 half-warp pipelined", time_it(launch_rows<4, 1>, &a, 5), nnz, mhz, mb, bps);
            }
        }
    }
    return 0;
}
