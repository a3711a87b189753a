// tcgen05 / TMEM / mbarrier primitives (sm_100a inline PTX) used by the tensor-core KNN filter.
//
// Shared-memory operand layout (K-major, no swizzle, "interleaved" core matrices), for a tile of
// ROWS rows x D fp16:  16-byte chunk c (8 consecutive k values) of row r lives at byte
//     c * (ROWS * 16) + r * 16
// so a core matrix (8 rows x 16 B) is 128 contiguous bytes, SBO (next 8 rows) = 128 B and
// LBO (next k-chunk) = ROWS * 16 B.  One tcgen05.mma of kind::f16 consumes K = 16 = two chunks.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1, SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;           // descriptor version for Blackwell
    return d;                         // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// SWIZZLE_128B K-major operand tile of ROWS rows x D fp16 (D a multiple of 64): the tile is split in
// K-blocks of 64 halves (128 B per row); inside K-block kb, row r occupies the 128 bytes at
//     kb * (ROWS * 128) + r * 128
// and its 16-byte chunk q (0..7) sits at position q ^ (r & 7)  (Swizzle<3,4,3>; tile base 1 KB aligned).
// SBO (next 8 rows) = 1024 B; LBO is not used by swizzled K-major layouts (encoded as 1).
// The k-th MMA (K = 16 halves = 32 B) of a K-block starts 32 * k bytes into the row.
__device__ __forceinline__ uint32_t sw128_offset(int rows, int r, int chunk /* 16-byte chunk of the row */) {
    const int kb = chunk >> 3, q = chunk & 7;
    return (uint32_t)(kb * rows * 128 + r * 128 + ((q ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // LBO (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version for Blackwell
    d |= (uint64_t)2 << 61;                       // layout_type = SWIZZLE_128B
    return d;
}
// start address of the k-th K=16 step (k = 0 .. D/16-1) of a swizzled tile
__device__ __forceinline__ uint32_t sw128_kstep_addr(uint32_t tile_addr, int rows, int k) {
    return tile_addr + (uint32_t)((k >> 2) * rows * 128 + (k & 3) * 32);
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): F16 x F16 -> F32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4)                      // c_format = F32
           | (0u << 7) | (0u << 10)       // a_format = b_format = F16
           | (0u << 15) | (0u << 16)      // a_major = b_major = K
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads smem through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T ; A = 128 rows (TMEM lanes) x 16 halves = 8 columns at tmem_a
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// this thread's lane: 32 consecutive 32-bit TMEM columns starting at taddr (warp-collective; thread i of the
// warp writes lane taddr.lane + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// one lane of a converged warp (elect.sync): the caller's operands stay warp-uniform, so the compiler keeps them in
// uniform registers and issues tcgen05.mma / commit directly.  Under `if (lane == 0)` it cannot prove that and wraps
// every UTCHMMA in an ELECT / R2UR loop: ~15 dependent instructions per MMA, 850 cycles per 8-MMA tile for one thread.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (no tensor map): `bytes` (multiple of 16) land at smem_dst and are
// accounted as transaction bytes on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// named barrier `id` over `nthreads` threads with an OR-reduction of a predicate
__device__ __forceinline__ int bar_red_or(int id, int nthreads, int pred) {
    int out;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "barrier.cta.red.or.pred.aligned p, %1, %2, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(out)
        : "r"(id), "r"(nthreads), "r"(pred)
        : "memory");
    return out;
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("barrier.cta.sync.aligned %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 32 lanes x 32 columns of fp32 accumulators: thread i of the warp gets lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
