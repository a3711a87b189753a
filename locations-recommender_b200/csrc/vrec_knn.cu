// KNN path: cosine similarity of a batch of targets against every person of a region-set over
// sparse place + category rating vectors (CSR, precomputed norms), weighted combine, top-K
// neighbours, similarity-weighted estimated_rating per place, ranked top-N places.
// Reference: knn/KnnRecommender.scala:22-96, knn/Distance.scala:7-16, knn/KnnRecommenderMain.scala:90-102.
#include <algorithm>
#include <numeric>
#include <string.h>

#include "vrec_internal.cuh"
#include "vrec_tc.cuh"

namespace {

#ifndef VREC_WS_ABLATE
#define VREC_WS_ABLATE 0          // > 0: timing experiments that break the results (never in a shipped build)
#endif

// ------------------------------------------------------------------ device views
struct KnnVec {                 // one rating-vector table in CSR form
    const int *rowptr;          // [P+1]
    const int *col;             // ascending within a row
    const double *val;
    const double *len;          // vectorLength per row (knn/Distance.scala:11-16)
};

struct KnnDev {
    long long P;
    const long long *person;    // ascending person ids
    KnnVec place, cat;
};

struct Nb {                     // one neighbour candidate
    double sim;
    int idx;                    // person index (== rank of person_id)
    int pad;
};

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_PER_THREAD = 4;
constexpr int TOPK_BUF = 2048;          // smem candidate buffer; requires K <= 1024
constexpr int TOPK_MAX_K = 1024;
constexpr int MERGE_THREADS = 256;
constexpr int RATE_WARPS = 4;

__device__ __forceinline__ bool nb_before(const Nb &a, const Nb &b) {
    return a.sim > b.sim || (a.sim == b.sim && a.idx < b.idx);
}

// spark-mllib-local 3.1.2 BLAS.dot(sparse, sparse) ("y catching x"), called from knn/Distance.scala:8
__device__ __forceinline__ double sparse_dot(const int *__restrict__ xi, const double *__restrict__ xv, int nx,
                                             const int *__restrict__ yi, const double *__restrict__ yv, int ny) {
    int kx = 0, ky = 0;
    double sum = 0.0;
    while (kx < nx && ky < ny) {
        int ix = xi[kx];
        while (ky < ny && yi[ky] < ix) ky++;
        if (ky < ny && yi[ky] == ix) {
            sum = xadd(sum, xmul(xv[kx], yv[ky]));
            ky++;
        }
        kx++;
    }
    return sum;
}

// cosineSimilarity(row, target) of one table, or 0 when the row is absent / not kept.
__device__ __forceinline__ double table_similarity(const KnnVec &v, long long i, int ts, int tn, double tlen,
                                                   bool &keep) {
    int s = v.rowptr[i], n = v.rowptr[i + 1] - s;
    if (n == 0) return 0.0;
    double d = sparse_dot(v.col + s, v.val + s, n, v.col + ts, v.val + ts, tn);
    double den = xmul(v.len[i], tlen);                 // vectorLength(v1) * vectorLength(v2)
    double c = xdiv(d, den);
    if (c > 0) {                                       // knn/KnnRecommender.scala:91
        keep = true;
        return c;
    }
    return 0.0;
}

struct TargetRows {
    int t;
    int ps, pn, cs, cn;
    double plen, clen;
    // the target's rows (global memory, or a shared-memory copy staged by the caller)
    const int *pcol, *ccol;
    const double *pval, *cval;
};

__device__ __forceinline__ TargetRows load_target(const KnnDev &d, int t) {
    TargetRows r;
    r.t = t;
    r.ps = d.place.rowptr[t];
    r.pn = d.place.rowptr[t + 1] - r.ps;
    r.cs = d.cat.rowptr[t];
    r.cn = d.cat.rowptr[t + 1] - r.cs;
    r.plen = d.place.len[t];
    r.clen = d.cat.len[t];
    r.pcol = d.place.col + r.ps;
    r.pval = d.place.val + r.ps;
    r.ccol = d.cat.col + r.cs;
    r.cval = d.cat.val + r.cs;
    return r;
}

// findSimilarPersons for one (candidate, target) pair: knn/KnnRecommender.scala:27-45.
// 0.0 means "not a candidate" (the target itself, or neither similarity > 0).
__device__ __forceinline__ double pair_similarity(const KnnDev &d, long long i, const TargetRows &t, double pw,
                                                  double cw) {
    if (i == t.t) return 0.0;                          // where(person_id =!= personId), :89
    bool keep = false;
    double ps = table_similarity(d.place, i, t.ps, t.pn, t.plen, keep);
    double cs = table_similarity(d.cat, i, t.cs, t.cn, t.clen, keep);
    if (!keep) return 0.0;
    return xadd(xmul(ps, pw), xmul(cs, cw));           // :42-44
}

// ------------------------------------------------------------------ kernels: load time

// vectorLength per row: values.map(v => v*v).sum left to right, sqrt (knn/Distance.scala:11-16)
__global__ void knn_norms_kernel(const int *__restrict__ rowptr, const double *__restrict__ val, long long P,
                                 double *__restrict__ len) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double s = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s = xadd(s, xmul(val[k], val[k]));
    len[i] = sqrt(s);
}

// ------------------------------------------------------------------ kernels: query

// target id -> person index; status ENOENT when the person has no row in either table
// ("No such person", knn/KnnRecommender.scala:77-83)
__global__ void knn_lookup_kernel(KnnDev d, const long long *__restrict__ targets, int n, int *__restrict__ tidx,
                                  int *__restrict__ status) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    long long id = targets[q];
    long long lo = 0, hi = d.P;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (d.person[mid] < id) lo = mid + 1; else hi = mid;
    }
    int t = -1;
    if (lo < d.P && d.person[lo] == id) {
        bool has_place = d.place.rowptr[lo + 1] > d.place.rowptr[lo];
        bool has_cat = d.cat.rowptr[lo + 1] > d.cat.rowptr[lo];
        if (has_place && has_cat) t = (int)lo;
    }
    tidx[q] = t;
    status[q] = t >= 0 ? VREC_OK : VREC_ENOENT;
}

__device__ __forceinline__ void bitonic_sort_desc(Nb *buf, int n_pow2, int tid, int nthreads) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pow2; i += nthreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    Nb a = buf[i], b = buf[ixj];
                    bool up = (i & k) == 0;            // "up" block: best first
                    bool swap = up ? nb_before(b, a) : nb_before(a, b);
                    if (swap) {
                        buf[i] = b;
                        buf[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Fused similarity + streaming top-K for K <= 1024.  Block (tt, sp) scans the candidate range
// sp of target tt and emits its best K as part[tt][sp][*] in (similarity desc, index asc) order.
__global__ void __launch_bounds__(TOPK_THREADS)
knn_topk_kernel(KnnDev d, const int *__restrict__ tidx, int K, int S, double pw, double cw,
                Nb *__restrict__ part, int *__restrict__ part_cnt) {
    const int tt = blockIdx.x, sp = blockIdx.y, tid = threadIdx.x;
    const int t = tidx[tt];
    if (t < 0) {
        if (tid == 0) part_cnt[tt * S + sp] = 0;
        return;
    }
    __shared__ Nb buf[TOPK_BUF];
    __shared__ int s_cnt;
    __shared__ int s_have_thr;
    __shared__ Nb s_thr;
    if (tid == 0) {
        s_cnt = 0;
        s_have_thr = 0;
    }
    __syncthreads();
    const TargetRows tr = load_target(d, t);
    const long long lo = d.P * sp / S, hi = d.P * (sp + 1) / S;
    for (long long base = lo; base < hi; base += TOPK_THREADS * TOPK_PER_THREAD) {
        const bool have_thr = s_have_thr;
        const Nb thr = s_thr;
#pragma unroll
        for (int c = 0; c < TOPK_PER_THREAD; ++c) {
            long long i = base + (long long)c * TOPK_THREADS + tid;
            if (i < hi) {
                double s = pair_similarity(d, i, tr, pw, cw);
                if (s > 0) {
                    Nb e;
                    e.sim = s;
                    e.idx = (int)i;
                    e.pad = 0;
                    if (!have_thr || nb_before(e, thr)) {
                        int slot = atomicAdd(&s_cnt, 1);
                        buf[slot] = e;
                    }
                }
            }
        }
        __syncthreads();
        int cnt = s_cnt;
        __syncthreads();                                           // nobody bumps s_cnt before all have read it
        if (cnt > TOPK_BUF - TOPK_THREADS * TOPK_PER_THREAD) {     // uniform
            int np2 = next_pow2(cnt);
            for (int i = cnt + tid; i < np2; i += TOPK_THREADS) {
                buf[i].sim = -1.0;
                buf[i].idx = 0x7fffffff;
            }
            __syncthreads();
            bitonic_sort_desc(buf, np2, tid, TOPK_THREADS);
            if (tid == 0) {
                if (cnt >= K) {
                    s_cnt = K;
                    s_thr = buf[K - 1];
                    s_have_thr = 1;
                }
            }
            __syncthreads();
        }
    }
    // final sort and emit
    int cnt = s_cnt;
    int np2 = next_pow2(max(cnt, 1));
    for (int i = cnt + tid; i < np2; i += TOPK_THREADS) {
        buf[i].sim = -1.0;
        buf[i].idx = 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_desc(buf, np2, tid, TOPK_THREADS);
    int keep = min(cnt, K);
    Nb *out = part + ((size_t)tt * S + sp) * K;
    for (int i = tid; i < keep; i += TOPK_THREADS) out[i] = buf[i];
    if (tid == 0) part_cnt[tt * S + sp] = keep;
}


// ---------------------------------------------------------------------------------------
// Tiled batch kernel (K <= 1024, non-negative data, cat_dim <= 32).
//
// Every person carries a dense fp32 feature vector of TILE_D = 32 dims: its category vector
// divided by its length, then the values of the `n_head` most popular ("head") places divided
// by the place-vector length.  For a target t with the weights folded in,
//     U(t, c) = sum_d tvec[t][d] * F[d][c]  =  cw * cos_cat + pw * (head part of cos_place)
// equals the combined similarity of every pair that shares no non-head ("tail") place, up to
// fp32 rounding (<= 2e-6, margin 3e-5).  The kernel
//   1. enumerates the pairs that DO share a tail place through the place postings and
//      evaluates them exactly (each pair once: at its smallest shared tail place);
//   2. streams all candidates of its range, computes U for TILE_T targets per candidate from
//      shared memory, and sends only pairs with U + margin >= current K-th best to the exact
//      fp64 evaluation (pairs found to share a tail place are skipped: step 1 owns them).
// Exact survivors go into per-target binary heaps in shared memory (worst neighbour at the
// root).  Results are identical to knn_topk_kernel: the filter only prunes.
// ---------------------------------------------------------------------------------------
constexpr int TILE_D = 32;
constexpr int TILE_THREADS = 256;
constexpr int TILE_QCAP = 2048;
constexpr int TILE_TVEC_STRIDE = 36;            // 32 dims + threshold + pad (16-byte aligned rows)
constexpr float TILE_MARGIN = 3e-5f;

struct TileAux {
    const float *feat;          // [TILE_D][fstride], dim-major
    long long fstride;
    const short *head_slot;     // per place: head slot or -1 (tail)
    const int *pcp;             // place postings (CSC of the place vectors)
    const int *pper;
    // packed per-person records for the exact evaluation (nullptr: read the CSR arrays instead).
    // meta[i] = offset (8-byte words, bits 0..39) | place count (bits 40..51) | category count (52..63)
    // record  = { |place|, |cat|, place cols (int, padded to 8 B), place vals, cat cols (padded), cat vals }
    // place cols carry "tail place" flags of the two head sets in bits 31 / 30.
    const unsigned long long *meta;
    const double *rec;
    unsigned tail_bit;          // which flag applies to this kernel's head set
    bool vals_f32 = false;      // record values stored as floats (see the record layout below)
    // knn_tc_ws_kernel: the targets' category vectors as dense rows [n_targets][cat_dim] (global scratch)
    double *tdense = nullptr;
    int cat_dim = 0;
    // compact records (round 2): the candidate side of the two hot evaluators (exact_pair_sig, exact_pair_staged)
    // when every value of the region-set is a small integer -- see the layout below.  nullptr: not available.
    const unsigned long long *cmeta = nullptr;
    const unsigned long long *crec = nullptr;
    const double *plen = nullptr, *clen = nullptr;   // |place| / |cat| per person (the targets' lengths)
};

constexpr unsigned REC_TAIL_TC = 0x80000000u, REC_TAIL_TILE = 0x40000000u, REC_COL_MASK = 0x3fffffffu;
// Record layout in 8-byte words, every section 16-byte aligned so that it can be fetched with 128-bit loads:
//   [0..2)  |place|, |cat|
//   place cols: ints, padded to a multiple of 4   -> rec_cols_words(np) words
//   place vals: doubles, padded to a multiple of 2 -> rec_vals_words(np) words; or, when every value of the
//               region-set is exactly representable in fp32 (visit counts always are), floats padded to a
//               multiple of 4 (TileAux::vals_f32): a third fewer bytes per record on the HBM-bound paths
//   cat cols, cat vals: same
__host__ __device__ constexpr int rec_cols_words(int n) { return ((n + 3) / 4) * 2; }
__host__ __device__ constexpr int rec_vals_words(int n, bool f32) { return f32 ? ((n + 3) / 4) * 2 : ((n + 1) / 2) * 2; }

// 4 entries (cols + vals) of a record section with three 128-bit loads
__device__ __forceinline__ void rec_load4(const int *cols, const double *vals, int k0, unsigned (&c)[4], double (&x)[4],
                                          bool f32) {
    const int4 ci = __ldg(reinterpret_cast<const int4 *>(cols + k0));
    c[0] = (unsigned)ci.x; c[1] = (unsigned)ci.y; c[2] = (unsigned)ci.z; c[3] = (unsigned)ci.w;
    if (f32) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(vals) + k0));
        x[0] = (double)v.x; x[1] = (double)v.y; x[2] = (double)v.z; x[3] = (double)v.w;
    } else {
        const double2 v01 = __ldg(reinterpret_cast<const double2 *>(vals + k0));
        const double2 v23 = __ldg(reinterpret_cast<const double2 *>(vals + k0 + 2));
        x[0] = v01.x; x[1] = v01.y; x[2] = v23.x; x[3] = v23.y;
    }
}
// one value of a record section
__device__ __forceinline__ double rec_val(const double *vals, int i, bool f32) {
    return f32 ? (double)__ldg(reinterpret_cast<const float *>(vals) + i) : __ldg(vals + i);
}

// ---- compact records -------------------------------------------------------------------------------
// Visit counts are small integers.  When every place value of a region-set is an integer in [0, 255], every
// category value an integer in [0, 1023], place_dim <= 2^22 and cat_dim <= 64, a person's record shrinks to
//   place entries, 4 bytes each:  col (bits 0..21) | value (22..29) | tail flags (30: tile head set, 31: tc head set)
//   category entries, 2 bytes each: col (bits 0..5) | value (6..15)
// both sections padded to 8 bytes, no header: 10^6 persons of BASELINE config 3 take ~50 MB instead of ~160 MB,
// which is what lets the ~3 K exact evaluations per target hit the L2 instead of HBM.  cmeta[i] has the layout
// of meta[i] (offset in 8-byte words | place count << 40 | category count << 52).  The vector lengths are not
// stored: sum(v^2) is an exact integer, so sqrt((double)sum) is bit for bit the length knn_norms_kernel
// computed (Distance.vectorLength, knn/Distance.scala:11-16).
#ifndef VREC_POST_SEARCH
#define VREC_POST_SEARCH 1        // 1: two-pass place matching in the postings evaluator (see exact_pair_staged_compact)
#endif
constexpr unsigned CREC_COL_MASK = 0x3fffffu;
constexpr int CREC_PW = 6, CREC_CW = 3;            // words fetched up front: 12 place + 12 category entries
__host__ __device__ constexpr int crec_place_words(int n) { return (n + 1) >> 1; }
__host__ __device__ constexpr int crec_cat_words(int n) { return (n + 3) >> 2; }
__device__ __forceinline__ unsigned long long ld_rec_u64(const unsigned long long *p, unsigned long long pol) {
    unsigned long long v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}

// The candidate's record, fetched with one round trip: all words of a typical record are requested before
// the first is used.  place(e) / cat(e) return entry e (any e; entries beyond the prefetched words are
// loaded on demand).
struct CRec {
    const unsigned long long *r;
    int np, nc, pwords;
    unsigned long long pw[CREC_PW], cw[CREC_CW];
    __device__ __forceinline__ void open(const TileAux &aux, unsigned long long m, unsigned long long pol) {
        np = (int)((m >> 40) & 0xfffu);
        nc = (int)(m >> 52);
        r = aux.crec + (m & 0xffffffffffULL);
        pwords = crec_place_words(np);
        const int cwords = crec_cat_words(nc);
#pragma unroll
        for (int j = 0; j < CREC_PW; ++j) pw[j] = j < pwords ? ld_rec_u64(r + j, pol) : 0ULL;
#pragma unroll
        for (int j = 0; j < CREC_CW; ++j) cw[j] = j < cwords ? ld_rec_u64(r + pwords + j, pol) : 0ULL;
    }
};

// Sparse dot of candidate row [s, s+n) with the target row [ts, ts+tn) of one table, in the exact
// order of the mllib merge (matches visited in ascending index), but with the candidate's entries
// fetched 8 at a time by independent loads instead of one dependent load per merge step.  The
// target row is shared by the whole block and stays L1-resident.
template <bool TRACK_TAIL>
__device__ __forceinline__ double bulk_dot(const KnnVec &v, int s, int n, const int *tcol, const double *tval, int tn,
                                           const short *head_slot, int &min_tail) {
    double sum = 0.0;
    int ky = 0;
    for (int k0 = 0; k0 < n && ky < tn; k0 += 8) {
        int c[8];
        double x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            bool ok = k0 + e < n;
            c[e] = ok ? __ldg(v.col + s + k0 + e) : 0x7fffffff;
            x[e] = ok ? __ldg(v.val + s + k0 + e) : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int ix = c[e];
            if (ix == 0x7fffffff) break;
            while (ky < tn && tcol[ky] < ix) ky++;
            if (ky < tn && tcol[ky] == ix) {
                sum = xadd(sum, xmul(x[e], tval[ky]));
                if (TRACK_TAIL) {
                    if (min_tail < 0 && head_slot[ix] < 0) min_tail = ix;
                }
                ky++;
            }
        }
    }
    return sum;
}

// bulk_dot over a packed record: same order of operations, one contiguous block of memory
template <bool TRACK_TAIL>
__device__ __forceinline__ double packed_dot(const int *__restrict__ pc, const double *__restrict__ pv, int n,
                                             const int *tcol, const double *tval, int tn, unsigned tail_bit,
                                             int &min_tail, bool F32) {
    double sum = 0.0;
    int ky = 0;
    for (int k0 = 0; k0 < n && ky < tn; k0 += 4) {
        unsigned c[4];
        double x[4];
        rec_load4(pc, pv, k0, c, x, F32);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (k0 + e >= n) break;
            int ix = (int)(c[e] & REC_COL_MASK);
            while (ky < tn && tcol[ky] < ix) ky++;
            if (ky < tn && tcol[ky] == ix) {
                sum = xadd(sum, xmul(x[e], tval[ky]));
                if (TRACK_TAIL) {
                    if (min_tail < 0 && (c[e] & tail_bit)) min_tail = ix;
                }
                ky++;
            }
        }
    }
    return sum;
}

__device__ unsigned long long g_probe[8];
__device__ int g_probe_on;

__device__ __forceinline__ double exact_pair_packed(const TileAux &aux, long long i, const TargetRows &t, double pw,
                                                    double cw, int &min_tail) {
    min_tail = -1;
    const bool F32 = aux.vals_f32;
    if (i == t.t) return 0.0;
    const bool pr = g_probe_on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32;
    long long c0 = clock64();
    const unsigned long long m = __ldg(aux.meta + i);                   // round trip 1 (8 MB table: L2)
    const int np = (int)((m >> 40) & 0xfffu), nc = (int)(m >> 52);
    const double *r = aux.rec + (m & 0xffffffffffULL);                  // round trip 2: one contiguous record
    if (pr && m == 0xffffffffffffffffULL) g_probe[7] = 1;              // keep the dependency
    __syncwarp(__activemask());
    long long c1 = clock64();
    const double2 lens = __ldg(reinterpret_cast<const double2 *>(r));
    const double plen = lens.x, clen = lens.y;
    const int *pc = reinterpret_cast<const int *>(r + 2);
    const double *pv = r + 2 + rec_cols_words(np);
    const int *cc = reinterpret_cast<const int *>(pv + rec_vals_words(np, F32));
    const double *cvp = pv + rec_vals_words(np, F32) + rec_cols_words(nc);
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (np > 0) {
        double sum = packed_dot<true>(pc, pv, np, t.pcol, t.pval, t.pn, aux.tail_bit, min_tail, F32);
        double c = xdiv(sum, xmul(plen, t.plen));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    __syncwarp(__activemask());
    long long c2 = clock64();
    if (nc > 0) {
        int dummy = 0;
        double sum = packed_dot<false>(cc, cvp, nc, t.ccol, t.cval, t.cn, 0u, dummy, F32);
        double c = xdiv(sum, xmul(clen, t.clen));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    __syncwarp(__activemask());
    if (pr && threadIdx.x == 0) {
        long long c3 = clock64();
        g_probe[0] += (unsigned long long)(c1 - c0);
        g_probe[1] += (unsigned long long)(c2 - c1);
        g_probe[2] += (unsigned long long)(c3 - c2);
        g_probe[3] += 1;
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// Target of the postings pass, staged in shared memory by its warp: sorted place row, a 64-bit
// signature of its places, and its category vector as a dense array.
struct StagedTarget {
    int t;
    int pn;
    const int *pcol;            // shared memory, ascending
    const double *pval;
    unsigned long long sig;     // bit h(col) set for every place of the target
    const double *cat_dense;    // [cat_dim] shared memory, 0.0 where the target has no rating
    double plen, clen;
};

__device__ __forceinline__ unsigned sig_bit(int col) { return ((unsigned)col * 0x9E3779B1u) >> 26; }

// exact_pair_staged on the candidate's compact record (same operations in the same order, so the same bits).
__device__ __forceinline__ double exact_pair_staged_compact(const TileAux &aux, long long i, const StagedTarget &t,
                                                            double pw, double cw, int &min_tail, double thr) {
    min_tail = -1;
    if (i == t.t) return 0.0;
    const unsigned long long pol = policy_evict_last();
    CRec rc;
    rc.open(aux, __ldg(aux.cmeta + i), pol);
    const int np = rc.np, nc = rc.nc;
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (np > 0) {
        double sum = 0.0;
        unsigned sq = 0u;
        // the target's entry for column ix, if any: binary search over its sorted places (shared memory)
        auto place_match = [&](unsigned craw) {
            const int ix = (int)(craw & CREC_COL_MASK);
            int lo = 0, hi = t.pn;                                   // first index with pcol >= ix
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (t.pcol[mid] < ix) lo = mid + 1; else hi = mid;
            }
            if (lo < t.pn && t.pcol[lo] == ix) {
                sum = xadd(sum, xmul((double)((craw >> 22) & 0xffu), t.pval[lo]));
                if (min_tail < 0 && (craw & aux.tail_bit)) min_tail = ix;
            }
        };
#if VREC_POST_SEARCH == 1
        // Pass 1 (all lanes, no memory): squares and the signature test of the prefetched entries -> a bit mask.
        // Pass 2: each lane searches only ITS entries that passed, in ascending order (the order the products are
        // added in), re-reading the entry word from the L1.  The search used to sit inside the 12 unrolled entry
        // slots and ran in nearly every slot for ~5 of 32 lanes: 47 % of the kernel's instructions
        // (profiles/r2_knn_hot_lines.txt); now the warp runs as many search rounds as its busiest lane has hits.
        unsigned pend = 0u;
#pragma unroll
        for (int j = 0; j < CREC_PW; ++j) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (2 * j + hf < np) {
                    const unsigned craw = hf ? (unsigned)(rc.pw[j] >> 32) : (unsigned)rc.pw[j];
                    const unsigned v = (craw >> 22) & 0xffu;
                    sq += v * v;
                    if ((t.sig >> sig_bit((int)(craw & CREC_COL_MASK))) & 1ULL) pend |= 1u << (2 * j + hf);
                }
            }
        }
        while (pend) {
            const int e = __ffs(pend) - 1;
            pend &= pend - 1;
            const unsigned long long wd = ld_rec_u64(rc.r + (e >> 1), pol);
            place_match((e & 1) ? (unsigned)(wd >> 32) : (unsigned)wd);
        }
#else
#pragma unroll
        for (int j = 0; j < CREC_PW; ++j) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (2 * j + hf < np) {
                    const unsigned craw = hf ? (unsigned)(rc.pw[j] >> 32) : (unsigned)rc.pw[j];
                    const unsigned v = (craw >> 22) & 0xffu;
                    sq += v * v;
                    if ((t.sig >> sig_bit((int)(craw & CREC_COL_MASK))) & 1ULL) place_match(craw);
                }
            }
        }
#endif
        for (int j = CREC_PW; j < rc.pwords; ++j) {                  // long records: the rest, entry by entry
            const unsigned long long wd = ld_rec_u64(rc.r + j, pol);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (2 * j + hf < np) {
                    const unsigned craw = hf ? (unsigned)(wd >> 32) : (unsigned)wd;
                    const unsigned v = (craw >> 22) & 0xffu;
                    sq += v * v;
                    if ((t.sig >> sig_bit((int)(craw & CREC_COL_MASK))) & 1ULL) place_match(craw);
                }
            }
        }
        const double c = xdiv(sum, xmul(sqrt((double)sq), t.plen));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    // see exact_pair_staged: the category section cannot lift the pair over `thr` any more
    if (xadd(xmul(ps_sim, pw), xmul(1.0000001, cw)) < thr) return 0.0;
    if (nc > 0) {
        double sum = 0.0;
        unsigned sq = 0u;
        auto cat_term = [&](unsigned h) {
            const unsigned v = (h >> 6) & 0x3ffu;
            sq += v * v;
            sum = xadd(sum, xmul((double)v, t.cat_dense[h & 63u]));
        };
#pragma unroll
        for (int j = 0; j < CREC_CW; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * j + q < nc) cat_term((unsigned)(rc.cw[j] >> (16 * q)) & 0xffffu);
        }
        for (int j = CREC_CW; j < crec_cat_words(nc); ++j) {
            const unsigned long long wd = ld_rec_u64(rc.r + rc.pwords + j, pol);
            for (int q = 0; q < 4; ++q)
                if (4 * j + q < nc) cat_term((unsigned)(wd >> (16 * q)) & 0xffffu);
        }
        const double c = xdiv(sum, xmul(sqrt((double)sq), t.clen));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// Same value as exact_pair_packed, without data-dependent merge loops:
//  * category dot: sum over the candidate's entries (ascending) of x * dense[col]; entries the target
//    lacks contribute x * 0.0 = +0.0, and s + 0.0 == s, so the sum equals the merge's bit for bit
//    (all values are non-negative on this path);
//  * place dot: candidate entries in ascending order; an entry can only match if its signature bit is
//    set, and then a fixed-depth binary search over the target row finds it.
__device__ __forceinline__ double exact_pair_staged(const TileAux &aux, long long i, const StagedTarget &t, double pw,
                                                    double cw, int &min_tail, double thr = 0.0) {
    if (aux.cmeta) return exact_pair_staged_compact(aux, i, t, pw, cw, min_tail, thr);
    min_tail = -1;
    const bool F32 = aux.vals_f32;
    if (i == t.t) return 0.0;
    const unsigned long long m = __ldg(aux.meta + i);
    const int np = (int)((m >> 40) & 0xfffu), nc = (int)(m >> 52);
    const double *r = aux.rec + (m & 0xffffffffffULL);
    const double2 lens = __ldg(reinterpret_cast<const double2 *>(r));
    const double plen = lens.x, clen = lens.y;
    const int *pc = reinterpret_cast<const int *>(r + 2);
    const double *pv = r + 2 + rec_cols_words(np);
    const int *cc = reinterpret_cast<const int *>(pv + rec_vals_words(np, F32));
    const double *cvp = pv + rec_vals_words(np, F32) + rec_cols_words(nc);
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (np > 0) {
        double sum = 0.0;
        auto place_term = [&](unsigned craw, double x) {
            const int ix = (int)(craw & REC_COL_MASK);
            if ((t.sig >> sig_bit(ix)) & 1ULL) {
                int lo = 0, hi = t.pn;                               // first index with pcol >= ix
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (t.pcol[mid] < ix) lo = mid + 1; else hi = mid;
                }
                if (lo < t.pn && t.pcol[lo] == ix) {
                    sum = xadd(sum, xmul(x, t.pval[lo]));
                    if (min_tail < 0 && (craw & aux.tail_bit)) min_tail = ix;
                }
            }
        };
        int k0 = 0;
        if (F32) {
            // float records: the first 8 place entries (most rows) with four loads issued together with the
            // header's, i.e. one memory round trip for header + place section
            const int4 c0 = __ldg(reinterpret_cast<const int4 *>(pc));
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(pv));
            int4 c1 = make_int4(0, 0, 0, 0);
            float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (np > 4) {
                c1 = __ldg(reinterpret_cast<const int4 *>(pc) + 1);
                v1 = __ldg(reinterpret_cast<const float4 *>(pv) + 1);
            }
            place_term((unsigned)c0.x, (double)v0.x);
            if (np > 1) place_term((unsigned)c0.y, (double)v0.y);
            if (np > 2) place_term((unsigned)c0.z, (double)v0.z);
            if (np > 3) place_term((unsigned)c0.w, (double)v0.w);
            if (np > 4) place_term((unsigned)c1.x, (double)v1.x);
            if (np > 5) place_term((unsigned)c1.y, (double)v1.y);
            if (np > 6) place_term((unsigned)c1.z, (double)v1.z);
            if (np > 7) place_term((unsigned)c1.w, (double)v1.w);
            k0 = 8;
        }
        for (; k0 < np; k0 += 4) {
            unsigned c[4];
            double x[4];
            rec_load4(pc, pv, k0, c, x, F32);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k0 + e < np) place_term(c[e], x[e]);
        }
        double c = xdiv(sum, xmul(plen, t.plen));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    // A cosine exceeds 1 by a few ulps at most: if even a category similarity of 1.0000001 leaves the pair
    // below `thr` (a lower bound of what the caller still accepts), the category section -- three more
    // dependent memory round trips -- is not read at all.
    if (xadd(xmul(ps_sim, pw), xmul(1.0000001, cw)) < thr) return 0.0;
    if (nc > 0) {
        double sum = 0.0;
        for (int k0 = 0; k0 < nc; k0 += 4) {
            unsigned c[4];
            double x[4];
            rec_load4(cc, cvp, k0, c, x, F32);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k0 + e < nc) sum = xadd(sum, xmul(x[e], t.cat_dense[c[e]]));
        }
        double c = xdiv(sum, xmul(clen, t.clen));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// One table of exact_pair_records: the target's section (<= 16 entries) is held in registers; the
// candidate's entries are visited in ascending order, each matches at most one target entry, and the
// products are added in that order -- the same sequence of operations as the mllib merge.
template <bool TRACK_TAIL>
__device__ __forceinline__ double records_dot(const int *__restrict__ cc, const double *__restrict__ cv, int nc,
                                              const int *__restrict__ tc_, const double *__restrict__ tv_, int nt,
                                              unsigned tail_bit, int &min_tail, bool F32) {
    unsigned tcol[16];
    double tval[16];
#pragma unroll
    for (int f0 = 0; f0 < 16; f0 += 4) {
        unsigned c4[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        double x4[4] = {0.0, 0.0, 0.0, 0.0};
        if (f0 < nt) rec_load4(tc_, tv_, f0, c4, x4, F32);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            tcol[f0 + f] = (f0 + f < nt) ? (c4[f] & REC_COL_MASK) : 0xffffffffu;
            tval[f0 + f] = x4[f];
        }
    }
    double sum = 0.0;
    for (int k0 = 0; k0 < nc; k0 += 4) {
        unsigned c[4];
        double x[4];
        rec_load4(cc, cv, k0, c, x, F32);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (k0 + e < nc) {
                const unsigned ix = c[e] & REC_COL_MASK;
                double m = 0.0;
                bool hit = false;
#pragma unroll
                for (int f = 0; f < 16; ++f) {
                    const bool eq = tcol[f] == ix;
                    m = eq ? tval[f] : m;
                    hit = hit || eq;
                }
                if (hit) {
                    sum = xadd(sum, xmul(x[e], m));
                    if (TRACK_TAIL) {
                        if (min_tail < 0 && (c[e] & tail_bit)) min_tail = (int)ix;
                    }
                }
            }
        }
    }
    return sum;
}

// exact_pair for two persons that both have a packed record and a target with <= 16 entries per table:
// two independent meta loads, then both records with 128-bit loads, the match in registers.
__device__ __forceinline__ bool exact_pair_records(const TileAux &aux, int cand, int tix, double pw, double cw,
                                                   int &min_tail, double &out) {
    min_tail = -1;
    const bool F32 = aux.vals_f32;
    out = 0.0;
    if (cand == tix) return true;
    const bool pr = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    long long z0 = clock64();
    const unsigned long long mc = __ldg(aux.meta + cand), mt = __ldg(aux.meta + tix);
    const int npc = (int)((mc >> 40) & 0xfffu), ncc = (int)(mc >> 52);
    const int npt = (int)((mt >> 40) & 0xfffu), nct = (int)(mt >> 52);
    if (npt > 16 || nct > 16) return false;                         // caller falls back to the general path
    const double *rc = aux.rec + (mc & 0xffffffffffULL), *rt = aux.rec + (mt & 0xffffffffffULL);
    {
        // the sections are read by dependent loads below: start all of the candidate's 128-byte lines
        // now, so that only the first of those loads pays the HBM latency
        const int words = 2 + rec_cols_words(npc) + rec_vals_words(npc, F32) + rec_cols_words(ncc) + rec_vals_words(ncc, F32);
        const char *pb = reinterpret_cast<const char *>(rc);
        const int lines = min(8, (int)(((reinterpret_cast<unsigned long long>(pb) & 127ULL) + 8ULL * words + 127ULL) >> 7));
        for (int l = 1; l < lines; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + 128 * l));
    }
    if (pr && (mc | mt) == 0xffffffffffffffffULL) g_probe[7] = 1;   // consume the meta words here
    long long z1 = clock64();
    const double2 lc = __ldg(reinterpret_cast<const double2 *>(rc)), lt = __ldg(reinterpret_cast<const double2 *>(rt));
    if (pr && lc.x + lt.x == -1.0) g_probe[7] = 2;                  // consume the headers here
    long long z2 = clock64();
    const int *pcc = reinterpret_cast<const int *>(rc + 2);
    const double *pvc = rc + 2 + rec_cols_words(npc);
    const int *ccc = reinterpret_cast<const int *>(pvc + rec_vals_words(npc, F32));
    const double *cvc = pvc + rec_vals_words(npc, F32) + rec_cols_words(ncc);
    const int *pct = reinterpret_cast<const int *>(rt + 2);
    const double *pvt = rt + 2 + rec_cols_words(npt);
    const int *cct = reinterpret_cast<const int *>(pvt + rec_vals_words(npt, F32));
    const double *cvt = pvt + rec_vals_words(npt, F32) + rec_cols_words(nct);
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (npc > 0) {
        double sum = records_dot<true>(pcc, pvc, npc, pct, pvt, npt, aux.tail_bit, min_tail, F32);
        double c = xdiv(sum, xmul(lc.x, lt.x));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    if (ncc > 0) {
        int dummy = 0;
        double sum = records_dot<false>(ccc, cvc, ncc, cct, cvt, nct, 0u, dummy, F32);
        double c = xdiv(sum, xmul(lc.y, lt.y));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (keep) out = xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
    if (pr) {
        long long z3 = clock64();
        g_probe[0] += (unsigned long long)(z1 - z0);
        g_probe[1] += (unsigned long long)(z2 - z1);
        g_probe[2] += (unsigned long long)(z3 - z2);
        g_probe[3] += 1;
    }
    return true;
}

// exact_pair_sig on the candidate's compact record; the target side (its own record: columns for the binary
// search when they are not in shared memory, values) stays on the full-size records, which the 128 targets of a
// CTA keep L1/L2-resident.
__device__ __forceinline__ double exact_pair_sig_compact(const TileAux &aux, int cand, int tix, unsigned long long tsig,
                                                         const double *__restrict__ tdense, const int *tcols_s,
                                                         double pw, double cw, int &min_tail) {
    min_tail = -1;
    const bool F32 = aux.vals_f32;
    if (cand == tix) return 0.0;
    const unsigned long long pol = policy_evict_last();
    CRec rc;
    rc.open(aux, __ldg(aux.cmeta + cand), pol);
    const unsigned long long mt = __ldg(aux.meta + tix);
    const int npt = (int)((mt >> 40) & 0xfffu);
    const double *rt = aux.rec + (mt & 0xffffffffffULL);
    const double tplen = __ldg(aux.plen + tix), tclen = __ldg(aux.clen + tix);
    const int *tpc = reinterpret_cast<const int *>(rt + 2);
    const double *tpv = rt + 2 + rec_cols_words(npt);
    const int np = rc.np, nc = rc.nc;
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (np > 0) {
        double sum = 0.0;
        unsigned sq = 0u;
        // the target's entry for a candidate entry that passed the signature test (binary search), and its product
        auto place_match = [&](unsigned craw) {
            const int ix = (int)(craw & CREC_COL_MASK);
            const unsigned v = (craw >> 22) & 0xffu;
            {
                int lo = 0, hi = npt;                                // first index with target col >= ix
                bool match;
                if (tcols_s) {                                       // the target's columns are in shared memory
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if (tcols_s[mid] < ix) lo = mid + 1; else hi = mid;
                    }
                    match = lo < npt && tcols_s[lo] == ix;
                } else {
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if ((int)((unsigned)__ldg(tpc + mid) & REC_COL_MASK) < ix) lo = mid + 1; else hi = mid;
                    }
                    match = lo < npt && (int)((unsigned)__ldg(tpc + lo) & REC_COL_MASK) == ix;
                }
                if (match) {
                    sum = xadd(sum, xmul((double)v, rec_val(tpv, lo, F32)));
                    if (min_tail < 0 && (craw & aux.tail_bit)) min_tail = ix;
                }
            }
        };
        auto place_term = [&](unsigned craw) {
            const unsigned v = (craw >> 22) & 0xffu;
            sq += v * v;
            if ((tsig >> sig_bit((int)(craw & CREC_COL_MASK))) & 1ULL) place_match(craw);
        };
        // Pass 1 (no memory): squares and the signature test of the prefetched entries -> a bit mask.  Pass 2: each
        // lane searches only ITS entries that passed, in ascending order (the order the products are added in).
        // With the search inside the unrolled entry slots the warp ran it in nearly every slot for a few lanes
        // (as in the postings evaluator, profiles/r2_knn_hot_lines.txt); now as many rounds as its busiest lane has hits.
        unsigned pend = 0u;
#pragma unroll
        for (int j = 0; j < CREC_PW; ++j) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (2 * j + hf < np) {
                    const unsigned craw = hf ? (unsigned)(rc.pw[j] >> 32) : (unsigned)rc.pw[j];
                    const unsigned v = (craw >> 22) & 0xffu;
                    sq += v * v;
                    if ((tsig >> sig_bit((int)(craw & CREC_COL_MASK))) & 1ULL) pend |= 1u << (2 * j + hf);
                }
            }
        }
        while (pend) {
            const int e = __ffs(pend) - 1;
            pend &= pend - 1;
            const unsigned long long wd = ld_rec_u64(rc.r + (e >> 1), pol);
            place_match((e & 1) ? (unsigned)(wd >> 32) : (unsigned)wd);
        }
        for (int j = CREC_PW; j < rc.pwords; ++j) {                  // long records: the rest, entry by entry
            const unsigned long long wd = ld_rec_u64(rc.r + j, pol);
            place_term((unsigned)wd);
            if (2 * j + 1 < np) place_term((unsigned)(wd >> 32));
        }
        const double c = xdiv(sum, xmul(sqrt((double)sq), tplen));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    if (nc > 0) {
        double sum = 0.0;
        unsigned sq = 0u;
        auto cat_term = [&](unsigned h) {
            const unsigned v = (h >> 6) & 0x3ffu;
            sq += v * v;
            sum = xadd(sum, xmul((double)v, __ldg(tdense + (h & 63u))));
        };
#pragma unroll
        for (int j = 0; j < CREC_CW; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * j + q < nc) cat_term((unsigned)(rc.cw[j] >> (16 * q)) & 0xffffu);
        }
        for (int j = CREC_CW; j < crec_cat_words(nc); ++j) {
            const unsigned long long wd = ld_rec_u64(rc.r + rc.pwords + j, pol);
            for (int q = 0; q < 4; ++q)
                if (4 * j + q < nc) cat_term((unsigned)(wd >> (16 * q)) & 0xffffu);
        }
        const double c = xdiv(sum, xmul(sqrt((double)sq), tclen));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// exact_pair for two persons with packed records, the target described by a 64-bit signature of its
// places and its category vector as a dense row (both prepared once per CTA): the same operations in the
// same order as exact_pair_staged, i.e. as the mllib merge, with ~10x fewer instructions than matching
// every candidate entry against every target entry.
//  * category dot: sum over the candidate's entries (ascending) of x * dense[col]; entries the target
//    lacks contribute x * 0.0 = +0.0 and leave the sum unchanged (values are non-negative on this path);
//  * place dot: an entry can only match if its signature bit is set; then a binary search over the
//    target's own record finds it.
__device__ __forceinline__ double exact_pair_sig(const TileAux &aux, int cand, int tix, unsigned long long tsig,
                                                 const double *__restrict__ tdense, const int *tcols_s, double pw,
                                                 double cw, int &min_tail) {
    if (aux.cmeta) return exact_pair_sig_compact(aux, cand, tix, tsig, tdense, tcols_s, pw, cw, min_tail);
    min_tail = -1;
    const bool F32 = aux.vals_f32;
    if (cand == tix) return 0.0;
    const bool pr = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    long long z0 = clock64();
    const unsigned long long mc = __ldg(aux.meta + cand), mt = __ldg(aux.meta + tix);
    const int np = (int)((mc >> 40) & 0xfffu), nc = (int)(mc >> 52);
    const int npt = (int)((mt >> 40) & 0xfffu);
    const double *r = aux.rec + (mc & 0xffffffffffULL), *rt = aux.rec + (mt & 0xffffffffffULL);
    {
        // the sections are read by dependent loads below: start all of the candidate's 128-byte lines now
        const int words = 2 + rec_cols_words(np) + rec_vals_words(np, F32) + rec_cols_words(nc) + rec_vals_words(nc, F32);
        const char *pb = reinterpret_cast<const char *>(r);
        const int lines = min(8, (int)(((reinterpret_cast<unsigned long long>(pb) & 127ULL) + 8ULL * words + 127ULL) >> 7));
        for (int l = 1; l < lines; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + 128 * l));
    }
    long long z1 = clock64();
    const double2 lens = __ldg(reinterpret_cast<const double2 *>(r)), lt = __ldg(reinterpret_cast<const double2 *>(rt));
    if (pr && lens.x + lt.x == -1.0) g_probe[5] = 2;                // consume the headers here
    long long z2 = clock64();
    const long long zh = z2;
    const int *pc = reinterpret_cast<const int *>(r + 2);
    const double *pv = r + 2 + rec_cols_words(np);
    const int *cc = reinterpret_cast<const int *>(pv + rec_vals_words(np, F32));
    const double *cvp = pv + rec_vals_words(np, F32) + rec_cols_words(nc);
    const int *tpc = reinterpret_cast<const int *>(rt + 2);
    const double *tpv = rt + 2 + rec_cols_words(npt);
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (np > 0) {
        double sum = 0.0;
        for (int k0 = 0; k0 < np; k0 += 4) {
            unsigned c[4];
            double x[4];
            rec_load4(pc, pv, k0, c, x, F32);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (k0 + e < np) {
                    const int ix = (int)(c[e] & REC_COL_MASK);
                    if ((tsig >> sig_bit(ix)) & 1ULL) {
                        int lo = 0, hi = npt;                        // first index with target col >= ix
                        bool match;
                        if (tcols_s) {                               // the target's columns are in shared memory
                            while (lo < hi) {
                                int mid = (lo + hi) >> 1;
                                if (tcols_s[mid] < ix) lo = mid + 1; else hi = mid;
                            }
                            match = lo < npt && tcols_s[lo] == ix;
                        } else {
                            while (lo < hi) {
                                int mid = (lo + hi) >> 1;
                                if ((int)((unsigned)__ldg(tpc + mid) & REC_COL_MASK) < ix) lo = mid + 1; else hi = mid;
                            }
                            match = lo < npt && (int)((unsigned)__ldg(tpc + lo) & REC_COL_MASK) == ix;
                        }
                        if (match) {
                            sum = xadd(sum, xmul(x[e], rec_val(tpv, lo, F32)));
                            if (min_tail < 0 && (c[e] & aux.tail_bit)) min_tail = ix;
                        }
                    }
                }
            }
        }
        double c = xdiv(sum, xmul(lens.x, lt.x));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    if (pr) {
        long long zz = clock64();
        g_probe[2] += (unsigned long long)(zz - z2);                  // place section + matching
        z2 = zz;
    }
    if (nc > 0) {
        double sum = 0.0;
        for (int k0 = 0; k0 < nc; k0 += 4) {
            unsigned c[4];
            double x[4];
            rec_load4(cc, cvp, k0, c, x, F32);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k0 + e < nc) sum = xadd(sum, xmul(x[e], __ldg(tdense + c[e])));
        }
        double c = xdiv(sum, xmul(lens.y, lt.y));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (pr) {
        long long z3 = clock64();
        g_probe[0] += (unsigned long long)(z1 - z0);
        g_probe[1] += (unsigned long long)(zh - z1);
        g_probe[7] += (unsigned long long)(z3 - z2);                  // category section + dense row + divisions
        g_probe[3] += 1;
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// exact combined similarity + the smallest shared tail place (-1 if none)
__device__ __forceinline__ double exact_pair(const KnnDev &d, const TileAux &aux, long long i, const TargetRows &t,
                                             double pw, double cw, int &min_tail) {
    if (aux.meta) return exact_pair_packed(aux, i, t, pw, cw, min_tail);
    min_tail = -1;
    if (i == t.t) return 0.0;
    // round trip 1: row extents and lengths of both tables (independent loads)
    const int ps = __ldg(d.place.rowptr + i), pe = __ldg(d.place.rowptr + i + 1);
    const int cs0 = __ldg(d.cat.rowptr + i), ce = __ldg(d.cat.rowptr + i + 1);
    const double plen = __ldg(d.place.len + i), clen = __ldg(d.cat.len + i);
    bool keep = false;
    double ps_sim = 0.0, cs_sim = 0.0;
    if (pe > ps) {
        double sum = bulk_dot<true>(d.place, ps, pe - ps, t.pcol, t.pval, t.pn, aux.head_slot, min_tail);
        double c = xdiv(sum, xmul(plen, t.plen));
        if (c > 0) {
            keep = true;
            ps_sim = c;
        }
    }
    if (ce > cs0) {
        int dummy = 0;
        double sum = bulk_dot<false>(d.cat, cs0, ce - cs0, t.ccol, t.cval, t.cn, nullptr, dummy);
        double c = xdiv(sum, xmul(clen, t.clen));
        if (c > 0) {
            keep = true;
            cs_sim = c;
        }
    }
    if (!keep) return 0.0;
    return xadd(xmul(ps_sim, pw), xmul(cs_sim, cw));
}

// event counters of the tiled kernel (rare events only): [0] exact evaluations from the
// postings pass, [1] exact evaluations of dense-filter survivors, [2] heap insert attempts,
// [3] survivors evaluated inline because the queue was full
__device__ unsigned long long g_tile_stats[4];
// phase cycle counters of knn_tc_kernel, block (0,0) thread 0 of the main pass:
// [0] wait for the B tile, [1] MMA issue + completion wait, [2] TMEM epilogue, [3] barriers + queue drain,
// [4] postings pass, [5] tiles
__device__ unsigned long long g_tc_cycles[12];
// per-block cycles of the main pass: [0][b] dense phase, [1][b] postings phase
__device__ unsigned long long g_tc_block_cycles[2][1024];

struct TileSmem {
    float *thr;                 // filter threshold of target slot t at thr[t * thr_stride]
    int thr_stride;
    float *tvec;                // [T][TILE_TVEC_STRIDE]; [32] = filter threshold
    double *hsim;               // [T][K] heaps, worst at the root
    int *hidx;                  // [T][K]
    int *hcnt;                  // [T]
    int *lock;                  // [T]
    int *tid_of;                // [T] person index of the target or -1
    unsigned long long *queue;  // [TILE_QCAP]  (t << 32 | candidate)
    int *qn;
    unsigned int *stats;        // [4] per-block event counters, flushed to g_tile_stats at the end
    unsigned long long *tsig = nullptr;   // [T] 64-bit signature of the target's places (knn_tc_ws_kernel)
    int t_base = 0;                       // batch position of target slot 0 (row of aux.tdense)
    const int *tpool = nullptr;           // place columns of the targets (ascending, flags stripped), packed
    const unsigned short *toff = nullptr; // [T] offset of target t's columns in tpool, 0xffff = not staged
};

__device__ __forceinline__ bool nb_worse(double sa, int ia, double sb, int ib) {
    return sa < sb || (sa == sb && ia > ib);
}

// insert (sim, idx) into the heap of target slot t (under its lock)
__device__ __forceinline__ void tile_heap_insert(const TileSmem &sm, int t, int K, double sim, int idx) {
    // inside the lock the heap is read and written with plain accesses (the fences after the acquire and before the
    // release order them against the other holders; `volatile` generic pointers made every access a strong
    // system-scope load and every child was read twice)
    double *hs = sm.hsim + (size_t)t * K;
    int *hi = sm.hidx + (size_t)t * K;
    volatile int *cntp = sm.hcnt + t;
    volatile float *thr = sm.thr + (size_t)t * sm.thr_stride;
    bool done = false;
    while (!done) {
        if (atomicCAS(sm.lock + t, 0, 1) == 0) {
            __threadfence_block();
            int n = *cntp;
            if (n < K) {
                int pos = n;
                while (pos > 0) {                       // sift up: parents must be worse
                    int par = (pos - 1) >> 1;
                    double ps_ = hs[par];
                    int pi_ = hi[par];
                    if (nb_worse(sim, idx, ps_, pi_)) {
                        hs[pos] = ps_;
                        hi[pos] = pi_;
                        pos = par;
                    } else {
                        break;
                    }
                }
                hs[pos] = sim;
                hi[pos] = idx;
                *cntp = n + 1;
                if (VREC_WS_ABLATE != 3 && n + 1 == K) *thr = fmaxf(*thr, __double2float_rd(hs[0]));
            } else if (nb_worse(hs[0], hi[0], sim, idx)) {
                int pos = 0;
                for (;;) {                              // sift down from the root
                    const int l = 2 * pos + 1, r = l + 1;
                    int w = -1;
                    double ws = sim;
                    int wi = idx;
                    if (l < K) {
                        const double ls = hs[l];
                        const int li = hi[l];
                        if (nb_worse(ls, li, ws, wi)) {
                            w = l;
                            ws = ls;
                            wi = li;
                        }
                    }
                    if (r < K) {
                        const double rs = hs[r];
                        const int ri = hi[r];
                        if (nb_worse(rs, ri, ws, wi)) {
                            w = r;
                            ws = rs;
                            wi = ri;
                        }
                    }
                    if (w < 0) break;
                    hs[pos] = ws;
                    hi[pos] = wi;
                    pos = w;
                }
                hs[pos] = sim;
                hi[pos] = idx;
                if (VREC_WS_ABLATE != 3) *thr = fmaxf(*thr, __double2float_rd(pos == 0 ? sim : hs[0]));
            }
            __threadfence_block();
            atomicExch(sm.lock + t, 0);
            done = true;
        }
    }
}

// exact evaluation of one filter survivor; from_postings = the place whose postings produced it
__device__ __forceinline__ void tile_process_rows(const KnnDev &d, const TileAux &aux, const TileSmem &sm, int t, int c,
                                                  int K, double pw, double cw, int from_postings,
                                                  const TargetRows &tr, double thr0 = 0.0) {
    atomicAdd(sm.stats + (from_postings >= 0 ? 0 : 1), 1u);
    int min_tail;
    double sim = exact_pair(d, aux, c, tr, pw, cw, min_tail);
    if (!(sim > 0) || sim < thr0) return;      // >= K candidates are known to reach thr0
    if (from_postings >= 0) {
        if (min_tail != from_postings) return;          // counted at its smallest shared tail place
    } else if (from_postings == -1 && min_tail >= 0) {
        return;                                         // the postings pass owns this pair
    }                                                   // (-2: seed pass, every pair counts)
    // cheap pre-check without the lock: the root similarity is written once per update and only
    // grows, so a stale read can only let too much through (the exact test is under the lock)
    volatile double *hs = sm.hsim + (size_t)t * K;
    if (*(volatile int *)(sm.hcnt + t) >= K && sim < hs[0]) return;
    atomicAdd(sm.stats + 2, 1u);
    tile_heap_insert(sm, t, K, sim, c);
}

// postings-pass survivor with a staged target: counted at its smallest shared tail place
__device__ __forceinline__ void tile_process_staged(const TileAux &aux, const TileSmem &sm, int t, int c, int K,
                                                    double pw, double cw, int from_postings, const StagedTarget &st,
                                                    double thr0 = 0.0) {
    atomicAdd(sm.stats + 0, 1u);
    int min_tail;
    // what can still enter: thr0 (>= K candidates are known to reach it) and, once the heap is full, its root
    double thr_eff = thr0;
    if (*(volatile int *)(sm.hcnt + t) >= K) thr_eff = fmax(thr_eff, *(volatile double *)(sm.hsim + (size_t)t * K));
    double sim = exact_pair_staged(aux, c, st, pw, cw, min_tail, thr_eff);
    if (!(sim > 0) || sim < thr0) return;      // >= K candidates are known to reach thr0
    if (min_tail != from_postings) return;
    volatile double *hs = sm.hsim + (size_t)t * K;
    if (*(volatile int *)(sm.hcnt + t) >= K && sim < hs[0]) return;
    atomicAdd(sm.stats + 2, 1u);
    tile_heap_insert(sm, t, K, sim, c);
}

// Survivors of knn_tc_ws_kernel's dense filter, one per lane of a converged warp (t < 0: none):
// signature / dense-row evaluation, then the heap inserts in rounds -- in every round the lanes that
// still hold a result elect one lane per distinct target (__match_any_sync), so the lanes of a warp
// never spin on the same heap lock (a warp's survivors all belong to the 32 targets of its lane quarter).
// One copy of the code (not inlined): the kernel stays small enough for the instruction cache.
__device__ __noinline__ void tile_process_sig(const TileAux &aux, const TileSmem &sm, int t, int c, int K, double pw,
                                              double cw) {
    double sim = 0.0;
    bool pending = false, evaluated = false;
    if (t >= 0) {
        const int tix = sm.tid_of[t];
        if (tix >= 0) {
            int min_tail;
            const unsigned short off = sm.toff[t];
            sim = exact_pair_sig(aux, c, tix, sm.tsig[t], aux.tdense + (size_t)(sm.t_base + t) * aux.cat_dim,
                                 off == 0xffffu ? nullptr : sm.tpool + off, pw, cw, min_tail);
            evaluated = true;
            pending = sim > 0 && min_tail < 0;      // pairs sharing a tail place belong to the postings kernel
        }
    }
    __syncwarp();
    // debug counters, one atomic per warp (32 atomics on one shared-memory word were ~250 cycles of every batch)
    const unsigned n_eval = __popc(__ballot_sync(0xffffffffu, evaluated));
    const unsigned n_ins = __popc(__ballot_sync(0xffffffffu, pending));      // an upper bound: some find the heap moved on
    if ((threadIdx.x & 31) == 0) {
        if (n_eval) atomicAdd(sm.stats + 1, n_eval);
        if (n_ins) atomicAdd(sm.stats + 2, n_ins);
    }
    for (;;) {
        if (pending) {
            volatile double *hs = sm.hsim + (size_t)t * K;
            if (*(volatile int *)(sm.hcnt + t) >= K && sim < hs[0]) pending = false;   // cannot enter any more
        }
        const unsigned todo = __ballot_sync(0xffffffffu, pending);
        if (!todo) break;
        if (pending) {
            const unsigned same = __match_any_sync(todo, t);
            if ((int)(__ffs(same) - 1) == (int)(threadIdx.x & 31)) {
                tile_heap_insert(sm, t, K, sim, c);
                pending = false;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void tile_process(const KnnDev &d, const TileAux &aux, const TileSmem &sm, int t, int c,
                                             int K, double pw, double cw, int from_postings) {
    int tix = sm.tid_of[t];
    if (tix < 0) return;
    if (aux.meta) {
        // both persons have packed records: no dependent loads, no data-dependent merge loops
        int min_tail;
        double sim;
        if (exact_pair_records(aux, c, tix, pw, cw, min_tail, sim)) {
            atomicAdd(sm.stats + (from_postings >= 0 ? 0 : 1), 1u);
            if (!(sim > 0)) return;
            if (from_postings >= 0) {
                if (min_tail != from_postings) return;
            } else if (from_postings == -1 && min_tail >= 0) {
                return;
            }
            volatile double *hs = sm.hsim + (size_t)t * K;
            if (*(volatile int *)(sm.hcnt + t) >= K && sim < hs[0]) return;
            atomicAdd(sm.stats + 2, 1u);
            long long h0 = clock64();
            tile_heap_insert(sm, t, K, sim, c);
            if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
                g_probe[4] += (unsigned long long)(clock64() - h0);
                g_probe[5] += 1;
            }
            return;
        }
    }
    TargetRows tr = load_target(d, tix);
    tile_process_rows(d, aux, sm, t, c, K, pw, cw, from_postings, tr);
}

__global__ void __launch_bounds__(TILE_THREADS, 2)
knn_tile_kernel(KnnDev d, TileAux aux, const int *__restrict__ tidx, int n_targets, int T, int K, int S,
                int cat_dim, double pw, double cw, Nb *__restrict__ part, int *__restrict__ part_cnt,
                long long cand_stride, long long cand_count, int seed_mode, double *__restrict__ seed_thr,
                int part_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, sp = blockIdx.y;
    const int t0 = tile * T;
    const int nt = min(T, n_targets - t0);
    TileSmem sm;
    {
        unsigned char *p = smem_raw;
        sm.tvec = (float *)p;                       p += sizeof(float) * (size_t)T * TILE_TVEC_STRIDE;   // 144*T: 16-B aligned rows
        sm.hsim = (double *)p;                      p += sizeof(double) * (size_t)T * K;
        sm.queue = (unsigned long long *)p;         p += sizeof(unsigned long long) * TILE_QCAP;
        sm.hidx = (int *)p;                         p += sizeof(int) * (size_t)T * K;
        sm.hcnt = (int *)p;                         p += sizeof(int) * T;
        sm.lock = (int *)p;                         p += sizeof(int) * T;
        sm.tid_of = (int *)p;                       p += sizeof(int) * T;
        sm.qn = (int *)p;
        sm.thr = sm.tvec + TILE_D;
        sm.thr_stride = TILE_TVEC_STRIDE;
    }
    __shared__ unsigned int s_stats[4];
    sm.stats = s_stats;
    if (tid < 4) s_stats[tid] = 0;
    // candidates are j * cand_stride for j in [jlo, jhi); the main pass has stride 1 (all persons),
    // the seed pass a strided sample.  [lo, hi) is the same range in person indices.
    const long long jlo = cand_count * sp / S, jhi = cand_count * (sp + 1) / S;
    const long long lo = jlo * cand_stride, hi = seed_mode == 1 ? d.P : jhi * cand_stride;
    // ---- target vectors (weights folded in), thresholds, empty heaps
    for (int t = tid; t < T; t += TILE_THREADS) {
        int tix = t < nt ? tidx[t0 + t] : -1;
        sm.tid_of[t] = tix;
        sm.hcnt[t] = 0;
        sm.lock[t] = 0;
        // invalid slots never pass the filter
        // the seed pass left a lower bound of the K-th best similarity (0 when it found < K)
        float thr0 = (seed_mode != 1 && tix >= 0) ? __double2float_rd(seed_thr[t0 + t]) : 0.0f;
        sm.tvec[(size_t)t * TILE_TVEC_STRIDE + TILE_D] = tix >= 0 ? thr0 : 3.0e38f;
    }
    if (tid == 0) *sm.qn = 0;
    __syncthreads();
    for (int e = tid; e < T * TILE_D; e += TILE_THREADS) {
        int t = e / TILE_D, dd = e % TILE_D;
        int tix = sm.tid_of[t];
        float v = 0.0f;
        if (tix >= 0) {
            double f = (double)aux.feat[(size_t)dd * aux.fstride + tix];
            v = (float)(f * (dd < cat_dim ? cw : pw));   // dims [0, cat_dim): category, rest: head places
        }
        sm.tvec[(size_t)t * TILE_TVEC_STRIDE + dd] = v;
    }
    __syncthreads();
    // ---- 1. dense filter over the candidate range (establishes strong thresholds first)
    for (long long base = jlo; base < jhi; base += TILE_THREADS) {
        long long c = (base + tid) * cand_stride;
        if (base + tid < jhi) {
            float f[TILE_D];
#pragma unroll
            for (int dd = 0; dd < TILE_D; ++dd) f[dd] = __ldg(aux.feat + (size_t)dd * aux.fstride + c);
            for (int t = 0; t < nt; ++t) {
                const float4 *tv = reinterpret_cast<const float4 *>(sm.tvec + (size_t)t * TILE_TVEC_STRIDE);
                float u0 = 0.0f, u1 = 0.0f, u2 = 0.0f, u3 = 0.0f;     // 4 independent FMA chains
#pragma unroll
                for (int q = 0; q < TILE_D / 4; ++q) {
                    float4 a = tv[q];
                    u0 = __fmaf_rn(a.x, f[4 * q + 0], u0);
                    u1 = __fmaf_rn(a.y, f[4 * q + 1], u1);
                    u2 = __fmaf_rn(a.z, f[4 * q + 2], u2);
                    u3 = __fmaf_rn(a.w, f[4 * q + 3], u3);
                }
                float u = (u0 + u1) + (u2 + u3);
                float thr = sm.tvec[(size_t)t * TILE_TVEC_STRIDE + TILE_D];
                if (u + TILE_MARGIN >= thr) {
                    int pos = atomicAdd(sm.qn, 1);
                    if (pos < TILE_QCAP) {
                        sm.queue[pos] = ((unsigned long long)t << 32) | (unsigned long long)(unsigned)c;
                    } else {
                        atomicAdd(sm.stats + 3, 1u);
                        tile_process(d, aux, sm, t, (int)c, K, pw, cw, seed_mode == 1 ? -2 : -1);   // queue full: now
                    }
                }
            }
        }
        __syncthreads();
        int qn = *sm.qn;
        __syncthreads();
        if (qn >= TILE_QCAP / 2 || base + TILE_THREADS >= jhi) {                // block-uniform
            int m = min(qn, TILE_QCAP);
            for (int i = tid; i < m; i += TILE_THREADS) {
                unsigned long long e = sm.queue[i];
                tile_process(d, aux, sm, (int)(e >> 32), (int)(unsigned)(e & 0xffffffffu), K, pw, cw,
                             seed_mode == 1 ? -2 : -1);
            }
            __syncthreads();
            if (tid == 0) *sm.qn = 0;
            __syncthreads();
        }
    }
    // ---- 2. pairs sharing a tail place, through the postings.  One warp per target; the
    // postings sub-ranges of 32 place entries at a time are flattened over the lanes.
    for (int t = warp; t < nt && seed_mode == 0; t += TILE_THREADS / 32) {
        int tix = sm.tid_of[t];
        if (tix < 0) continue;                                                  // warp-uniform
        int ps = d.place.rowptr[tix], pn = d.place.rowptr[tix + 1] - ps;
        for (int e0 = 0; e0 < pn; e0 += 32) {
            int e = e0 + lane, start = 0, len = 0, pl = -1;
            if (e < pn) {
                pl = d.place.col[ps + e];
                if (aux.head_slot[pl] < 0) {
                    int b = aux.pcp[pl], en = aux.pcp[pl + 1];
                    int l0 = b, h0 = en;                 // postings (ascending person) within [lo, hi)
                    while (l0 < h0) {
                        int mid = (l0 + h0) >> 1;
                        if (aux.pper[mid] < lo) l0 = mid + 1; else h0 = mid;
                    }
                    int l1 = l0, h1 = en;
                    while (l1 < h1) {
                        int mid = (l1 + h1) >> 1;
                        if (aux.pper[mid] < hi) l1 = mid + 1; else h1 = mid;
                    }
                    start = l0;
                    len = l1 - l0;
                }
            }
            int incl = len;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - len;
            for (int j0 = 0; j0 < total; j0 += 32) {
                int j = j0 + lane;
                int L = 0;                               // smallest lane whose inclusive prefix exceeds j
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    int probe = __shfl_sync(0xffffffffu, incl, L + step - 1);
                    if (probe <= j) L += step;
                }
                L = min(L, 31);
                int ex_l = __shfl_sync(0xffffffffu, excl, L);
                int st_l = __shfl_sync(0xffffffffu, start, L);
                int pl_l = __shfl_sync(0xffffffffu, pl, L);
                if (j < total) tile_process(d, aux, sm, t, aux.pper[st_l + (j - ex_l)], K, pw, cw, pl_l);
            }
        }
    }
    __syncthreads();
    if (seed_mode == 1) {
        // K real candidates with similarity >= root exist, so the K-th best overall is >= root
        for (int t = tid; t < nt; t += TILE_THREADS) seed_thr[t0 + t] = sm.hcnt[t] >= K ? sm.hsim[(size_t)t * K] : 0.0;
        return;
    }
    if (tid < 4) atomicAdd(&g_tile_stats[tid], (unsigned long long)s_stats[tid]);
    // ---- emit the heaps (unsorted; the merge kernel sorts)
    for (int t = 0; t < nt; ++t) {
        int cnt = sm.hcnt[t];
        Nb *out = part + ((size_t)(t0 + t) * part_stride + sp) * K;
        for (int j = tid; j < cnt; j += TILE_THREADS) {
            Nb e;
            e.sim = sm.hsim[(size_t)t * K + j];
            e.idx = sm.hidx[(size_t)t * K + j];
            e.pad = 0;
            out[j] = e;
        }
        if (tid == 0) part_cnt[(t0 + t) * part_stride + sp] = cnt;
    }
}


// ---------------------------------------------------------------------------------------
// Tensor-core variant of the tiled batch kernel (tcgen05 / TMEM, fp16 features, D = 128).
//
// Operand tiles use the SWIZZLE_128B K-major layout of vrec_tc.cuh (un-swizzled tiles made the
// tensor core's shared-memory reads ~6x slower).
// Same algorithm as knn_tile_kernel, but the dense bound U(t, c) of a 128-target x 128-candidate
// tile is ONE tcgen05.mma chain (8 instructions of K = 16) into 128 TMEM columns, and the filter
// `U * (1 + 2e-3) + 2e-5 >= K-th best` is applied to the accumulators as they come out of TMEM.
// With 128 dims the head covers the 128 - cat_dim most visited places, so far fewer pairs are left
// to the postings pass.  fp16 rounding (2 x 2^-11 relative) only loosens the filter; survivors
// are evaluated in fp64 exactly as in the other kernels, so results stay bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int TC_D = 128;
constexpr int TC_M = 128;
constexpr int TC_N = 128;
constexpr int TC_THREADS = 512;                          // one CTA per SM
constexpr int TC_MAX_K = 56;                             // heaps of 128 targets must fit shared memory
constexpr int TC_TILE_BYTES = TC_M * TC_D * 2;          // 32 KB per operand tile
constexpr int TC_LBO = TC_M * 16;                        // next k-chunk
constexpr int TC_SBO = 128;                              // next 8 rows
constexpr int TC_STAGES = 3;                             // B tiles in flight
constexpr int TC_QCAP = 1024;                            // survivor queue entries (fits next to 3 B stages)

// The first tcgen05 filter kernel (no warp specialisation, cp.async loads, row-major fp16 features) is superseded
// by knn_tc_ws_kernel and compiled only on request (-DVREC_WITH_TC_BASELINE=1, A/B measurements): without it the
// region-set does not carry the 256 B / person row-major feature copy either, and knn_kernel = 3 and the
// "tc_seed" option select the warp-specialised kernel / nothing.
#ifndef VREC_WITH_TC_BASELINE
#define VREC_WITH_TC_BASELINE 0
#endif
#if VREC_WITH_TC_BASELINE
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    unsigned bytes = valid ? 16u : 0u;                   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(KnnDev d, TileAux aux, const __half *__restrict__ feat16, const int *__restrict__ tidx,
              int n_targets, int K, int S, int cat_dim, double pw, double cw, Nb *__restrict__ part,
              int *__restrict__ part_cnt, long long cand_stride, long long cand_count, int seed_mode,
              double *__restrict__ seed_thr, int part_stride) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_m = blockIdx.x, sp = blockIdx.y;
    const int t0 = tile_m * TC_M;
    const int nt = min(TC_M, n_targets - t0);
    // operand tiles on a 1 KB boundary (the descriptor start address ignores its low bits)
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = base, *sB0 = base + TC_TILE_BYTES;            // sB: TC_STAGES stages
    TileSmem sm;
    {
        unsigned char *p = base + (1 + TC_STAGES) * TC_TILE_BYTES;
        sm.hsim = (double *)p;                      p += sizeof(double) * (size_t)TC_M * K;
        sm.queue = (unsigned long long *)p;         p += sizeof(unsigned long long) * TC_QCAP;
        sm.hidx = (int *)p;                         p += sizeof(int) * (size_t)TC_M * K;
        sm.thr = (float *)p;                        p += sizeof(float) * TC_M;
        sm.thr_stride = 1;
        sm.tvec = nullptr;
        sm.hcnt = (int *)p;                         p += sizeof(int) * TC_M;
        sm.lock = (int *)p;                         p += sizeof(int) * TC_M;
        sm.tid_of = (int *)p;                       p += sizeof(int) * TC_M;
        sm.qn = (int *)p;
    }
    __shared__ unsigned int s_stats[4];
    sm.stats = s_stats;
    if (tid < 4) s_stats[tid] = 0;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * TC_N);        // two accumulators
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_init(&bar[1], 1);
        tc::mbar_init_fence();
        *sm.qn = 0;
    }
    const long long jlo = cand_count * sp / S, jhi = cand_count * (sp + 1) / S;
    const long long lo = jlo * cand_stride, hi = seed_mode == 1 ? d.P : jhi * cand_stride;
    for (int t = tid; t < TC_M; t += TC_THREADS) {
        int tix = t < nt ? tidx[t0 + t] : -1;
        sm.tid_of[t] = tix;
        sm.hcnt[t] = 0;
        sm.lock[t] = 0;
        float thr0 = (seed_mode != 1 && tix >= 0) ? __double2float_rd(seed_thr[t0 + t]) : 0.0f;
        sm.thr[t] = tix >= 0 ? thr0 : 3.0e38f;
    }
    __syncthreads();
    // candidate tiles are visited in a per-CTA rotated order, so that the CTAs do not all pull the
    // same 32 KB out of L2 at the same moment
    const int ntiles = (int)((jhi - jlo + TC_N - 1) / TC_N);
    const int rot = ntiles > 0 ? (int)(((long long)blockIdx.x * 67 + (long long)blockIdx.y * 29) % ntiles) : 0;
    auto tile_start = [&](int i) {
        int w = i + rot;
        if (w >= ntiles) w -= ntiles;
        return jlo + (long long)w * TC_N;
    };
    // B tile loader: 16-byte chunk c of candidate row r -> chunk-major UMMA layout, zero fill past the end
    auto load_b = [&](int i) {
        if (i < ntiles) {
            const long long tile = tile_start(i);
            unsigned char *sB = sB0 + (size_t)(i % TC_STAGES) * TC_TILE_BYTES;
            for (int q = tid; q < TC_N * (TC_D / 8); q += TC_THREADS) {
                // 16 consecutive lanes fetch one 256-byte feature row (coalesced); the swizzle keeps the
                // shared-memory side conflict-free
                int r = q >> 4, c = q & 15;
                long long j = tile + r;
                bool ok = j < jhi;
                const __half *src = feat16 + (size_t)((ok ? j : jlo) * cand_stride) * TC_D + c * 8;
                cp_async16(sB + tc::sw128_offset(TC_N, r, c), src, ok);
            }
        }
        cp_async_commit();                       // always commit: keeps the group count uniform
    };
    load_b(0);
    load_b(1);
    // ---- A tile: the targets' features with the weights folded in (chunk-major UMMA layout)
    for (int q = tid; q < TC_M * (TC_D / 8); q += TC_THREADS) {
        int r = q % TC_M, c = q / TC_M;
        int tix = sm.tid_of[r];
        __align__(16) __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = __float2half(0.0f);
        if (tix >= 0) {
            uint4 raw = *reinterpret_cast<const uint4 *>(feat16 + (size_t)tix * TC_D + c * 8);
            const __half *src = reinterpret_cast<const __half *>(&raw);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int dd = c * 8 + e;
                h[e] = __float2half(__half2float(src[e]) * (float)(dd < cat_dim ? cw : pw));
            }
        }
        *reinterpret_cast<uint4 *>(sA + tc::sw128_offset(TC_M, r, c)) = *reinterpret_cast<const uint4 *>(h);
    }
    const uint32_t idesc = tc::make_idesc_f16(TC_M, TC_N);
    const uint32_t a_addr = tc::smem_u32(sA), b_addr0 = tc::smem_u32(sB0);
    // issue the MMA chain of tile i into accumulator i & 1 (one thread)
    auto issue_mma = [&](int i, uint32_t tbase_) {
        tc::fence_after_sync();
        const uint32_t b_addr = b_addr0 + (uint32_t)(i % TC_STAGES) * TC_TILE_BYTES;
        const uint32_t acc = tbase_ + (uint32_t)(i & 1) * TC_N;
#pragma unroll
        for (int k = 0; k < TC_D / 16; ++k) {
            uint64_t da = tc::make_desc_sw128(tc::sw128_kstep_addr(a_addr, TC_M, k));
            uint64_t db = tc::make_desc_sw128(tc::sw128_kstep_addr(b_addr, TC_N, k));
            tc::mma_f16(acc, da, db, idesc, k > 0);
        }
        tc::mma_commit(&bar[i & 1]);
    };
    cp_async_wait<1>();                          // tile 0 has landed
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    if (tid == 0 && ntiles > 0) issue_mma(0, tbase);
    const int lq = warp & 3, cq = warp >> 2;                  // TMEM lane quarter, column quarter (16 warps)
    const int my_t = lq * 32 + lane;
    uint32_t phase0 = 0u, phase1 = 0u;
    const bool prof = tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && seed_mode != 1;
    long long tk0 = clock64(), tk1;
    const long long blk_t0 = tk0;
#define TC_TICK(slot)                                                  \
    if (prof) {                                                        \
        tk1 = clock64();                                               \
        g_tc_cycles[slot] += (unsigned long long)(tk1 - tk0);          \
        tk0 = tk1;                                                     \
    }
    // ---- 1. dense filter on the tensor cores.  Iteration i: B(i+1) has landed -> MMA(i+1) is issued
    // into the other accumulator, B(i+2) starts streaming, then the epilogue of tile i runs under them.
    for (int i = 0; i <= ntiles; ++i) {                           // one extra turn drains the last survivors
        const bool live = i < ntiles, has_next = i + 1 < ntiles;
        cp_async_wait<0>();                      // B(i+1) (issued one iteration ago) has landed
        TC_TICK(6)
        tc::fence_proxy_async();
        tc::fence_before_sync();
        // ONE barrier per tile: B(i+1) visible to all, epilogue(i-1) finished, and a block-uniform
        // decision whether the survivor queue must be drained now
        const int drain = __syncthreads_or((*(volatile int *)sm.qn >= TC_QCAP / 2) || !live);
        TC_TICK(7)
        if (tid == 0 && has_next) issue_mma(i + 1, tbase);
        TC_TICK(8)
        if (live) load_b(i + 2);                 // stage (i+2)%3 was last read by MMA(i-1), long complete
        TC_TICK(3)
        if (drain) {
            int m = min(*(volatile int *)sm.qn, TC_QCAP);
            for (int qi = tid; qi < m; qi += TC_THREADS) {
                unsigned long long e = sm.queue[qi];
                tile_process(d, aux, sm, (int)(e >> 32), (int)(unsigned)(e & 0xffffffffu), K, pw, cw,
                             seed_mode == 1 ? -2 : -1);
            }
            __syncthreads();
            if (tid == 0) *sm.qn = 0;
            __syncthreads();
        }
        if (!live) break;
        TC_TICK(0)
        if (prof) g_tc_cycles[5] += 1;
        if (i & 1) {
            tc::mbar_wait(&bar[1], phase1);
            phase1 ^= 1u;
        } else {
            tc::mbar_wait(&bar[0], phase0);
            phase0 ^= 1u;
        }
        tc::fence_after_sync();
        TC_TICK(1)
        const float thr = *(volatile float *)(sm.thr + my_t);
        const long long tile = tile_start(i);
        {
            const int c0 = cq * 32;
            float v[32];
            tc::tmem_ld32(tbase + ((uint32_t)(lq * 32) << 16) + (uint32_t)((i & 1) * TC_N + c0), v);
            unsigned pass = 0;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) pass |= (v[jj] * 1.002f + 2e-5f >= thr ? 1u : 0u) << jj;
            while (pass) {                       // rare: a single copy of the survivor path
                const int jj = __ffs(pass) - 1;
                pass &= pass - 1;
                const long long j = tile + c0 + jj;
                if (j < jhi) {
                    int c = (int)(j * cand_stride);
                    int pos = atomicAdd(sm.qn, 1);
                    if (pos < TC_QCAP) {
                        sm.queue[pos] = ((unsigned long long)my_t << 32) | (unsigned long long)(unsigned)c;
                    } else {
                        atomicAdd(sm.stats + 3, 1u);
                        tile_process(d, aux, sm, my_t, c, K, pw, cw, seed_mode == 1 ? -2 : -1);
                    }
                }
            }
        }
        TC_TICK(2)
    }
    cp_async_wait<0>();
    const long long blk_t1 = clock64();
    __syncthreads();                                               // every MMA and epilogue is done: B stages are free
    // ---- 2. pairs sharing a tail place, through the postings (as in knn_tile_kernel).  The current
    // target's rows are staged in shared memory (the idle B stages): with ~220 KB of shared memory
    // carved out there is hardly any L1 left to keep them hot.
    constexpr int STG_P = 64, STG_C = 32, STG_DENSE = 64;
    constexpr int STG_BYTES = STG_P * 12 + STG_C * 12 + STG_DENSE * 8;
    unsigned char *wbuf = sB0 + (size_t)warp * STG_BYTES;
    double *s_pval = (double *)wbuf, *s_cval = (double *)(wbuf + STG_P * 8);
    double *s_dense = (double *)(wbuf + STG_P * 8 + STG_C * 8);
    int *s_pcol = (int *)(wbuf + STG_P * 8 + STG_C * 8 + STG_DENSE * 8), *s_ccol = s_pcol + STG_P;
    const bool can_stage = aux.meta != nullptr && cat_dim <= STG_DENSE;
    for (int t = warp; t < nt && seed_mode == 0; t += TC_THREADS / 32) {
        int tix = sm.tid_of[t];
        if (tix < 0) continue;
        TargetRows tr = load_target(d, tix);
        __syncwarp();
        StagedTarget stg;
        bool staged = false;
        if (tr.pn <= STG_P && tr.cn <= STG_C) {
            unsigned long long sig = 0ULL;
            for (int e = lane; e < STG_DENSE; e += 32) s_dense[e] = 0.0;
            __syncwarp();
            for (int e = lane; e < tr.pn; e += 32) {
                int col = tr.pcol[e];
                s_pcol[e] = col;
                s_pval[e] = tr.pval[e];
                sig |= 1ULL << sig_bit(col);
            }
            for (int e = lane; e < tr.cn; e += 32) {
                s_ccol[e] = tr.ccol[e];
                s_cval[e] = tr.cval[e];
                if (can_stage) s_dense[tr.ccol[e]] = tr.cval[e];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sig |= __shfl_xor_sync(0xffffffffu, sig, off);
            tr.pcol = s_pcol;
            tr.pval = s_pval;
            tr.ccol = s_ccol;
            tr.cval = s_cval;
            staged = can_stage;
            stg.t = tr.t;
            stg.pn = tr.pn;
            stg.pcol = s_pcol;
            stg.pval = s_pval;
            stg.sig = sig;
            stg.cat_dense = s_dense;
            stg.plen = tr.plen;
            stg.clen = tr.clen;
        }
        __syncwarp();
        int ps = d.place.rowptr[tix], pn = d.place.rowptr[tix + 1] - ps;
        for (int e0 = 0; e0 < pn; e0 += 32) {
            int e = e0 + lane, start = 0, len = 0, pl = -1;
            if (e < pn) {
                pl = d.place.col[ps + e];
                if (aux.head_slot[pl] < 0) {
                    int b = aux.pcp[pl], en = aux.pcp[pl + 1];
                    int l0 = b, h0 = en;
                    while (l0 < h0) {
                        int mid = (l0 + h0) >> 1;
                        if (aux.pper[mid] < lo) l0 = mid + 1; else h0 = mid;
                    }
                    int l1 = l0, h1 = en;
                    while (l1 < h1) {
                        int mid = (l1 + h1) >> 1;
                        if (aux.pper[mid] < hi) l1 = mid + 1; else h1 = mid;
                    }
                    start = l0;
                    len = l1 - l0;
                }
            }
            int incl = len;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - len;
            // Software pipeline over the flattened postings: candidate id + meta word are fetched two
            // iterations ahead and the record of the next iteration is pulled into L2 while the current
            // one is evaluated, so that the evaluation itself only sees L2 latencies.
            auto locate = [&](int j, int &pl_out) -> int {        // postings index of flat position j, or -1
                int L = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    int probe = __shfl_sync(0xffffffffu, incl, L + step - 1);
                    if (probe <= j) L += step;
                }
                L = min(L, 31);
                int ex_l = __shfl_sync(0xffffffffu, excl, L);
                int st_l = __shfl_sync(0xffffffffu, start, L);
                pl_out = __shfl_sync(0xffffffffu, pl, L);
                return j < total ? st_l + (j - ex_l) : -1;
            };
            auto fetch = [&](int j, int &cand_out, unsigned long long &meta_out, int &pl_out) {
                int at = locate(j, pl_out);
                cand_out = at >= 0 ? __ldg(aux.pper + at) : -1;
                meta_out = (cand_out >= 0 && aux.meta) ? __ldg(aux.meta + cand_out) : 0ULL;
            };
            auto prefetch_record = [&](int cand_, unsigned long long m_) {
                if (cand_ >= 0 && aux.meta) {
                    const char *r = reinterpret_cast<const char *>(aux.rec + (m_ & 0xffffffffffULL));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128));
                }
            };
            int c0 = -1, c1 = -1, p0 = -1, p1 = -1;
            unsigned long long m0 = 0, m1 = 0;
            fetch(lane, c0, m0, p0);
            fetch(32 + lane, c1, m1, p1);
            prefetch_record(c0, m0);
            for (int j0 = 0; j0 < total; j0 += 32) {
                int c2, p2;
                unsigned long long m2;
                fetch(j0 + 64 + lane, c2, m2, p2);                 // two iterations ahead
                prefetch_record(c1, m1);                           // next iteration's record -> L2
                const int cand = c0, pl_l = p0;
                if (prof) g_probe_on = 1;
                long long q1 = clock64();
                if (cand >= 0) {
                    if (staged) tile_process_staged(aux, sm, t, cand, K, pw, cw, pl_l, stg);
                    else tile_process_rows(d, aux, sm, t, cand, K, pw, cw, pl_l, tr);
                }
                __syncwarp();
                if (prof) {
                    g_tc_cycles[10] += (unsigned long long)(clock64() - q1);
                    g_tc_cycles[11] += 1;
                    g_probe_on = 0;
                }
                c0 = c1; m0 = m1; p0 = p1;
                c1 = c2; m1 = m2; p1 = p2;
            }
        }
    }
    __syncthreads();
    TC_TICK(4)
#undef TC_TICK
    if (tid < 4) atomicAdd(&g_tile_stats[tid], (unsigned long long)s_stats[tid]);
    if (tid == 0 && seed_mode != 1 && blockIdx.y == 0 && blockIdx.x < 1024) {
        g_tc_block_cycles[0][blockIdx.x] = (unsigned long long)(blk_t1 - blk_t0);
        g_tc_block_cycles[1][blockIdx.x] = (unsigned long long)(clock64() - blk_t1);
    }
    if (seed_mode == 1) {
        for (int t = tid; t < nt; t += TC_THREADS)
            seed_thr[t0 + t] = sm.hcnt[t] >= K ? *(volatile double *)(sm.hsim + (size_t)t * K) : 0.0;
    } else {
        for (int t = 0; t < nt; ++t) {
            int cnt = sm.hcnt[t];
            Nb *out = part + ((size_t)(t0 + t) * part_stride + sp) * K;
            for (int j = tid; j < cnt; j += TC_THREADS) {
                Nb e;
                e.sim = *(volatile double *)(sm.hsim + (size_t)t * K + j);
                e.idx = *(volatile int *)(sm.hidx + (size_t)t * K + j);
                e.pad = 0;
                out[j] = e;
            }
            if (tid == 0) part_cnt[(t0 + t) * part_stride + sp] = cnt;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 2 * TC_N);
}
#endif   // VREC_WITH_TC_BASELINE


// ---------------------------------------------------------------------------------------
// Postings pass as its own kernel: one warp per target, targets handed out dynamically.
// Evaluates exactly every pair (target, candidate) that shares a tail place (each once, at its
// smallest shared tail place) and emits the best K of them as one more partial list (slot S).
// The K-th best similarity the dense kernel already found is a valid lower bound: candidates
// below it are dropped without touching the heap.  Small shared-memory footprint -> twice the
// resident warps of the tensor-core kernel, and no load imbalance between blocks.
// ---------------------------------------------------------------------------------------
constexpr int POST_WARPS = 8;
constexpr int POST_STG_P = 64, POST_STG_C = 32, POST_DENSE = 64;

__host__ __device__ constexpr int post_warp_bytes(int K) {
    return ((8 * K + 8 * POST_STG_P + 8 * POST_STG_C + 8 * POST_DENSE + 4 * K + 4 * POST_STG_P + 4 * POST_STG_C + 64) + 15) / 16 * 16;
}

// Work items of the postings kernel, longest first.  A target's work = the visitors of its tail places;
// a heavy target is split into up to `max_parts` parts (every part takes every max_parts-th block of 32
// candidates and keeps its own heap -> its own partial list), so the kernel does not end on one long
// target.  Item = target * 8 + part * 2... encoded as (target << 4) | (part << 2) | (parts - 1); items are
// binned by floor(log2(work per part)) and handed out from the heaviest bin down.
constexpr int POST_MAX_PARTS = 4;
constexpr int POST_PART_WORK = 6144;     // candidates per part before a target is split further

__global__ void knn_post_work_kernel(KnnDev d, TileAux aux, const int *__restrict__ tidx, int n_targets, int max_parts,
                                     int *__restrict__ bin_of, int *__restrict__ parts_of, int *__restrict__ bin_cnt) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const int tix = tidx[t];
    long long work = 0;
    if (tix >= 0) {
        for (int e = d.place.rowptr[tix]; e < d.place.rowptr[tix + 1]; ++e) {
            const int pl = d.place.col[e];
            if (aux.head_slot[pl] < 0) work += aux.pcp[pl + 1] - aux.pcp[pl];
        }
    }
    int parts = (int)min((long long)max_parts, 1 + work / POST_PART_WORK);
    work /= parts;
    int b = 0;
    while (b < 31 && (work >> (b + 1)) > 0) ++b;
    bin_of[t] = b;
    parts_of[t] = parts;
    atomicAdd(bin_cnt + b, parts);
}
// bin counts -> start offsets, heaviest bin first (32 bins: one thread)
__global__ void knn_post_order_kernel(int *__restrict__ bin_cnt, int *__restrict__ n_items) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int at = 0;
        for (int b = 31; b >= 0; --b) {
            int c = bin_cnt[b];
            bin_cnt[b] = at;
            at += c;
        }
        *n_items = at;
    }
}
__global__ void knn_post_scatter_kernel(int n_targets, const int *__restrict__ bin_of, const int *__restrict__ parts_of,
                                        const int *__restrict__ bin_start, int *__restrict__ bin_fill,
                                        int *__restrict__ items) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const int b = bin_of[t], parts = parts_of[t];
    const int at = bin_start[b] + atomicAdd(bin_fill + b, parts);
    for (int p_ = 0; p_ < parts; ++p_) items[at + p_] = (t << 4) | (p_ << 2) | (parts - 1);
}

// start thresholds of the dense kernel from the postings lists: a full list's smallest similarity is reached by
// K candidates, hence a lower bound of the target's K-th best
__global__ void knn_seed_from_post_kernel(int n_targets, int K, int S, int parts, int part_stride,
                                          const Nb *__restrict__ part, const int *__restrict__ part_cnt,
                                          double *__restrict__ seed_thr) {
    const int lane = threadIdx.x & 31;
    const int tt = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (tt >= n_targets) return;
    double v = seed_thr[tt];
    for (int p_ = 0; p_ < parts; ++p_) {
        if (part_cnt[tt * part_stride + S + p_] >= K) {
            const Nb *src = part + ((size_t)tt * part_stride + S + p_) * K;
            double mn = 1.0e300;
            for (int j = lane; j < K; j += 32) mn = fmin(mn, src[j].sim);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
            v = fmax(v, mn);
        }
    }
    if (lane == 0) seed_thr[tt] = v;
}

#ifndef VREC_POST_MINB
#define VREC_POST_MINB 4
#endif
__global__ void __launch_bounds__(POST_WARPS * 32, VREC_POST_MINB)
knn_postings_kernel(KnnDev d, TileAux aux, const int *__restrict__ tidx, int n_targets, int K, int S, int part_stride,
                    int cat_dim, double pw, double cw, Nb *__restrict__ part, int *__restrict__ part_cnt,
                    const double *__restrict__ seed_thr, int *__restrict__ work_counter,
                    const int *__restrict__ items, const int *__restrict__ n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *w = smem_raw + (size_t)warp * post_warp_bytes(K);
    TileSmem sm;
    double *s_pval, *s_cval, *s_dense;
    int *s_pcol, *s_ccol;
    {
        unsigned char *p = w;
        sm.hsim = (double *)p;          p += 8 * K;
        s_pval = (double *)p;           p += 8 * POST_STG_P;
        s_cval = (double *)p;           p += 8 * POST_STG_C;
        s_dense = (double *)p;          p += 8 * POST_DENSE;
        sm.hidx = (int *)p;             p += 4 * K;
        s_pcol = (int *)p;              p += 4 * POST_STG_P;
        s_ccol = (int *)p;              p += 4 * POST_STG_C;
        sm.hcnt = (int *)p;             p += 4;
        sm.lock = (int *)p;             p += 4;
        sm.tid_of = (int *)p;           p += 4;
        sm.thr = (float *)p;            p += 4;
        sm.stats = (unsigned int *)p;   p += 16;
        sm.thr_stride = 1;
        sm.tvec = nullptr;
        sm.queue = nullptr;
        sm.qn = nullptr;
    }
    if (lane < 4) sm.stats[lane] = 0;
    const bool can_stage = aux.meta != nullptr && cat_dim <= POST_DENSE;
    const long long lo = 0, hi = d.P;
    for (;;) {
        int tt = 0;
        if (lane == 0) tt = atomicAdd(work_counter, 1);
        tt = __shfl_sync(0xffffffffu, tt, 0);
        if (tt >= (items ? *n_items : n_targets)) break;
        int part_i = 0, n_parts = 1;
        if (items) {                                    // heaviest work items first
            const int it_ = items[tt];
            tt = it_ >> 4;
            part_i = (it_ >> 2) & 3;
            n_parts = (it_ & 3) + 1;
        }
        const int tix = tidx[tt];
        __syncwarp();
        if (lane == 0) {
            *sm.hcnt = 0;
            *sm.lock = 0;
            *sm.tid_of = tix;
            *sm.thr = 0.0f;
        }
        __syncwarp();
        if (tix >= 0) {
            // lower bound of the K-th best: the seed, and every full list the dense kernel produced
            double thr0 = seed_thr[tt];
            for (int sp = 0; sp < S; ++sp) {
                const int cnt = part_cnt[tt * part_stride + sp];
                if (cnt >= K) {
                    const Nb *src = part + ((size_t)tt * part_stride + sp) * K;
                    double mn = 1.0e300;
                    for (int j = lane; j < K; j += 32) mn = fmin(mn, src[j].sim);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                    thr0 = fmax(thr0, mn);
                }
            }
            TargetRows tr = load_target(d, tix);
            StagedTarget stg;
            bool staged = false;
            if (tr.pn <= POST_STG_P && tr.cn <= POST_STG_C) {
                unsigned long long sig = 0ULL;
                for (int e = lane; e < POST_DENSE; e += 32) s_dense[e] = 0.0;
                __syncwarp();
                for (int e = lane; e < tr.pn; e += 32) {
                    int col = tr.pcol[e];
                    s_pcol[e] = col;
                    s_pval[e] = tr.pval[e];
                    sig |= 1ULL << sig_bit(col);
                }
                for (int e = lane; e < tr.cn; e += 32) {
                    s_ccol[e] = tr.ccol[e];
                    s_cval[e] = tr.cval[e];
                    if (can_stage) s_dense[tr.ccol[e]] = tr.cval[e];
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sig |= __shfl_xor_sync(0xffffffffu, sig, off);
                tr.pcol = s_pcol;
                tr.pval = s_pval;
                tr.ccol = s_ccol;
                tr.cval = s_cval;
                staged = can_stage;
                stg.t = tr.t;
                stg.pn = tr.pn;
                stg.pcol = s_pcol;
                stg.pval = s_pval;
                stg.sig = sig;
                stg.cat_dense = s_dense;
                stg.plen = tr.plen;
                stg.clen = tr.clen;
            }
            __syncwarp();
            const int ps = d.place.rowptr[tix], pn = d.place.rowptr[tix + 1] - ps;
            for (int e0 = 0; e0 < pn; e0 += 32) {
                int e = e0 + lane, start = 0, len = 0, pl = -1;
                if (e < pn) {
                    pl = d.place.col[ps + e];
                    if (aux.head_slot[pl] < 0) {
                        int b = aux.pcp[pl], en = aux.pcp[pl + 1];
                        int l0 = b, h0 = en;
                        while (l0 < h0) {
                            int mid = (l0 + h0) >> 1;
                            if (aux.pper[mid] < lo) l0 = mid + 1; else h0 = mid;
                        }
                        int l1 = l0, h1 = en;
                        while (l1 < h1) {
                            int mid = (l1 + h1) >> 1;
                            if (aux.pper[mid] < hi) l1 = mid + 1; else h1 = mid;
                        }
                        start = l0;
                        len = l1 - l0;
                    }
                }
                int incl = len;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - len;
                auto locate = [&](int j, int &pl_out) -> int {
                    int L = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        int probe = __shfl_sync(0xffffffffu, incl, L + step - 1);
                        if (probe <= j) L += step;
                    }
                    L = min(L, 31);
                    int ex_l = __shfl_sync(0xffffffffu, excl, L);
                    int st_l = __shfl_sync(0xffffffffu, start, L);
                    pl_out = __shfl_sync(0xffffffffu, pl, L);
                    return j < total ? st_l + (j - ex_l) : -1;
                };
                auto fetch = [&](int j, int &cand_out, unsigned long long &meta_out, int &pl_out) {
                    int at = locate(j, pl_out);
                    cand_out = at >= 0 ? __ldg(aux.pper + at) : -1;
                    const unsigned long long *mp = aux.cmeta ? aux.cmeta : aux.meta;
                    meta_out = (cand_out >= 0 && mp) ? __ldg(mp + cand_out) : 0ULL;
                };
                auto prefetch_record = [&](int cand_, unsigned long long m_) {
                    if (cand_ >= 0 && aux.cmeta) {
                        const char *r = reinterpret_cast<const char *>(aux.crec + (m_ & 0xffffffffffULL));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
                    } else if (cand_ >= 0 && aux.meta) {
                        const char *r = reinterpret_cast<const char *>(aux.rec + (m_ & 0xffffffffffULL));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128));
                    }
                };
                // this part's blocks of 32 candidates: part_i, part_i + n_parts, ...
                const int jstep = 32 * n_parts;
                int c0 = -1, c1 = -1, p0 = -1, p1 = -1;
                unsigned long long m0 = 0, m1 = 0;
                fetch(32 * part_i + lane, c0, m0, p0);
                fetch(32 * part_i + jstep + lane, c1, m1, p1);
                prefetch_record(c0, m0);
                for (int j0 = 32 * part_i; j0 < total; j0 += jstep) {
                    int c2, p2;
                    unsigned long long m2;
                    fetch(j0 + 2 * jstep + lane, c2, m2, p2);
                    prefetch_record(c1, m1);
                    if (c0 >= 0) {
                        if (staged) tile_process_staged(aux, sm, 0, c0, K, pw, cw, p0, stg, thr0);
                        else tile_process_rows(d, aux, sm, 0, c0, K, pw, cw, p0, tr, thr0);
                    }
                    __syncwarp();
                    c0 = c1; m0 = m1; p0 = p1;
                    c1 = c2; m1 = m2; p1 = p2;
                }
            }
        }
        __syncwarp();
        const int cnt = tix >= 0 ? *(volatile int *)sm.hcnt : 0;
        Nb *out = part + ((size_t)tt * part_stride + S + part_i) * K;
        for (int j = lane; j < cnt; j += 32) {
            Nb e;
            e.sim = *(volatile double *)(sm.hsim + j);
            e.idx = *(volatile int *)(sm.hidx + j);
            e.pad = 0;
            out[j] = e;
        }
        if (lane == 0) part_cnt[tt * part_stride + S + part_i] = cnt;
        __syncwarp();
    }
    if (lane < 4) atomicAdd(&g_tile_stats[lane], (unsigned long long)sm.stats[lane]);
}


// ---------------------------------------------------------------------------------------
// Warp-specialised tensor-core kernel (the default batch path).  Roles by warp:
//   loader warp (one thread): streams B tiles with TMA bulk copies -- the fp16 features are stored in HBM as
//       ready-made 32 KB shared-memory images (128 persons, K-major, SWIZZLE_128B), so one cp.async.bulk per tile
//       lands them in place -- WS_STAGES tiles ahead;
//   MMA warp (warp-uniform loop, one elected lane issues): a tile's 8 tcgen05.mma into one of WS_NACC rotating
//       TMEM accumulators, A operand in tensor memory;
//   filter warps (WS_FILTER_WARPS, a multiple of 4: one group per TMEM lane quarter): pull 32 x 32 units of an
//       accumulator out of TMEM, apply the threshold filter, push the survivors into a ring in shared memory;
//   evaluator warps (WS_EVAL_WARPS): take 32 survivors at a time and evaluate them exactly (heaps, thresholds).
// Synchronisation is mbarrier-only on the data path: full[stage] (TMA transaction bytes), tfull[acc] /
// sempty[stage] (tcgen05.commit), tempty[acc] (one arrival per unit = 16 per tile).
// ---------------------------------------------------------------------------------------
#ifndef VREC_WS_WORKERS
#define VREC_WS_WORKERS 640     // 12 filter + 8 evaluator warps (measured: profiles/r2_knn_dense_variants.log)
#endif
constexpr int WS_WORKERS = VREC_WS_WORKERS;   // consumer threads: filter warps + evaluator warps
constexpr int WS_THREADS = WS_WORKERS + 64;           // + MMA warp + loader warp
constexpr int WS_VOTE_EVERY = 4;
#ifndef VREC_WS_QCAP
#define VREC_WS_QCAP 1024
#endif
constexpr int WS_QCAP = VREC_WS_QCAP;      // survivor ring entries (power of two)
constexpr int WS_WQ = WS_QCAP / (WS_WORKERS / 32);   // survivor queue entries per consumer warp
constexpr int WS_NACC = 3;           // TMEM accumulators (3 x 128 columns); the A operand (128 targets x 128 halves) takes 64 more
#ifndef VREC_WS_STAGES
#define VREC_WS_STAGES 4
#endif
constexpr int WS_STAGES = VREC_WS_STAGES;         // B tiles in flight (the A operand lives in tensor memory, not in shared memory)
constexpr int WS_TMEM_A = WS_NACC * TC_N;   // first TMEM column of the A operand
constexpr int WS_STAGGER = 1;        // tiles between the starting points of neighbouring CTAs (small: the CTAs share each tile through L2)
#ifndef VREC_WS_STATIC_UNITS
#define VREC_WS_STATIC_UNITS 1   // 1: a lane quarter's units go round robin over its filter warps; 0: claimed with an atomic
#endif
#ifndef VREC_WS_PUSH_BALLOT
#define VREC_WS_PUSH_BALLOT 1   // survivor ring positions from a ballot per round (0: prefix scan over the lanes' counts)
#endif
#ifndef VREC_WS_EBATCH
#define VREC_WS_EBATCH 32        // ring entries an evaluator warp takes at a time (one per lane)
#endif
constexpr int WS_EBATCH = VREC_WS_EBATCH;
#ifndef VREC_WS_POLL_NS
#define VREC_WS_POLL_NS 256      // pause of an evaluator lane between two looks at its ring slot
#endif
#ifndef VREC_WS_EVAL_WARPS
#define VREC_WS_EVAL_WARPS 8     // consumer warps that only evaluate survivors (0: every consumer warp filters AND evaluates, round 1)
#endif
constexpr int WS_EVAL_WARPS = VREC_WS_EVAL_WARPS;
constexpr int WS_FILTER_WARPS = WS_WORKERS / 32 - WS_EVAL_WARPS;     // a multiple of 4: TMEM lane quarters
static_assert(WS_FILTER_WARPS >= 4 && WS_FILTER_WARPS % 4 == 0, "filter warps must cover the four TMEM lane quarters");
constexpr unsigned long long WS_Q_FREE = 0xffffffffffffULL;          // payload of a free ring slot (48 bits)
constexpr unsigned long long WS_Q_NONE = 0xffffffffffffffffULL;      // "this lane holds no survivor"
#ifndef VREC_WS_MARGIN_MUL
#define VREC_WS_MARGIN_MUL 1.002f
#define VREC_WS_MARGIN_ADD 2e-5f
#endif
constexpr float WS_MARGIN_MUL = VREC_WS_MARGIN_MUL, WS_MARGIN_ADD = VREC_WS_MARGIN_ADD;   // filter: U * MUL + ADD >= threshold
#ifndef VREC_WS_BOOT_TILES
#define VREC_WS_BOOT_TILES 1024    // measured at P = 10^6, K = 50: 256 / 512 / 1024 / 1536 / 2048 tiles -> 6.67 / 6.49 / 6.36 / 6.40 / 6.47 ms
#endif
constexpr int WS_BOOT_TILES = VREC_WS_BOOT_TILES;   // multiple of 16; 64 groups of 2 048 candidates per target

__global__ void __launch_bounds__(WS_THREADS, 1)
knn_tc_ws_kernel(KnnDev d, TileAux aux, const __half *__restrict__ featsw, const int *__restrict__ tidx,
                 int n_targets, int K, int S, int cat_dim, double pw, double cw, Nb *__restrict__ part,
                 int *__restrict__ part_cnt, const double *__restrict__ seed_thr, int part_stride, int total_tiles,
                 int pool_ints) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[WS_STAGES], sempty[WS_STAGES], tfull[WS_NACC], tempty[WS_NACC];
    __shared__ unsigned short s_toff[TC_M];
    __shared__ unsigned short s_tnp[TC_M];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int s_stats[4];
    __shared__ int s_next_unit[4];
    __shared__ int s_q_res, s_q_take, s_q_done;     // survivor ring (WS_EVAL_WARPS > 0): reserved / claimed entries, finished filter warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = tid < WS_WORKERS;
    const int tile_m = blockIdx.x, sp = blockIdx.y;
    const int t0 = tile_m * TC_M;
    const int nt = min(TC_M, n_targets - t0);
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sB0 = base;
    TileSmem sm;
    {
        unsigned char *p = base + WS_STAGES * TC_TILE_BYTES;
        sm.hsim = (double *)p;                      p += sizeof(double) * (size_t)TC_M * K;
        sm.queue = (unsigned long long *)p;         p += sizeof(unsigned long long) * WS_QCAP;
        sm.hidx = (int *)p;                         p += sizeof(int) * (size_t)TC_M * K;
        sm.thr = (float *)p;                        p += sizeof(float) * TC_M;
        sm.thr_stride = 1;
        sm.tvec = nullptr;
        sm.hcnt = (int *)p;                         p += sizeof(int) * TC_M;
        sm.lock = (int *)p;                         p += sizeof(int) * TC_M;
        sm.tid_of = (int *)p;                       p += sizeof(int) * TC_M;
        sm.qn = (int *)p;                           p += 16;
        sm.tsig = (unsigned long long *)p;          p += sizeof(unsigned long long) * TC_M;
        sm.tpool = (const int *)p;
        sm.toff = s_toff;
        sm.t_base = t0;
        sm.stats = s_stats;
    }
    if (tid < 4) s_stats[tid] = 0;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s_ = 0; s_ < WS_STAGES; ++s_) {
            tc::mbar_init(&full[s_], 1);
            tc::mbar_init(&sempty[s_], 1);
        }
        for (int a = 0; a < WS_NACC; ++a) {
            tc::mbar_init(&tfull[a], 1);
            tc::mbar_init(&tempty[a], 16);              // one arrival per unit: 4 lane quarters x 4 column quarters
        }
        tc::mbar_init_fence();
        *sm.qn = 0;
    }
    // this CTA's share of the candidate tiles, visited in a rotated order
    const int tlo = (int)((long long)total_tiles * sp / S), thi = (int)((long long)total_tiles * (sp + 1) / S);
    const int ntiles = thi - tlo;
    const int rot = ntiles > 0 ? (int)(((long long)blockIdx.x * WS_STAGGER + (long long)blockIdx.y * 29) % ntiles) : 0;
    // Bootstrap: the first `boot` tiles of the sequence are only scanned for group maxima (below) and
    // are visited again, normally, at the end of the sequence.
    const int boot = ntiles >= 64 ? min(WS_BOOT_TILES, (ntiles / 4) / 16 * 16) : 0;   // a quarter of the tiles at most
    const int nseq = ntiles + boot;
    if (tid < 4) s_next_unit[tid] = boot * 4;
    auto tile_index = [&](int i) {
        int w_ = (i >= ntiles ? i - ntiles : i) + rot;
        if (w_ >= ntiles) w_ -= ntiles;
        return tlo + w_;
    };
    if (worker) {
        for (int t = tid; t < TC_M; t += WS_WORKERS) {
            int tix = t < nt ? tidx[t0 + t] : -1;
            sm.tid_of[t] = tix;
            sm.hcnt[t] = 0;
            sm.lock[t] = 0;
            sm.thr[t] = tix >= 0 ? __double2float_rd(seed_thr[t0 + t]) : 3.0e38f;
        }
    }
    __syncthreads();
    __syncthreads();                              // tmem_base_s
    if (tid < TC_M) {
        // A operand in tensor memory: thread = target row = TMEM lane; 32-bit column j holds the halves
        // (2j, 2j+1) of the row -- the target's features (read from their tile image) with the weights folded in
        const int tix = sm.tid_of[tid];
        const uint32_t ta = tmem_base_s + ((uint32_t)(warp * 32) << 16) + WS_TMEM_A;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {                       // 16-byte chunk c = h * 8 + c4: 8 halves
                const int c = h * 8 + c4;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (tix >= 0) {
                    const unsigned char *img = reinterpret_cast<const unsigned char *>(featsw) + (size_t)(tix >> 7) * TC_TILE_BYTES;
                    raw = *reinterpret_cast<const uint4 *>(img + tc::sw128_offset(TC_N, tix & 127, c));
                }
                const __half *src = reinterpret_cast<const __half *>(&raw);
                __align__(16) __half hh[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int dd = c * 8 + e;
                    hh[e] = __float2half(__half2float(src[e]) * (float)(dd < cat_dim ? cw : pw));
                }
                const uint32_t *pk = reinterpret_cast<const uint32_t *>(hh);
#pragma unroll
                for (int e = 0; e < 4; ++e) r[c4 * 4 + e] = pk[e];
            }
            tc::tmem_st32(ta + h * 32, r);
        }
        tc::tmem_wait_st();
    }
    if (worker) {
        // the targets' place columns, packed into what is left of shared memory (first come, first served)
        int *pool = const_cast<int *>(sm.tpool);
        if (tid < TC_M) {
            const int tix = sm.tid_of[tid];
            s_tnp[tid] = tix >= 0 ? (unsigned short)((__ldg(aux.meta + tix) >> 40) & 0xfffu) : 0;
        }
        tc::bar_sync(1, WS_WORKERS);
        if (tid == 0) {
            int at = 0;
            for (int t = 0; t < TC_M; ++t) {
                const int n = s_tnp[t];
                if (at + n <= pool_ints && at + n < 0xffff) {
                    s_toff[t] = (unsigned short)at;
                    at += n;
                } else {
                    s_toff[t] = 0xffffu;
                }
            }
        }
        tc::bar_sync(1, WS_WORKERS);
        // per target: signature of its places (shared memory) and its category vector as a dense row
        for (int t = warp; t < TC_M; t += WS_WORKERS / 32) {
            const int tix = sm.tid_of[t];
            unsigned long long sig = 0ULL;
            if (tix >= 0) {
                const bool F32 = aux.vals_f32;
                const unsigned long long mt = __ldg(aux.meta + tix);
                const int npt = (int)((mt >> 40) & 0xfffu), nct = (int)(mt >> 52);
                const double *rt = aux.rec + (mt & 0xffffffffffULL);
                const int *tpc = reinterpret_cast<const int *>(rt + 2);
                const double *tpv = rt + 2 + rec_cols_words(npt);
                const int *tcc = reinterpret_cast<const int *>(tpv + rec_vals_words(npt, F32));
                const double *tcv = tpv + rec_vals_words(npt, F32) + rec_cols_words(nct);
                for (int k = lane; k < npt; k += 32) {
                    const int col = (int)((unsigned)tpc[k] & REC_COL_MASK);
                    sig |= 1ULL << sig_bit(col);
                    if (s_toff[t] != 0xffffu) pool[s_toff[t] + k] = col;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sig |= __shfl_xor_sync(0xffffffffu, sig, off);
                // The S CTAs of a tile all fill the same rows, and not at the same time (a grid of more than
                // one wave starts some of them while others already evaluate): every element is stored ONCE,
                // with its final value, so a row never passes through a state a concurrent reader must not see.
                double *row = aux.tdense + (size_t)(t0 + t) * aux.cat_dim;
                for (int c = lane; c < aux.cat_dim; c += 32) {
                    double v = 0.0;
                    for (int k = 0; k < nct; ++k)
                        if (tcc[k] == c) v = rec_val(tcv, k, F32);
                    row[c] = v;
                }
            }
            if (lane == 0) sm.tsig[t] = sig;
        }
        __threadfence_block();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;

    if (!worker) {
        const uint32_t idesc = tc::make_idesc_f16(TC_M, TC_N);
        const uint32_t b_addr0 = tc::smem_u32(sB0);
        if (warp == WS_WORKERS / 32 + 1) {
            // ================= loader warp: TMA bulk copies, up to TC_STAGES tiles ahead =================
            if (lane == 0) {
                for (int li = 0; li < nseq; ++li) {
                    const int st = li % WS_STAGES;
                    if (li >= WS_STAGES) tc::mbar_wait(&sempty[st], (uint32_t)((li / WS_STAGES - 1) & 1));   // MMA(li-4) done
                    tc::mbar_expect_tx(&full[st], TC_TILE_BYTES);
                    tc::bulk_copy_g2s(sB0 + (size_t)st * TC_TILE_BYTES,
                                      reinterpret_cast<const unsigned char *>(featsw) + (size_t)tile_index(li) * TC_TILE_BYTES,
                                      TC_TILE_BYTES, &full[st]);
                }
            }
        } else {
            // ================= MMA warp: the whole warp walks the tiles, one elected lane issues =================
            // (warp-uniform control flow and operands: the descriptors live in uniform registers; with the loop
            // under `if (lane == 0)` every tcgen05.mma was wrapped in an ELECT / R2UR loop and the issue of a
            // tile's 8 MMAs took ~850 cycles of dependent instructions -- more than their 542 cycles of math)
            const bool pprof = lane == 0 && blockIdx.x == 0 && blockIdx.y == 0;
            long long pk0 = clock64(), pk1;
            unsigned long long pacc0 = 0, pacc7 = 0, pacc8 = 0;      // phase cycles, in registers until the end
#define WS_PTICK(slot)                                                 \
    if (pprof) {                                                       \
        pk1 = clock64();                                               \
        pacc##slot += (unsigned long long)(pk1 - pk0);                 \
        pk0 = pk1;                                                     \
    }
            const uint32_t tbase_u = __shfl_sync(0xffffffffu, tbase, 0);
            const uint32_t b_addr_u = __shfl_sync(0xffffffffu, b_addr0, 0);
            for (int i = 0; i < nseq; ++i) {
                const int st = i % WS_STAGES, a = i % WS_NACC;
                tc::mbar_wait(&full[st], (uint32_t)((i / WS_STAGES) & 1));                 // B(i) landed
                WS_PTICK(0)
                if (i >= WS_NACC) tc::mbar_wait(&tempty[a], (uint32_t)((i / WS_NACC - 1) & 1));   // accumulator drained
                WS_PTICK(7)
                tc::fence_after_sync();
                const uint32_t b_addr = b_addr_u + (uint32_t)st * TC_TILE_BYTES;
                const uint32_t acc = tbase_u + (uint32_t)a * TC_N;
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        uint64_t db = tc::make_desc_sw128(tc::sw128_kstep_addr(b_addr, TC_N, k));
                        tc::mma_f16_ts(acc, tbase_u + WS_TMEM_A + (uint32_t)(k * 8), db, idesc, k > 0);
                    }
                    tc::mma_commit(&tfull[a]);        // accumulator ready for the consumers
                    tc::mma_commit(&sempty[st]);      // B stage reusable
                }
                __syncwarp();
                WS_PTICK(8)
            }
            if (pprof) {
                g_tc_cycles[0] += pacc0;
                g_tc_cycles[7] += pacc7;
                g_tc_cycles[8] += pacc8;
                g_tc_cycles[5] += (unsigned long long)nseq;
            }
#undef WS_PTICK
        }
        __syncwarp();                             // the idle lanes of the producer warps wait for lane 0 here
    } else {
        // ================= consumers: TMEM epilogue + exact survivors =================
        const int lq = warp & 3, cq = warp >> 2;
        const int my_t = lq * 32 + lane;
        const bool cprof = tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
        long long ck0 = clock64(), ck1;
        unsigned long long cacc1 = 0, cacc2 = 0, cacc10 = 0;
#define WS_CTICK(slot)                                                 \
    if (cprof) {                                                       \
        ck1 = clock64();                                               \
        cacc##slot += (unsigned long long)(ck1 - ck0);                 \
        ck0 = ck1;                                                     \
    }
        // ---- bootstrap: 16 running maxima of U per thread (group = tile % 16 within this thread's 32
        // columns).  A target row has 4 threads -> 64 disjoint groups; each group maximum is attained by a
        // distinct candidate, at most one of them the target itself, so >= 63 >= K real candidates have
        // U >= theta = min of the 64 maxima, and their exact similarity is >= theta / 1.002 - 2e-5.
        if (boot > 0) {
            float gmax[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) gmax[g] = 0.0f;
            for (int i = 0; i < (warp < 16 ? boot : 0); ++i) {     // 16 warps = the 16 units of a tile
                const int a = i % WS_NACC;
                tc::mbar_wait(&tfull[a], (uint32_t)((i / WS_NACC) & 1));
                tc::fence_after_sync();
                float v[32];
                tc::tmem_ld32(tbase + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * TC_N + cq * 32), v);
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[a]);
                float m = v[0];
#pragma unroll
                for (int jj = 1; jj < 32; ++jj) m = fmaxf(m, v[jj]);
#pragma unroll
                for (int g = 0; g < 16; ++g)
                    if (g == (i & 15)) gmax[g] = fmaxf(gmax[g], m);
            }
            // the minimum of this thread's 16 maxima, then of the row's 4 threads.  Scratch = the survivor
            // queues (8 KB, not in use before the main sequence; 2 KB needed).  It used to be 64 floats per
            // row in the heap area, which holds 8 * K bytes per row: K < 32 overran it into the queues, the
            // heap indices and the target tables (illegal address at K = 7, P = 10^6).
            static_assert(sizeof(unsigned long long) * WS_QCAP >= sizeof(float) * 4 * TC_M, "bootstrap scratch");
            float gmin = gmax[0];
#pragma unroll
            for (int g = 1; g < 16; ++g) gmin = fminf(gmin, gmax[g]);
            float *scratch = reinterpret_cast<float *>(sm.queue);
            if (warp < 16) scratch[my_t * 4 + cq] = gmin;
            tc::bar_sync(1, WS_WORKERS);
            if (tid < TC_M && sm.tid_of[tid] >= 0) {
                float theta = scratch[tid * 4];
                for (int g = 1; g < 4; ++g) theta = fminf(theta, scratch[tid * 4 + g]);
                float thr0 = theta / 1.002f - 2e-5f;
                thr0 = nextafterf(thr0, 0.0f);                             // keep the bound on the safe side
                if (thr0 > sm.thr[tid]) sm.thr[tid] = thr0;
            }
            tc::bar_sync(1, WS_WORKERS);
        }
        // ---- main sequence.  The unit of consumer work is (tile, lane quarter, column quarter) = 32 target
        // rows x 32 candidates; the warps of a lane quarter claim units in order from a shared counter.
#if VREC_WS_EVAL_WARPS > 0
        // Round 2: the consumer warps are split.  FILTER warps (WS_FILTER_WARPS / 4 per TMEM lane quarter) only drain accumulators,
        // test thresholds and push survivors into a ring in shared memory; EVALUATOR warps only pull survivors
        // (32 at a time, one per lane) and evaluate them exactly.  An exact evaluation is a chain of dependent
        // memory reads (a drain of 32 took ~25 K cycles = 13 tile periods): with every warp doing both, a lane
        // quarter whose warps happened to be evaluating held up its accumulator, and the MMA thread waited 737
        // cycles per tile for one (measured, tools/knn_bench.py) -- now no warp that touches TMEM ever blocks.
        // Ring slot = lap (16 bits) | payload (48): FREE(L) -> entry of lap L (target << 32 | candidate) -> FREE(L+1).
        // A writer of lap L waits for FREE(L), a reader of lap L for an entry tagged L: writers of different laps
        // can never race for a slot (with a plain "empty" marker a lap-L+1 entry could overtake the lap-L one,
        // and with every evaluator waiting for a lap-L entry nobody would claim it -- a deadlock the reduced-size
        // tests, whose loose thresholds flood the ring, ran into at once).
        for (int q = tid; q < WS_QCAP; q += WS_WORKERS) sm.queue[q] = WS_Q_FREE;                 // FREE(0)
        if (tid == 0) {
            s_q_res = 0;
            s_q_take = 0;
            s_q_done = 0;
        }
        tc::bar_sync(1, WS_WORKERS);
        volatile unsigned long long *ring = sm.queue;
        if (warp >= WS_FILTER_WARPS) {
            // ================= evaluator warps =================
            const bool eprof = warp == WS_FILTER_WARPS && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0;
            unsigned long long e_wait = 0, e_eval = 0, e_n = 0;     // first evaluator warp of block 0: cycles waiting / evaluating, batches
            for (;;) {
                const long long ek0 = clock64();
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_q_take, WS_EBATCH);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const int my = pos + lane;
                const unsigned long long lap = (unsigned long long)((my / WS_QCAP) & 0xffff);
                unsigned long long e = WS_Q_NONE;
                for (; lane < WS_EBATCH;) {
                    const unsigned long long v = ring[my & (WS_QCAP - 1)];
                    if ((v >> 48) == lap && (v & WS_Q_FREE) != WS_Q_FREE) {
                        e = v;
                        ring[my & (WS_QCAP - 1)] = (((lap + 1) & 0xffff) << 48) | WS_Q_FREE;     // FREE(lap + 1)
                        break;
                    }
                    if (*(volatile int *)&s_q_done == WS_FILTER_WARPS && my >= *(volatile int *)&s_q_res) break;
                    __nanosleep(VREC_WS_POLL_NS);
                }
                __syncwarp();
                if (!__any_sync(0xffffffffu, e != WS_Q_NONE)) break;           // nothing left for this warp
                const int tt = e != WS_Q_NONE ? (int)((e >> 32) & 0xffu) : -1;
                const long long ek1 = clock64();
                tile_process_sig(aux, sm, tt, (int)(unsigned)(e & 0xffffffffu), K, pw, cw);
                if (eprof) {
                    e_wait += (unsigned long long)(ek1 - ek0);
                    e_eval += (unsigned long long)(clock64() - ek1);
                    e_n += 1;
                }
            }
            if (eprof) {
                g_tc_cycles[3] += e_wait;
                g_tc_cycles[4] += e_eval;
                g_tc_cycles[6] += e_n;
            }
        } else {
            // ================= filter warps =================
            const int n_units = nseq * 4;
            const int ulq = lq;
#if VREC_WS_STATIC_UNITS
            // a lane quarter's units go round robin over its filter warps (they never block, so a dynamic claim --
            // an atomic on shared memory plus a shuffle per unit -- buys nothing)
            for (int u = boot * 4 + (warp >> 2); u < n_units; u += WS_FILTER_WARPS / 4) {
#else
            for (;;) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&s_next_unit[ulq], 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= n_units) break;
#endif
                const int i = u >> 2, ucq = u & 3;
                const int a = i % WS_NACC;
                tc::mbar_wait(&tfull[a], (uint32_t)((i / WS_NACC) & 1));
                tc::fence_after_sync();
                WS_CTICK(1)                                       // claim + wait for the accumulator
                const int ut = ulq * 32 + lane;
                const float thr = *(volatile float *)(sm.thr + ut);
                const long long tile = (long long)tile_index(i) * TC_N;
                const int c0 = ucq * 32;
                float v[32];
                tc::tmem_ld32(tbase + ((uint32_t)(ulq * 32) << 16) + (uint32_t)(a * TC_N + c0), v);
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[a]);            // this unit is out of the accumulator
                float m8[8];                                           // maximum by a tree: depth 5 instead of 31
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) m8[jj] = fmaxf(fmaxf(v[jj], v[jj + 8]), fmaxf(v[jj + 16], v[jj + 24]));
                const float vmax = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                         fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                unsigned pass = 0;
                if (VREC_WS_ABLATE != 2 && vmax * WS_MARGIN_MUL + WS_MARGIN_ADD >= thr) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) pass |= (v[jj] * WS_MARGIN_MUL + WS_MARGIN_ADD >= thr ? 1u : 0u) << jj;
                }
                if (tile + c0 + 32 > d.P) {                            // last tile: columns past the last person
                    const long long over = tile + c0 + 32 - d.P;
                    pass = over >= 32 ? 0u : (pass & (0xffffffffu >> over));
                }
#if VREC_WS_ABLATE == 1 || VREC_WS_ABLATE == 2
                pass = 0;                                          // timing experiment only: drop every survivor
#endif
                const unsigned any = __ballot_sync(0xffffffffu, pass != 0u);
                WS_CTICK(2)                                       // epilogue
                if (any) {
#if VREC_WS_PUSH_BALLOT
                    // one round per survivor of the busiest lane (nearly always one): positions from a ballot
                    unsigned has = any;
                    while (has) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s_q_res, __popc(has));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (pass) {
                            const int pos = base + __popc(has & ((1u << lane) - 1u));
                            const int jj = __ffs(pass) - 1;
                            pass &= pass - 1;
                            const unsigned long long tag = (unsigned long long)((pos / WS_QCAP) & 0xffff) << 48;
                            while (ring[pos & (WS_QCAP - 1)] != (tag | WS_Q_FREE)) __nanosleep(128);   // previous lap not taken yet
                            ring[pos & (WS_QCAP - 1)] = tag | ((unsigned long long)ut << 32) |
                                                        (unsigned long long)(unsigned)(tile + c0 + jj);
                            if (aux.meta) asm volatile("prefetch.global.L2 [%0];" ::"l"((aux.cmeta ? aux.cmeta : aux.meta) + tile + c0 + jj));
                        }
                        has = __ballot_sync(0xffffffffu, pass != 0u);
                    }
#else
                    int cnt = __popc(pass), incl = cnt;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        int o = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += o;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_q_res, total);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    int pos = base + incl - cnt;
                    while (pass) {
                        const int jj = __ffs(pass) - 1;
                        pass &= pass - 1;
                        const unsigned long long tag = (unsigned long long)((pos / WS_QCAP) & 0xffff) << 48;
                        while (ring[pos & (WS_QCAP - 1)] != (tag | WS_Q_FREE)) __nanosleep(128);   // previous lap not taken yet
                        ring[pos & (WS_QCAP - 1)] = tag | ((unsigned long long)ut << 32) |
                                                    (unsigned long long)(unsigned)(tile + c0 + jj);
                        ++pos;
                        if (aux.meta) asm volatile("prefetch.global.L2 [%0];" ::"l"((aux.cmeta ? aux.cmeta : aux.meta) + tile + c0 + jj));
                    }
#endif
                    __syncwarp();
                    WS_CTICK(10)                                  // queueing
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&s_q_done, 1);
        }
#else
        // (round 1) the warps claim units in order from shared counters, so a warp that is busy
        // with exact evaluations (a chain of dependent HBM reads, tens of thousands of cycles) never holds up
        // an accumulator: the others take over its units and the tensor pipe keeps running.  Survivors go
        // to a per-warp queue that the warp drains by itself once it holds one candidate per lane.
        unsigned long long *wq = sm.queue + (size_t)warp * WS_WQ;
        int wn = 0;                                              // warp-uniform
        auto drain_own = [&]() {
            for (int q0 = 0; q0 < wn; q0 += 32) {
                const int qi = q0 + lane;
                const unsigned long long e = qi < wn ? wq[qi] : 0xffffffff00000000ULL;      // target -1: idle lane
                tile_process_sig(aux, sm, (int)(e >> 32), (int)(unsigned)(e & 0xffffffffu), K, pw, cw);
            }
            __syncwarp();
            wn = 0;
        };
        // (a warp can only read the TMEM lanes 32 * (warp % 4) ..: the four warps of a lane quarter share
        // that quarter's units)
        const int n_units = nseq * 4;
        const int ulq = lq;
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_next_unit[ulq], 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= n_units) break;
            const int i = u >> 2, ucq = u & 3;
            const int a = i % WS_NACC;
            tc::mbar_wait(&tfull[a], (uint32_t)((i / WS_NACC) & 1));
            tc::fence_after_sync();
            WS_CTICK(1)                                       // claim + wait for the accumulator
            const int ut = ulq * 32 + lane;
            const float thr = *(volatile float *)(sm.thr + ut);
            const long long tile = (long long)tile_index(i) * TC_N;
            const int c0 = ucq * 32;
            float v[32];
            tc::tmem_ld32(tbase + ((uint32_t)(ulq * 32) << 16) + (uint32_t)(a * TC_N + c0), v);
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[a]);            // this unit is out of the accumulator
            // most rows have no survivor among their 32 values: one running maximum first, the per-value
            // test only for the rows whose maximum passes a (slightly lower, hence safe) bound
            float vmax = v[0];
#pragma unroll
            for (int jj = 1; jj < 32; ++jj) vmax = fmaxf(vmax, v[jj]);
            unsigned pass = 0;
            if (vmax * 1.002f + 2e-5f >= thr) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) pass |= (v[jj] * 1.002f + 2e-5f >= thr ? 1u : 0u) << jj;
            }
            if (tile + c0 + 32 > d.P) {                            // last tile: columns past the last person
                const long long over = tile + c0 + 32 - d.P;
                pass = over >= 32 ? 0u : (pass & (0xffffffffu >> over));
            }
            const unsigned any = __ballot_sync(0xffffffffu, pass != 0u);
            WS_CTICK(2)                                       // epilogue
            if (any) {
                int cnt = __popc(pass), incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    int o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (wn + total > WS_WQ) drain_own();
                if (total > WS_WQ) {
                    // more survivors in one unit than the queue holds (thresholds still loose): inline
                    while (__any_sync(0xffffffffu, pass != 0u)) {
                        const int jj = pass ? __ffs(pass) - 1 : 0;
                        const int tt = pass ? ut : -1;
                        pass &= pass - 1;
                        tile_process_sig(aux, sm, tt, (int)(tile + c0 + jj), K, pw, cw);
                    }
                    __syncwarp();
                } else {
                    int pos = wn + incl - cnt;
                    while (pass) {
                        const int jj = __ffs(pass) - 1;
                        pass &= pass - 1;
                        wq[pos++] = ((unsigned long long)ut << 32) | (unsigned long long)(unsigned)(tile + c0 + jj);
                        if (aux.meta) asm volatile("prefetch.global.L2 [%0];" ::"l"((aux.cmeta ? aux.cmeta : aux.meta) + tile + c0 + jj));
                    }
                    __syncwarp();
                    wn += total;
                    if (wn >= 32) drain_own();
                }
                WS_CTICK(10)                                  // queueing + own drains
            }
        }
        drain_own();
#endif
        if (cprof) {
            g_tc_cycles[1] += cacc1;
            g_tc_cycles[2] += cacc2;
            g_tc_cycles[10] += cacc10;
        }
#undef WS_CTICK
    }
    __syncthreads();
    if (tid < 4) atomicAdd(&g_tile_stats[tid], (unsigned long long)s_stats[tid]);
    if (worker) {
        for (int t = 0; t < nt; ++t) {
            int cnt = sm.hcnt[t];
            Nb *out = part + ((size_t)(t0 + t) * part_stride + sp) * K;
            for (int j = tid; j < cnt; j += WS_WORKERS) {
                Nb e;
                e.sim = sm.hsim[(size_t)t * K + j];
                e.idx = sm.hidx[(size_t)t * K + j];
                e.pad = 0;
                out[j] = e;
            }
            if (tid == 0) part_cnt[(t0 + t) * part_stride + sp] = cnt;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// fp16 features as ready-made shared-memory tile images: tile b = persons [128 b, 128 b + 128),
// 32 KB each in the SWIZZLE_128B K-major layout of vrec_tc.cuh (rows past P stay zero)
__global__ void knn_features16sw_kernel(KnnDev d, const short *__restrict__ head_slot, int cat_dim,
                                        unsigned char *__restrict__ featsw) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.P) return;
    __half row[TC_D];
    for (int dd = 0; dd < TC_D; ++dd) row[dd] = __float2half(0.0f);
    double cl = d.cat.len[i], pl = d.place.len[i];
    for (int k = d.cat.rowptr[i]; k < d.cat.rowptr[i + 1]; ++k) row[d.cat.col[k]] = __double2half(d.cat.val[k] / cl);
    for (int k = d.place.rowptr[i]; k < d.place.rowptr[i + 1]; ++k) {
        int slot = head_slot[d.place.col[k]];
        if (slot >= 0) row[cat_dim + slot] = __double2half(d.place.val[k] / pl);
    }
    unsigned char *img = featsw + (size_t)(i >> 7) * TC_TILE_BYTES;
    for (int c = 0; c < TC_D / 8; ++c)
        *reinterpret_cast<uint4 *>(img + tc::sw128_offset(TC_N, (int)(i & 127), c)) = *reinterpret_cast<const uint4 *>(row + c * 8);
}

#if VREC_WITH_TC_BASELINE
// fp16 features, row-major [P][TC_D]: category vector / length, then the head places' values / length
__global__ void knn_features16_kernel(KnnDev d, const short *__restrict__ head_slot, int cat_dim,
                                      __half *__restrict__ feat16) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.P) return;
    __half *row = feat16 + (size_t)i * TC_D;
    for (int dd = 0; dd < TC_D; ++dd) row[dd] = __float2half(0.0f);
    double cl = d.cat.len[i], pl = d.place.len[i];
    for (int k = d.cat.rowptr[i]; k < d.cat.rowptr[i + 1]; ++k) row[d.cat.col[k]] = __double2half(d.cat.val[k] / cl);
    for (int k = d.place.rowptr[i]; k < d.place.rowptr[i + 1]; ++k) {
        int slot = head_slot[d.place.col[k]];
        if (slot >= 0) row[cat_dim + slot] = __double2half(d.place.val[k] / pl);
    }
}
#endif

// Dense fp32 features, dim-major: F[d][i] = cat[i][d] / |cat_i| for d < cat_dim, then the head
// places' values / |place_i|.  `wscale` (host) is applied per target in knn_tile_kernel.
__global__ void knn_features_kernel(KnnDev d, const short *__restrict__ head_slot, int cat_dim, long long fstride,
                                    float *__restrict__ feat) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.P) return;
    for (int dd = 0; dd < TILE_D; ++dd) feat[(size_t)dd * fstride + i] = 0.0f;
    double cl = d.cat.len[i], pl = d.place.len[i];
    for (int k = d.cat.rowptr[i]; k < d.cat.rowptr[i + 1]; ++k)
        feat[(size_t)d.cat.col[k] * fstride + i] = (float)(d.cat.val[k] / cl);
    for (int k = d.place.rowptr[i]; k < d.place.rowptr[i + 1]; ++k) {
        int slot = head_slot[d.place.col[k]];
        if (slot >= 0) feat[(size_t)(cat_dim + slot) * fstride + i] = (float)(d.place.val[k] / pl);
    }
}

// Bitonic sort of 32 * NQ (key, payload) pairs held by one warp, element i = q * 32 + lane, "best first" in the
// (similarity desc, index asc) order: partner distances below 32 are shuffles, the others stay in the thread.
template <int NQ>
__device__ __forceinline__ void warp_bitonic_best_first(double (&ks)[NQ], int (&ki)[NQ], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * NQ; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dq = j >> 5;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if ((q & dq) == 0) {
                        const int i = q * 32 + lane;
                        const bool up = (i & k) == 0;                       // best first inside an "up" block
                        const bool second_before = ks[q ^ dq] > ks[q] || (ks[q ^ dq] == ks[q] && ki[q ^ dq] < ki[q]);
                        const bool first_before = ks[q] > ks[q ^ dq] || (ks[q] == ks[q ^ dq] && ki[q] < ki[q ^ dq]);
                        if (up ? second_before : first_before) {
                            const double ts = ks[q];
                            const int ti = ki[q];
                            ks[q] = ks[q ^ dq];
                            ki[q] = ki[q ^ dq];
                            ks[q ^ dq] = ts;
                            ki[q ^ dq] = ti;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int i = q * 32 + lane;
                    const double os = __shfl_xor_sync(0xffffffffu, ks[q], j);
                    const int oi = __shfl_xor_sync(0xffffffffu, ki[q], j);
                    const bool up = (i & k) == 0;
                    const bool lower = (lane & j) == 0;                       // this lane holds the smaller position
                    const bool other_before = os > ks[q] || (os == ks[q] && oi < ki[q]);
                    const bool mine_before = ks[q] > os || (ks[q] == os && ki[q] < oi);
                    // the smaller position of an "up" block keeps the better element, and so on
                    const bool take = (up == lower) ? other_before : mine_before;
                    if (take) {
                        ks[q] = os;
                        ki[q] = oi;
                    }
                }
            }
        }
    }
}

// knn_merge_kernel for S * K <= 256 (the batch path: 5 lists of 50): one warp per target, everything in
// registers, no block barriers.  Same outputs.
constexpr int MERGE_WARP_MAX = 256;
__global__ void __launch_bounds__(128)
knn_merge_warp_kernel(const Nb *__restrict__ part, const int *__restrict__ part_cnt, int n_targets, int K, int S,
                      Nb *__restrict__ nb_rank, Nb *__restrict__ nb_idx, int *__restrict__ nb_cnt) {
    const int lane = threadIdx.x & 31;
    const int tt = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (tt >= n_targets) return;
    constexpr int NQ = MERGE_WARP_MAX / 32;
    double ks[NQ];
    int ki[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        ks[q] = -1.0;                                   // sentinels come last
        ki[q] = 0x7fffffff;
    }
    // concatenate the lists: element position p = running offset + index within the list
    int total = 0;
    for (int s_ = 0; s_ < S; ++s_) {
        const int c = part_cnt[tt * S + s_];
        const Nb *src = part + ((size_t)tt * S + s_) * K;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int p_ = q * 32 + lane - total;
            if (p_ >= 0 && p_ < c) {
                const Nb e = src[p_];
                ks[q] = e.sim;
                ki[q] = e.idx;
            }
        }
        total += c;
    }
    warp_bitonic_best_first<NQ>(ks, ki, lane);
    const int keep = min(total, K);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = q * 32 + lane;
        if (i < keep) {
            Nb e;
            e.sim = ks[q];
            e.idx = ki[q];
            e.pad = 0;
            nb_rank[(size_t)tt * K + i] = e;
        }
    }
    if (lane == 0) nb_cnt[tt] = keep;
    __syncwarp();                                       // nb_rank is read back below by other lanes
    // the kept entries again, in ascending person index: sort (key = -index, payload = rank position)
    constexpr int NQ2 = 2;                              // keep <= K <= 64 on this path
    double k2[NQ2];
    int p2[NQ2];
#pragma unroll
    for (int q = 0; q < NQ2; ++q) {
        const int i = q * 32 + lane;
        k2[q] = i < keep ? -(double)ki[q] : -4.0e9;      // ascending index == descending -index
        p2[q] = i;
    }
    warp_bitonic_best_first<NQ2>(k2, p2, lane);
#pragma unroll
    for (int q = 0; q < NQ2; ++q) {
        const int i = q * 32 + lane;
        if (i < keep) {
            // p2[q] = rank position of the entry with the i-th smallest index; fetch its similarity from its holder
            const int rp = p2[q];
            Nb e = nb_rank[(size_t)tt * K + rp];         // written above by this warp
            e.pad = 0;
            nb_idx[(size_t)tt * K + i] = e;
        }
    }
}

// Merges the S partial lists of a target (S*K <= TOPK_BUF): final neighbours in
// (similarity desc, index asc) order -> nb_rank, and the same set in ascending index -> nb_idx.
__global__ void __launch_bounds__(MERGE_THREADS)
knn_merge_kernel(const Nb *__restrict__ part, const int *__restrict__ part_cnt, int K, int S,
                 Nb *__restrict__ nb_rank, Nb *__restrict__ nb_idx, int *__restrict__ nb_cnt) {
    const int tt = blockIdx.x, tid = threadIdx.x;
    __shared__ Nb buf[TOPK_BUF];
    __shared__ int s_off[34];
    if (tid == 0) {
        int o = 0;
        for (int s = 0; s < S; ++s) {
            s_off[s] = o;
            o += part_cnt[tt * S + s];
        }
        s_off[S] = o;
    }
    __syncthreads();
    const int total = s_off[S];
    for (int s = 0; s < S; ++s) {
        int c = s_off[s + 1] - s_off[s];
        const Nb *src = part + ((size_t)tt * S + s) * K;
        for (int i = tid; i < c; i += MERGE_THREADS) buf[s_off[s] + i] = src[i];
    }
    int np2 = next_pow2(max(total, 1));
    for (int i = total + tid; i < np2; i += MERGE_THREADS) {
        buf[i].sim = -1.0;
        buf[i].idx = 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_desc(buf, np2, tid, MERGE_THREADS);
    const int keep = min(total, K);
    for (int i = tid; i < keep; i += MERGE_THREADS) nb_rank[(size_t)tt * K + i] = buf[i];
    if (tid == 0) nb_cnt[tt] = keep;
    __syncthreads();
    // re-sort the kept entries by ascending index: flip the key (sim := -idx) and reuse the sorter
    int kp2 = next_pow2(max(keep, 1));
    Nb mine[TOPK_BUF / MERGE_THREADS];
    for (int i = tid, j = 0; i < kp2; i += MERGE_THREADS, ++j) {
        Nb e = buf[i];
        if (i >= keep) {
            e.sim = -1.0;
            e.idx = 0x7fffffff;
        }
        mine[j] = e;
    }
    __syncthreads();
    // second array lives in registers while we sort (idx, original sim) pairs by idx:
    // encode as sim' = -(double)idx (exact for idx < 2^31), keep the true sim in a side array
    __shared__ double s_true[TOPK_MAX_K];
    for (int i = tid, j = 0; i < kp2; i += MERGE_THREADS, ++j) {
        Nb e = mine[j];
        if (i < keep) s_true[i] = e.sim;
        Nb f;
        f.idx = i;                                       // remember the rank position
        f.pad = e.idx;                                   // person index
        f.sim = (i < keep) ? -(double)e.idx : -4.0e9;    // ascending idx == descending -idx
        buf[i] = f;
    }
    __syncthreads();
    bitonic_sort_desc(buf, kp2, tid, MERGE_THREADS);
    for (int i = tid; i < keep; i += MERGE_THREADS) {
        Nb f = buf[i];
        Nb e;
        e.sim = s_true[f.idx];
        e.idx = f.pad;
        e.pad = 0;
        nb_idx[(size_t)tt * K + i] = e;
    }
}

// makeRecommendations0 (knn/KnnRecommender.scala:51-70) by neighbour-row gather, one warp per
// target: neighbours in ascending person index, num/den accumulated in the same pass in a
// dense per-warp scratch row; then the region filter + ranked top-N
// (knn/KnnRecommenderMain.scala:96-100).
// The scratch rows (2 x 8 bytes per place and warp: 3.8 GB for 2 368 warps x 100 000 places) are far larger than
// L2, so every num/den update of a dense row is a DRAM round trip (2.8 GB read per 18 944 targets, 1.1 ms).  A
// target whose neighbours hold at most RATE_HASH_MAX ratings (K = 50: ~400) instead maps place -> slot through a
// RATE_HASH-entry open-addressing table in shared memory and keeps num/den in the first RATE_HASH entries of its
// rows (38 MB over all warps: L2-resident).  Same additions in the same order; the rows are zero again afterwards.
constexpr int RATE_HASH = 1024;          // table entries per warp (power of two)
constexpr int RATE_HASH_MAX = 512;       // ratings of a target's neighbours up to which the table is used (= 32 lanes x RG below)
__global__ void __launch_bounds__(RATE_WARPS * 32)
knn_rate_gather_kernel(const int *__restrict__ rrp, const int *__restrict__ rpl, const double *__restrict__ rv,
                       const Nb *__restrict__ nb_idx, const int *__restrict__ nb_cnt, int T, int K, int rdim,
                       const unsigned char *__restrict__ flag, double *__restrict__ num, double *__restrict__ den,
                       int *__restrict__ touched, int touch_cap, int max_recs, long long *__restrict__ out_place,
                       double *__restrict__ out_rating, int *__restrict__ out_count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * RATE_WARPS + warp, nslots = gridDim.x * RATE_WARPS;
    double *mynum = num + (size_t)slot * rdim;
    double *myden = den + (size_t)slot * rdim;
    int *mytouched = touched + (size_t)slot * touch_cap;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    __shared__ int s_keys[RATE_WARPS][RATE_HASH];
    volatile int *keys = s_keys[warp];
    for (int q = lane; q < RATE_HASH; q += 32) keys[q] = -1;
    __syncwarp();
    for (int tt = slot; tt < T; tt += nslots) {
        const int n = nb_cnt[tt];
        int ntouched = 0;
        bool use_hash = false;
        if (rdim > 4 * RATE_HASH) {                                    // small rows stay in L2 anyway
            int total_all = 0;
            for (int k0 = 0; k0 < n && total_all <= RATE_HASH_MAX; k0 += 32) {
                int c = 0;
                if (k0 + lane < n) {
                    const int pi = nb_idx[(size_t)tt * K + k0 + lane].idx;
                    c = rrp[pi + 1] - rrp[pi];
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
                total_all += c;
            }
            use_hash = total_all <= RATE_HASH_MAX;
        }
        // 32 neighbours at a time: their (similarity, rating row extent) one per lane; then the ratings of those
        // neighbours as ONE flat list, 32 entries per step whatever the row boundaries are.  Entries of one step
        // that hit the same place are applied in lane order (= ascending neighbour) in successive rounds
        // (__match_any_sync), everything else in parallel: the sums keep their left-to-right order.
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int kk = k0 + lane;
            double l_sim = 0.0;
            int l_rs = 0, l_rn = 0;
            if (kk < n) {
                const Nb nbl = nb_idx[(size_t)tt * K + kk];
                l_sim = nbl.sim;
                l_rs = rrp[nbl.idx];
                l_rn = rrp[nbl.idx + 1] - l_rs;
            }
            int incl = l_rn;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - l_rn;
            for (int j0 = 0; j0 < total; j0 += 32) {
                const int j = j0 + lane;
                // neighbour (lane L of this block) that owns flat entry j: largest L with excl_L <= j
                int L = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int probe = __shfl_sync(0xffffffffu, excl, min(31, L + step));
                    if (L + step < 32 && probe <= j) L += step;
                }
                // rows of length 0 share their offset with the next row: move on to the row that really holds j
                const int ex_l = __shfl_sync(0xffffffffu, excl, L);
                const int rs_l = __shfl_sync(0xffffffffu, l_rs, L);
                const double sim_l = __shfl_sync(0xffffffffu, l_sim, L);
                const bool valid = j < total;
                int pl = -1 - lane;                                   // distinct dummies: no match among idle lanes
                double w = 0.0;
                if (valid) {
                    pl = rpl[rs_l + (j - ex_l)];
                    w = xmul(rv[rs_l + (j - ex_l)], sim_l);           // rating * similarity, :60
                }
                const unsigned same = __match_any_sync(0xffffffffu, pl);
                const int rank = __popc(same & lt_mask);
                int rounds = __popc(same);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, off));
                bool first = false;
                int slot_at = 0;                                      // index of this entry's num / den
                for (int r = 0; r < rounds; ++r) {
                    if (valid && rank == r) {
                        int at = pl;
                        bool fresh = false;
                        if (use_hash) {
                            // find or insert (the lanes of one round hold different places)
                            at = (int)(((unsigned)pl * 2654435761u) >> 22) & (RATE_HASH - 1);
                            for (;;) {
                                int kq = keys[at];
                                if (kq == -1) {
                                    kq = atomicCAS((int *)&keys[at], -1, pl);
                                    if (kq == -1) {
                                        fresh = true;
                                        break;
                                    }
                                }
                                if (kq == pl) break;
                                at = (at + 1) & (RATE_HASH - 1);
                            }
                        }
                        const double dold = myden[at];
                        first = use_hash ? fresh : dold == 0.0;
                        mynum[at] = xadd(mynum[at], w);               // :63
                        myden[at] = xadd(dold, sim_l);                // :64
                        slot_at = at;
                    }
                    __syncwarp();
                }
                const unsigned m = __ballot_sync(0xffffffffu, first);
                if (first) {
                    const int pos = ntouched + __popc(m & lt_mask);
                    if (pos < touch_cap) mytouched[pos] = slot_at;
                }
                ntouched += __popc(m);
            }
            __syncwarp();
        }
        if (ntouched > touch_cap) ntouched = touch_cap;
        __syncwarp();
        // estimated ratings of the touched places; up to 16 per lane stay in registers for the ranking rounds
        // (and the scratch rows are cleared in the same pass), more than that go back through the scratch row
        constexpr int RG = 16;
        const bool in_regs = ntouched <= 32 * RG;
        double rv_[RG];
        int rp_[RG];
#pragma unroll
        for (int q = 0; q < RG; ++q) {
            rv_[q] = nan;
            rp_[q] = 0;
        }
        if (in_regs) {
#pragma unroll
            for (int q = 0; q < RG; ++q) {
                const int j = lane + 32 * q;
                if (j < ntouched) {
                    const int at = mytouched[j];
                    const int pl = use_hash ? keys[at] : at;
                    const double est = xdiv(mynum[at], myden[at]);      // :68
                    const bool ok = !flag || flag[pl];
                    rv_[q] = ok ? est : nan;
                    rp_[q] = pl;
                    mynum[at] = 0.0;
                    myden[at] = 0.0;
                    if (use_hash) keys[at] = -1;
                }
            }
        } else {
            for (int j = lane; j < ntouched; j += 32) {
                int pl = mytouched[j];
                double est = xdiv(mynum[pl], myden[pl]);          // :68
                bool ok = !flag || flag[pl];
                mynum[pl] = ok ? est : nan;
            }
        }
        __syncwarp();
        bool have_last = false;
        double last_val = 0.0;
        long long last_key = 0;
        int count = 0;
        for (int r = 0; r < max_recs; ++r) {
            bool has = false;
            double bv = 0.0;
            long long bk = 0;
            if (in_regs) {
#pragma unroll
                for (int q = 0; q < RG; ++q) {
                    const double x = rv_[q];
                    const int pl = rp_[q];
                    if (x != x) continue;
                    if (have_last && !ranks_before(last_val, last_key, x, pl)) continue;
                    if (!has || ranks_before(x, pl, bv, bk)) {
                        has = true;
                        bv = x;
                        bk = pl;
                    }
                }
            } else {
                for (int j = lane; j < ntouched; j += 32) {
                    int pl = mytouched[j];
                    double x = mynum[pl];
                    if (x != x) continue;
                    if (have_last && !ranks_before(last_val, last_key, x, pl)) continue;
                    if (!has || ranks_before(x, pl, bv, bk)) {
                        has = true;
                        bv = x;
                        bk = pl;
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                int oh = __shfl_xor_sync(0xffffffffu, (int)has, off);
                double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                long long ok2 = __shfl_xor_sync(0xffffffffu, bk, off);
                if (oh && (!has || ranks_before(ov, ok2, bv, bk))) {
                    has = true;
                    bv = ov;
                    bk = ok2;
                }
            }
            if (!has) break;
            have_last = true;
            last_val = bv;
            last_key = bk;
            if (lane == 0) {
                out_place[(size_t)tt * max_recs + r] = bk;
                out_rating[(size_t)tt * max_recs + r] = bv;
            }
            count = r + 1;
        }
        if (lane == 0) out_count[tt] = count;
        if (!in_regs) {
            for (int j = lane; j < ntouched; j += 32) {
                int pl = mytouched[j];
                mynum[pl] = 0.0;
                myden[pl] = 0.0;
            }
        }
        __syncwarp();
    }
}

// Dense similarity rows for the large-K path: sim[tt][i] for every person i.
__global__ void __launch_bounds__(256)
knn_sim_rows_kernel(KnnDev d, const int *__restrict__ tidx, double pw, double cw, double *__restrict__ sim) {
    const int tt = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.P) return;
    const int t = tidx[tt];
    double s = 0.0;
    if (t >= 0) {
        const TargetRows tr = load_target(d, t);
        s = pair_similarity(d, i, tr, pw, cw);
    }
    sim[(size_t)tt * d.P + i] = s;
}

// orderBy(similarity desc).limit(K) on a dense row (knn/KnnRecommender.scala:46-48): radix select
// of the K-th largest positive similarity, ties taken in ascending index; everything else := 0.
__global__ void __launch_bounds__(1024)
knn_select_mask_kernel(double *__restrict__ sim, long long P, int K) {
    double *row = sim + (size_t)blockIdx.x * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix, s_mask;
    __shared__ long long s_remaining;
    __shared__ long long s_count;
    __shared__ long long s_wcnt[32];
    if (tid == 0) s_count = 0;
    __syncthreads();
    long long c = 0;
    for (long long i = tid; i < P; i += 1024) c += row[i] > 0;
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) atomicAdd((unsigned long long *)&s_count, (unsigned long long)c);
    __syncthreads();
    if ((long long)K >= s_count) return;
    if (tid == 0) {
        s_prefix = 0;
        s_mask = 0;
        s_remaining = K;
    }
    for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix, mask = s_mask;
        for (long long i = tid; i < P; i += 1024) {
            double v = row[i];
            unsigned long long b = (unsigned long long)__double_as_longlong(v);
            if (v > 0 && (b & mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            long long cum = 0, rem = s_remaining;
            int bin = 255;
            for (; bin > 0; --bin) {
                if (cum + (long long)hist[bin] >= rem) break;
                cum += hist[bin];
            }
            s_remaining = rem - cum;
            s_prefix = prefix | ((unsigned long long)bin << shift);
            s_mask = mask | (0xffULL << shift);
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    const long long need = s_remaining;
    // ties in ascending index: warp w owns the contiguous range [P*w/32, P*(w+1)/32)
    const long long lo = P * warp / 32, hi = P * (warp + 1) / 32;
    long long eqc = 0;
    for (long long i = lo + lane; i < hi; i += 32) {
        double v = row[i];
        eqc += (v > 0 && (unsigned long long)__double_as_longlong(v) == kth);
    }
    for (int off = 16; off > 0; off >>= 1) eqc += __shfl_xor_sync(0xffffffffu, eqc, off);
    if (lane == 0) s_wcnt[warp] = eqc;
    __syncthreads();
    long long base = 0;
    for (int w = 0; w < warp; ++w) base += s_wcnt[w];
    const unsigned lt_mask = (1u << lane) - 1u;
    for (long long i0 = lo; i0 < hi; i0 += 32) {
        long long i = i0 + lane;
        double v = i < hi ? row[i] : 0.0;
        unsigned long long b = (unsigned long long)__double_as_longlong(v);
        bool eq = v > 0 && b == kth;
        bool gt = v > 0 && b > kth;
        unsigned m = __ballot_sync(0xffffffffu, eq);
        long long rank = base + __popc(m & lt_mask);
        bool keep = gt || (eq && rank < need);
        if (i < hi && !keep && v != 0.0) row[i] = 0.0;
        base += __popc(m);
    }
}

// makeRecommendations0 by column scan (large K): one thread per place walks the persons that
// rated it in ascending person index; a person contributes iff its masked similarity is > 0.
__global__ void __launch_bounds__(128)
knn_rate_cols_kernel(const int *__restrict__ colptr, const int *__restrict__ cper, const double *__restrict__ crv,
                     int rdim, const double *__restrict__ simrow, double *__restrict__ est) {
    // One WARP per place.  The two sums keep their left-to-right order over the persons (ascending index), but
    // the loads no longer sit on that serial chain: 32 raters' (index, rating, similarity) are fetched by the
    // 32 lanes at once and then folded in lane order.  (With one thread per place a place with 175 K raters
    // took 14 ms on its own.)
    __shared__ double sw[4][32], ss[4][32];          // blockDim.x == 128
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pl = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (pl >= rdim) return;
    double num = 0.0, den = 0.0;
    bool any = false;
    const int e0 = colptr[pl], e1 = colptr[pl + 1];
    // two chunks ahead: rater indices; one chunk ahead: their similarities and ratings
    int p1 = (e0 + lane < e1) ? cper[e0 + lane] : -1;
    int p2 = (e0 + 32 + lane < e1) ? cper[e0 + 32 + lane] : -1;
    double s1 = p1 >= 0 ? simrow[p1] : 0.0;
    double r1 = p1 >= 0 ? crv[e0 + lane] : 0.0;
    for (int eb = e0; eb < e1; eb += 32) {
        const double s = s1, w = xmul(r1, s1);
        const int p3 = (eb + 64 + lane < e1) ? cper[eb + 64 + lane] : -1;
        s1 = p2 >= 0 ? simrow[p2] : 0.0;
        r1 = p2 >= 0 ? crv[eb + 32 + lane] : 0.0;
        p2 = p3;
        const unsigned pos = __ballot_sync(0xffffffffu, s > 0);
        // fold in ascending order by ONE lane from shared memory: the serial chain is then just the two
        // dependent additions per rater (a rater that is no neighbour has s = +0.0 and w = rating * 0.0 = +0.0,
        // which leave the sums unchanged, so no test sits on the chain)
        sw[wid][lane] = w;
        ss[wid][lane] = s;
        __syncwarp();
        if (lane == 0 && pos) {
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                num = xadd(num, sw[wid][l]);
                den = xadd(den, ss[wid][l]);
            }
        }
        __syncwarp();
        any = any || pos != 0u;
    }
    if (lane == 0) est[pl] = any ? xdiv(num, den) : __longlong_as_double(0x7ff8000000000000LL);
}

__global__ void knn_fill_int_stride_kernel(int *p, int n, int stride, int offset, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[(size_t)i * stride + offset] = v;
}

__global__ void knn_fill_int_kernel(int *p, long long n, int v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// ------------------------------------------------------------------ host object
struct vrec_knn {
    vrec_ctx *ctx = nullptr;
    cudaEvent_t ev_dense[2] = {nullptr, nullptr};   // around the last dense-filter kernel launch (bench roofline)
    int64_t P = 0;
    int place_dim = 0, cat_dim = 0, rdim = 0;
    int max_rating_row = 0;
    int64_t nnz_place = 0, nnz_cat = 0, nnz_rat = 0;
    std::vector<int64_t> h_person;
    DevBuf<long long> d_person;
    DevBuf<int> d_prp, d_pci, d_crp, d_cci;
    DevBuf<double> d_pv, d_plen, d_cv, d_clen;
    DevBuf<int> d_rrp, d_rpl;          // ratings by person
    DevBuf<double> d_rv;
    DevBuf<int> d_ccp, d_cper;         // ratings by place (CSC)
    DevBuf<double> d_crv;
    DevBuf<unsigned char> d_flag;
    bool has_filter = false;
    // tiled batch kernel: dense fp32 features, head-place slots, place postings
    bool tile_ok = false;
    int n_head = 0;
    long long fstride = 0;
    DevBuf<float> d_feat;
    DevBuf<short> d_head_slot;
    DevBuf<int> d_pcp, d_pper;
    DevBuf<double> d_seed_thr;
    bool rec_f32 = false;                    // packed records hold their values as floats
    DevBuf<int> d_post_bin, d_post_order, d_post_bins;   // longest-first order of the postings kernel's targets
    int64_t opt_post_first = 0;               // 1: postings kernel before the dense kernel, its lists seed the thresholds (measured: 187 instead of 338 dense survivors per target, but the postings kernel loses its cut-off: 18.9 ms vs 17.7 ms per step)
    DevBuf<double> d_tdense;                 // [targets of the batch][cat_dim] dense category rows (knn_tc_ws_kernel)
    DevBuf<int> d_work;
    DevBuf<unsigned long long> d_meta;     // packed records for the exact evaluation
    DevBuf<double> d_rec;
    DevBuf<unsigned long long> d_cmeta, d_crec;   // compact records (small-integer region-sets), see CRec
    int64_t opt_compact = 1;               // 0: evaluate on the full-size records even when compact ones exist (A/B)
    void attach_compact(TileAux &aux) const {
        aux.plen = d_plen.p;
        aux.clen = d_clen.p;
        if (opt_compact && d_cmeta.p && d_crec.p && d_meta.p) {
            aux.cmeta = d_cmeta.p;
            aux.crec = d_crec.p;
        }
    }
    // tensor-core variant: fp16 row-major features over TC_D dims, its own (larger) head set
    bool tc_ok = false;
    DevBuf<__half> d_feat16;
    DevBuf<unsigned char> d_featsw;          // the same features as 32 KB swizzled tile images
    int tc_tiles = 0;
    DevBuf<short> d_head_slot_tc;
    // options
    int64_t opt_rating_path = 0, opt_tile = 0, opt_splits = 0, opt_kernel = 0;
    int64_t opt_debug_skip_postings = 0;     // timing experiments only: results are then WRONG
    int64_t opt_tc_seed = 0;                 // run the sampled seed pass before the tensor-core main pass
    // scratch
    DevBuf<long long> d_targets;
    DevBuf<int> d_tidx, d_status;
    DevBuf<Nb> d_part, d_nb_rank, d_nb_idx;
    DevBuf<int> d_part_cnt, d_nb_cnt;
    DevBuf<double> d_num, d_den;
    DevBuf<int> d_touched;
    int rate_slots = 0, rate_touch_cap = 0;
    DevBuf<double> d_sim, d_est;
    DevBuf<long long> d_out_place;
    DevBuf<double> d_out_rating;
    DevBuf<int> d_out_count;

    KnnDev dev() const {
        KnnDev d;
        d.P = P;
        d.person = d_person.p;
        d.place = KnnVec{d_prp.p, d_pci.p, d_pv.p, d_plen.p};
        d.cat = KnnVec{d_crp.p, d_cci.p, d_cv.p, d_clen.p};
        return d;
    }
};

namespace {

int check_csr(const char *what, int64_t P, const int64_t *rowptr, const int32_t *col, int32_t dim) {
    if (rowptr[0] != 0) {
        vrec_set_error("vrec_knn_load: %s rowptr[0] != 0", what);
        return VREC_EINVAL;
    }
    for (int64_t i = 0; i < P; ++i) {
        if (rowptr[i + 1] < rowptr[i]) {
            vrec_set_error("vrec_knn_load: %s rowptr not monotone at row %lld", what, (long long)i);
            return VREC_EINVAL;
        }
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (col[k] < 0 || col[k] >= dim || (k > rowptr[i] && col[k] <= col[k - 1])) {
                vrec_set_error("vrec_knn_load: %s row %lld: indices must be strictly ascending in [0, %d)", what,
                               (long long)i, dim);
                return VREC_EINVAL;
            }
        }
    }
    if (rowptr[P] >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_knn_load: %s has too many non-zeros", what);
        return VREC_EINVAL;
    }
    return VREC_OK;
}

// permuted CSR copy (rows reordered by `order`), int32 row pointers
void permute_csr(int64_t P, const std::vector<int64_t> &order, const int64_t *rowptr, const int32_t *col,
                 const double *val, std::vector<int> &orp, std::vector<int> &oc, std::vector<double> &ov) {
    orp.assign((size_t)P + 1, 0);
    oc.resize((size_t)rowptr[P]);
    ov.resize((size_t)rowptr[P]);
    int pos = 0;
    for (int64_t r = 0; r < P; ++r) {
        int64_t src = order[r];
        orp[r] = pos;
        for (int64_t k = rowptr[src]; k < rowptr[src + 1]; ++k) {
            oc[pos] = col[k];
            ov[pos] = val[k];
            ++pos;
        }
    }
    orp[P] = pos;
}

}  // namespace

extern "C" int vrec_knn_load(vrec_ctx *ctx, int64_t P, const int64_t *person_id, const int64_t *place_rowptr,
                             const int32_t *place_col, const double *place_val, int32_t place_dim,
                             const int64_t *cat_rowptr, const int32_t *cat_col, const double *cat_val,
                             int32_t cat_dim, int64_t n_ratings, const int64_t *rating_person,
                             const int64_t *rating_place, const int64_t *rating_value, vrec_knn **out) {
    if (!ctx || !out || P < 0 || (P > 0 && !person_id) || !place_rowptr || !cat_rowptr || place_dim < 0 ||
        cat_dim < 0 || P >= (int64_t)0x7fffffff) {
        vrec_set_error("vrec_knn_load: bad argument");
        return VREC_EINVAL;
    }
    *out = nullptr;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(check_csr("place_rating_vectors", P, place_rowptr, place_col, place_dim));
    VREC_TRY(check_csr("category_rating_vectors", P, cat_rowptr, cat_col, cat_dim));
    // persons in ascending id
    std::vector<int64_t> order((size_t)P);
    std::iota(order.begin(), order.end(), (int64_t)0);
    bool sorted = true;
    for (int64_t i = 1; i < P; ++i)
        if (person_id[i] <= person_id[i - 1]) sorted = false;
    if (!sorted) std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return person_id[a] < person_id[b]; });
    vrec_knn *k = new vrec_knn();
    k->ctx = ctx;
    k->P = P;
    k->place_dim = place_dim;
    k->cat_dim = cat_dim;
    k->h_person.resize((size_t)P);
    for (int64_t i = 0; i < P; ++i) k->h_person[i] = person_id[order[i]];
    for (int64_t i = 1; i < P; ++i)
        if (k->h_person[i] == k->h_person[i - 1]) {
            vrec_set_error("vrec_knn_load: duplicate person_id %lld", (long long)k->h_person[i]);
            delete k;
            return VREC_EINVAL;
        }
    std::vector<int> prp, pci, crp, cci;
    std::vector<double> pv, cv;
    permute_csr(P, order, place_rowptr, place_col, place_val, prp, pci, pv);
    permute_csr(P, order, cat_rowptr, cat_col, cat_val, crp, cci, cv);
    k->nnz_place = prp[P];
    k->nnz_cat = crp[P];
    // place_ratings grouped by person
    std::vector<int> rrp, rpl;
    std::vector<double> rv;
    int rdim = place_dim;
    if (!rating_person) {
        rrp = prp;
        rpl = pci;
        rv = pv;
    } else {
        if (n_ratings < 0 || n_ratings >= (int64_t)0x7fffffff || !rating_place || !rating_value) {
            vrec_set_error("vrec_knn_load: bad place_ratings arguments");
            delete k;
            return VREC_EINVAL;
        }
        std::vector<int> pidx((size_t)n_ratings);
        rrp.assign((size_t)P + 1, 0);
        for (int64_t e = 0; e < n_ratings; ++e) {
            auto it = std::lower_bound(k->h_person.begin(), k->h_person.end(), rating_person[e]);
            int idx = (it != k->h_person.end() && *it == rating_person[e]) ? (int)(it - k->h_person.begin()) : -1;
            pidx[e] = idx;       // persons without rating vectors can never be neighbours: dropped
            if (idx >= 0) {
                if (rating_place[e] < 0 || rating_place[e] >= (int64_t)0x7ffffff0) {
                    vrec_set_error("vrec_knn_load: place_id %lld out of Int range", (long long)rating_place[e]);
                    delete k;
                    return VREC_EINVAL;
                }
                rrp[idx + 1]++;
                rdim = std::max<int64_t>(rdim, rating_place[e] + 1);
            }
        }
        for (int64_t i = 0; i < P; ++i) rrp[i + 1] += rrp[i];
        std::vector<int> pos(rrp.begin(), rrp.end() - 1);
        rpl.resize((size_t)rrp[P]);
        rv.resize((size_t)rrp[P]);
        for (int64_t e = 0; e < n_ratings; ++e) {
            if (pidx[e] < 0) continue;
            int p = pos[pidx[e]]++;
            rpl[p] = (int)rating_place[e];
            rv[p] = (double)rating_value[e];           // long -> double, as `rating * similarity` does
        }
    }
    k->rdim = std::max(rdim, 1);
    k->nnz_rat = rrp[P];
    // by place (CSC), persons ascending inside a column; reject duplicate (person, place) rows
    std::vector<int> ccp((size_t)k->rdim + 1, 0), cper((size_t)k->nnz_rat);
    std::vector<double> crv((size_t)k->nnz_rat);
    for (int64_t e = 0; e < k->nnz_rat; ++e) ccp[rpl[e] + 1]++;
    for (int64_t c = 0; c < k->rdim; ++c) ccp[c + 1] += ccp[c];
    {
        std::vector<int> pos(ccp.begin(), ccp.end() - 1);
        for (int64_t i = 0; i < P; ++i) {
            k->max_rating_row = std::max(k->max_rating_row, rrp[i + 1] - rrp[i]);
            for (int e = rrp[i]; e < rrp[i + 1]; ++e) {
                int c = rpl[e];
                int p = pos[c]++;
                if (p > ccp[c] && cper[p - 1] == (int)i) {
                    vrec_set_error("vrec_knn_load: duplicate (person_id, place_id) = (%lld, %d) in place_ratings",
                                   (long long)k->h_person[i], c);
                    delete k;
                    return VREC_EINVAL;
                }
                cper[p] = (int)i;
                crv[p] = rv[e];
            }
        }
    }
    cudaStream_t s = ctx->stream;
    int rc = k->d_person.upload((const long long *)k->h_person.data(), (size_t)P, s);
    if (rc == VREC_OK) rc = k->d_prp.upload(prp.data(), prp.size(), s);
    if (rc == VREC_OK) rc = k->d_pci.upload(pci.data(), pci.size(), s);
    if (rc == VREC_OK) rc = k->d_pv.upload(pv.data(), pv.size(), s);
    if (rc == VREC_OK) rc = k->d_crp.upload(crp.data(), crp.size(), s);
    if (rc == VREC_OK) rc = k->d_cci.upload(cci.data(), cci.size(), s);
    if (rc == VREC_OK) rc = k->d_cv.upload(cv.data(), cv.size(), s);
    if (rc == VREC_OK) rc = k->d_rrp.upload(rrp.data(), rrp.size(), s);
    if (rc == VREC_OK) rc = k->d_rpl.upload(rpl.data(), rpl.size(), s);
    if (rc == VREC_OK) rc = k->d_rv.upload(rv.data(), rv.size(), s);
    if (rc == VREC_OK) rc = k->d_ccp.upload(ccp.data(), ccp.size(), s);
    if (rc == VREC_OK) rc = k->d_cper.upload(cper.data(), cper.size(), s);
    if (rc == VREC_OK) rc = k->d_crv.upload(crv.data(), crv.size(), s);
    if (rc == VREC_OK) rc = k->d_plen.alloc((size_t)P);
    if (rc == VREC_OK) rc = k->d_clen.alloc((size_t)P);
    if (rc == VREC_OK) rc = k->d_flag.alloc((size_t)k->rdim);
    if (rc == VREC_OK && P > 0) {
        int grid = (int)((P + 255) / 256);
        knn_norms_kernel<<<grid, 256, 0, s>>>(k->d_prp.p, k->d_pv.p, P, k->d_plen.p);
        knn_norms_kernel<<<grid, 256, 0, s>>>(k->d_crp.p, k->d_cv.p, P, k->d_clen.p);
        ctx->launches += 2;
        if (cudaGetLastError() != cudaSuccess) rc = VREC_ECUDA;
    }
    // ---- inputs of the tiled batch kernel
    bool nonneg = true;
    for (double v : pv) nonneg &= v >= 0;
    for (double v : cv) nonneg &= v >= 0;
    if (rc == VREC_OK && nonneg && cat_dim <= TILE_D && P > 0 && place_dim < (1 << 30)) {
        // postings of the place vectors (persons ascending) and the n_head most visited places
        std::vector<int> pcp((size_t)place_dim + 1, 0), pper((size_t)k->nnz_place);
        for (int64_t e = 0; e < k->nnz_place; ++e) pcp[pci[e] + 1]++;
        for (int64_t c = 0; c < place_dim; ++c) pcp[c + 1] += pcp[c];
        {
            std::vector<int> pos(pcp.begin(), pcp.end() - 1);
            for (int64_t i = 0; i < P; ++i)
                for (int e = prp[i]; e < prp[i + 1]; ++e) pper[pos[pci[e]]++] = (int)i;
        }
        k->n_head = std::min<int>(TILE_D - cat_dim, place_dim);
        std::vector<int> byc((size_t)place_dim);
        std::iota(byc.begin(), byc.end(), 0);
        std::partial_sort(byc.begin(), byc.begin() + k->n_head, byc.end(), [&](int a, int b) {
            int ca = pcp[a + 1] - pcp[a], cb = pcp[b + 1] - pcp[b];
            return ca > cb || (ca == cb && a < b);
        });
        std::vector<short> head_slot((size_t)place_dim, (short)-1);
        for (int h = 0; h < k->n_head; ++h) head_slot[byc[h]] = (short)h;
        k->fstride = (P + 31) / 32 * 32;
        rc = k->d_pcp.upload(pcp.data(), pcp.size(), s);
        if (rc == VREC_OK) rc = k->d_pper.upload(pper.data(), pper.size(), s);
        if (rc == VREC_OK) rc = k->d_head_slot.upload(head_slot.data(), head_slot.size(), s);
        if (rc == VREC_OK) rc = k->d_feat.alloc((size_t)TILE_D * (size_t)k->fstride);
        if (rc == VREC_OK) {
            knn_features_kernel<<<(int)((P + 255) / 256), 256, 0, s>>>(k->dev(), k->d_head_slot.p, cat_dim,
                                                                      k->fstride, k->d_feat.p);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = VREC_ECUDA;
        }
        k->tile_ok = rc == VREC_OK;
        std::vector<short> hs_tc((size_t)place_dim, (short)-1);
        if (rc == VREC_OK && cat_dim <= 64) {
            int n_head_tc = std::min<int>(TC_D - cat_dim, place_dim);
            std::partial_sort(byc.begin(), byc.begin() + n_head_tc, byc.end(), [&](int a, int b) {
                int ca = pcp[a + 1] - pcp[a], cb = pcp[b + 1] - pcp[b];
                return ca > cb || (ca == cb && a < b);
            });
            std::vector<short> &hs = hs_tc;
            for (int h = 0; h < n_head_tc; ++h) hs[byc[h]] = (short)h;
            rc = k->d_head_slot_tc.upload(hs.data(), hs.size(), s);
#if VREC_WITH_TC_BASELINE
            if (rc == VREC_OK) rc = k->d_feat16.alloc((size_t)TC_D * (size_t)P);
            if (rc == VREC_OK) {
                knn_features16_kernel<<<(int)((P + 127) / 128), 128, 0, s>>>(k->dev(), k->d_head_slot_tc.p, cat_dim,
                                                                            k->d_feat16.p);
                ctx->launches++;
                if (cudaGetLastError() != cudaSuccess) rc = VREC_ECUDA;
            }
#endif
            if (rc == VREC_OK) {
                k->tc_tiles = (int)((P + TC_N - 1) / TC_N);
                rc = k->d_featsw.alloc((size_t)k->tc_tiles * TC_TILE_BYTES);
                if (rc == VREC_OK && cudaMemsetAsync(k->d_featsw.p, 0, (size_t)k->tc_tiles * TC_TILE_BYTES, s) != cudaSuccess)
                    rc = VREC_ECUDA;
                if (rc == VREC_OK) {
                    knn_features16sw_kernel<<<(int)((P + 127) / 128), 128, 0, s>>>(k->dev(), k->d_head_slot_tc.p, cat_dim,
                                                                                  k->d_featsw.p);
                    ctx->launches++;
                    if (cudaGetLastError() != cudaSuccess) rc = VREC_ECUDA;
                }
            }
            k->tc_ok = rc == VREC_OK;
        }
        // packed records (needs the device-computed lengths): one contiguous block per person
        bool small_rows = true;
        for (int64_t i = 0; i < P && small_rows; ++i)
            small_rows = prp[i + 1] - prp[i] < 4096 && crp[i + 1] - crp[i] < 4096;
        if (rc == VREC_OK && small_rows) {
            std::vector<double> hpl((size_t)P), hcl((size_t)P);
            if (cudaMemcpyAsync(hpl.data(), k->d_plen.p, sizeof(double) * (size_t)P, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaMemcpyAsync(hcl.data(), k->d_clen.p, sizeof(double) * (size_t)P, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess) {
                vrec_set_error("vrec_knn_load: %s", cudaGetErrorString(cudaGetLastError()));
                rc = VREC_ECUDA;
            }
            std::vector<unsigned long long> meta((size_t)P);
            std::vector<double> rec;
            rec.reserve((size_t)(10 * P + 2 * (k->nnz_place + k->nnz_cat)));
            // values as floats when that loses nothing (visit counts: always)
            bool F32 = true;
            for (int64_t e = 0; e < k->nnz_place && F32; ++e) F32 = (double)(float)pv[e] == pv[e];
            for (int64_t e = 0; e < k->nnz_cat && F32; ++e) F32 = (double)(float)cv[e] == cv[e];
            k->rec_f32 = F32;
            for (int64_t i = 0; i < P && rc == VREC_OK; ++i) {
                const int np = prp[i + 1] - prp[i], nc = crp[i + 1] - crp[i];
                meta[i] = (unsigned long long)rec.size() | ((unsigned long long)np << 40) | ((unsigned long long)nc << 52);
                rec.push_back(hpl[i]);
                rec.push_back(hcl[i]);
                size_t at = rec.size();
                rec.resize(at + (size_t)rec_cols_words(np), 0.0);
                unsigned *pc = reinterpret_cast<unsigned *>(rec.data() + at);
                for (int e = 0; e < np; ++e) {
                    int col = pci[prp[i] + e];
                    pc[e] = (unsigned)col | (hs_tc[col] < 0 ? REC_TAIL_TC : 0u) | (head_slot[col] < 0 ? REC_TAIL_TILE : 0u);
                }
                at = rec.size();
                rec.resize(at + (size_t)rec_vals_words(np, F32), 0.0);
                if (F32) {
                    float *fv = reinterpret_cast<float *>(rec.data() + at);
                    for (int e = 0; e < np; ++e) fv[e] = (float)pv[prp[i] + e];
                } else {
                    for (int e = 0; e < np; ++e) rec[at + e] = pv[prp[i] + e];
                }
                at = rec.size();
                rec.resize(at + (size_t)rec_cols_words(nc), 0.0);
                unsigned *cc = reinterpret_cast<unsigned *>(rec.data() + at);
                for (int e = 0; e < nc; ++e) cc[e] = (unsigned)cci[crp[i] + e];
                at = rec.size();
                rec.resize(at + (size_t)rec_vals_words(nc, F32), 0.0);
                if (F32) {
                    float *fv = reinterpret_cast<float *>(rec.data() + at);
                    for (int e = 0; e < nc; ++e) fv[e] = (float)cv[crp[i] + e];
                } else {
                    for (int e = 0; e < nc; ++e) rec[at + e] = cv[crp[i] + e];
                }
            }
            rec.resize(rec.size() + 64, 0.0);         // slack: the 128-bit loads may touch the padding of the last record
            if (rc == VREC_OK) rc = k->d_meta.upload(meta.data(), meta.size(), s);
            if (rc == VREC_OK) rc = k->d_rec.upload(rec.data(), rec.size(), s);
            if (rc == VREC_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = VREC_ECUDA;
            if (rc != VREC_OK) {
                k->d_meta.release();
                k->d_rec.release();
            }
            // compact records: small-integer values only (visit counts), see CRec
            bool compact = rc == VREC_OK && place_dim <= (1 << 22) && cat_dim <= 64;
            for (int64_t e = 0; e < k->nnz_place && compact; ++e)
                compact = pv[e] >= 0.0 && pv[e] <= 255.0 && pv[e] == (double)(int)pv[e];
            for (int64_t e = 0; e < k->nnz_cat && compact; ++e)
                compact = cv[e] >= 0.0 && cv[e] <= 1023.0 && cv[e] == (double)(int)cv[e];
            if (compact) {
                std::vector<unsigned long long> cmeta((size_t)P), crec;
                crec.reserve((size_t)(k->nnz_place / 2 + k->nnz_cat / 4 + 2 * P + 16));
                for (int64_t i = 0; i < P; ++i) {
                    const int np = prp[i + 1] - prp[i], nc = crp[i + 1] - crp[i];
                    cmeta[i] = (unsigned long long)crec.size() | ((unsigned long long)np << 40) |
                               ((unsigned long long)nc << 52);
                    size_t at = crec.size();
                    crec.resize(at + (size_t)crec_place_words(np), 0ULL);
                    unsigned *pe = reinterpret_cast<unsigned *>(crec.data() + at);
                    for (int e = 0; e < np; ++e) {
                        const int col = pci[prp[i] + e];
                        pe[e] = (unsigned)col | ((unsigned)(int)pv[prp[i] + e] << 22) | (hs_tc[col] < 0 ? REC_TAIL_TC : 0u) |
                                (head_slot[col] < 0 ? REC_TAIL_TILE : 0u);
                    }
                    at = crec.size();
                    crec.resize(at + (size_t)crec_cat_words(nc), 0ULL);
                    unsigned short *ce = reinterpret_cast<unsigned short *>(crec.data() + at);
                    for (int e = 0; e < nc; ++e)
                        ce[e] = (unsigned short)((unsigned)cci[crp[i] + e] | ((unsigned)(int)cv[crp[i] + e] << 6));
                }
                crec.resize(crec.size() + 16, 0ULL);
                int rc2 = k->d_cmeta.upload(cmeta.data(), cmeta.size(), s);
                if (rc2 == VREC_OK) rc2 = k->d_crec.upload(crec.data(), crec.size(), s);
                if (rc2 != VREC_OK || cudaStreamSynchronize(s) != cudaSuccess) {
                    k->d_cmeta.release();
                    k->d_crec.release();
                }
            }
        }
    }
    if (rc == VREC_OK && cudaStreamSynchronize(s) != cudaSuccess) {
        vrec_set_error("vrec_knn_load: %s", cudaGetErrorString(cudaGetLastError()));
        rc = VREC_ECUDA;
    }
    if (rc != VREC_OK) {
        delete k;
        return rc;
    }
    *out = k;
    return VREC_OK;
}

extern "C" void vrec_knn_free(vrec_knn *knn) {
    if (!knn) return;
    cudaSetDevice(knn->ctx->device);
    cudaStreamSynchronize(knn->ctx->stream);
    for (cudaEvent_t e : knn->ev_dense)
        if (e) cudaEventDestroy(e);
    delete knn;
}

extern "C" int64_t vrec_knn_resident_bytes(vrec_knn *k) {
    if (!k) return 0;
    return (int64_t)(k->d_person.bytes() + k->d_prp.bytes() + k->d_pci.bytes() + k->d_pv.bytes() +
                     k->d_plen.bytes() + k->d_crp.bytes() + k->d_cci.bytes() + k->d_cv.bytes() +
                     k->d_clen.bytes());
}

extern "C" int vrec_knn_person_ids(vrec_knn *k, int64_t *out) {
    if (!k || !out) return VREC_EINVAL;
    std::copy(k->h_person.begin(), k->h_person.end(), out);
    return VREC_OK;
}

extern "C" int vrec_knn_set_option(vrec_knn *k, const char *name, int64_t value) {
    if (!k || !name) return VREC_EINVAL;
    if (!strcmp(name, "rating_path") && value >= 0 && value <= 2) {
        k->opt_rating_path = value;
        return VREC_OK;
    }
    if (!strcmp(name, "tile") && value >= 0) {
        k->opt_tile = value;
        return VREC_OK;
    }
    if (!strcmp(name, "knn_kernel") && value >= 0 && value <= 4) {
        if (value >= 3 && !k->tc_ok) {
            vrec_set_error("knn_kernel=3 (tensor cores) needs cat_dim <= 64 and non-negative rating values");
            return VREC_EINVAL;
        }
        if (value == 2 && !k->tile_ok) {
            vrec_set_error("knn_kernel=2 (tiled) needs cat_dim <= 32 and non-negative rating values");
            return VREC_EINVAL;
        }
        k->opt_kernel = value;
        return VREC_OK;
    }
    if (!strcmp(name, "tc_seed")) {
        k->opt_tc_seed = value;
        return VREC_OK;
    }
    if (!strcmp(name, "post_first")) {
        k->opt_post_first = value;
        return VREC_OK;
    }
    if (!strcmp(name, "debug_skip_postings")) {
        k->opt_debug_skip_postings = value;
        return VREC_OK;
    }
    if (!strcmp(name, "compact_records") && (value == 0 || value == 1)) {
        k->opt_compact = value;            // 0: exact evaluations on the full-size records (A/B of the compact ones)
        return VREC_OK;
    }
    if (!strcmp(name, "splits") && value >= 0 && value <= 32) {
        k->opt_splits = value;
        return VREC_OK;
    }
    vrec_set_error("vrec_knn_set_option: unknown option or bad value: %s=%lld", name, (long long)value);
    return VREC_EINVAL;
}

// Debug: reads and clears the event counters of the tiled kernel (see g_tile_stats).
// Duration of the last knn_tc_ws_kernel launch (CUDA events on the library's stream), in milliseconds;
// 0 if that kernel has not run.  Synchronises the stream.
extern "C" int vrec_knn_last_dense_ms(vrec_knn *k, double *out_ms) {
    if (!k || !out_ms) return VREC_EINVAL;
    *out_ms = 0.0;
    if (!k->ev_dense[0]) return VREC_OK;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    float ms = 0.0f;
    VREC_CUDA(cudaEventElapsedTime(&ms, k->ev_dense[0], k->ev_dense[1]));
    *out_ms = ms;
    return VREC_OK;
}

extern "C" int vrec_knn_debug_stats(vrec_knn *k, uint64_t *out4) {
    if (!k || !out4) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    unsigned long long z[4] = {0, 0, 0, 0};
    VREC_CUDA(cudaMemcpyFromSymbol(out4, g_tile_stats, sizeof(z)));
    VREC_CUDA(cudaMemcpyToSymbol(g_tile_stats, z, sizeof(z)));
    return VREC_OK;
}

// Debug: the neighbour list (ascending person index, as the rating reduction reads it) of target t of the
// last vrec_knn_query pass over the fused top-K kernels.  capacity >= the K of that query.
extern "C" int vrec_knn_debug_last_neighbours(vrec_knn *k, int32_t t, int32_t K, int64_t *out_person_id,
                                              double *out_similarity, int32_t *out_count) {
    if (!k || !out_person_id || !out_similarity || !out_count || t < 0 || K <= 0) return VREC_EINVAL;
    if ((size_t)(t + 1) * (size_t)K > k->d_nb_idx.n || (size_t)t >= k->d_nb_cnt.n) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    int cnt = 0;
    std::vector<Nb> h((size_t)K);
    VREC_CUDA(cudaMemcpyAsync(&cnt, k->d_nb_cnt.p + t, sizeof(int), cudaMemcpyDeviceToHost, k->ctx->stream));
    VREC_CUDA(cudaMemcpyAsync(h.data(), k->d_nb_idx.p + (size_t)t * K, sizeof(Nb) * (size_t)K, cudaMemcpyDeviceToHost,
                              k->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    cnt = std::max(0, std::min(cnt, (int)K));
    for (int i = 0; i < cnt; ++i) {
        out_person_id[i] = k->h_person[(size_t)h[i].idx];
        out_similarity[i] = h[i].sim;
    }
    *out_count = cnt;
    return VREC_OK;
}

// Debug: reads and clears the exact-evaluation probe (block 0 warp 0, postings pass): out4 = {meta wait,
// record + place merge + division, category merge + division, evaluations}.
extern "C" int vrec_knn_debug_probe(vrec_knn *k, uint64_t *out8) {
    uint64_t *out4 = out8;
    if (!k || !out4) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    VREC_CUDA(cudaMemcpyFromSymbol(out4, g_probe, 64));        // 8 values
    VREC_CUDA(cudaMemcpyToSymbol(g_probe, z, sizeof(z)));
    return VREC_OK;
}

// Debug: per-block cycles of the last tensor-core main pass: out[0..n) dense phase, out[n..2n) postings phase.
extern "C" int vrec_knn_debug_tc_block_cycles(vrec_knn *k, uint64_t *out, int n) {
    if (!k || !out || n < 0 || n > 1024) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    VREC_CUDA(cudaMemcpyFromSymbol(out, g_tc_block_cycles, sizeof(uint64_t) * n, 0));
    VREC_CUDA(cudaMemcpyFromSymbol(out + n, g_tc_block_cycles, sizeof(uint64_t) * n, sizeof(uint64_t) * 1024));
    return VREC_OK;
}

// Debug: reads and clears the phase cycle counters of knn_tc_kernel (see g_tc_cycles).
extern "C" int vrec_knn_debug_tc_cycles(vrec_knn *k, uint64_t *out8) {
    if (!k || !out8) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    VREC_CUDA(cudaMemcpyFromSymbol(out8, g_tc_cycles, sizeof(z)));       // out8 holds 12 values
    VREC_CUDA(cudaMemcpyToSymbol(g_tc_cycles, z, sizeof(z)));
    return VREC_OK;
}

extern "C" int vrec_knn_set_filter(vrec_knn *k, const int64_t *place_filter, int64_t n_filter) {
    if (!k || n_filter < 0) return VREC_EINVAL;
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    if (!place_filter) {
        k->has_filter = false;
        return VREC_OK;
    }
    std::vector<unsigned char> flag((size_t)k->rdim, 0);
    for (int64_t i = 0; i < n_filter; ++i)
        if (place_filter[i] >= 0 && place_filter[i] < k->rdim) flag[(size_t)place_filter[i]] = 1;
    VREC_CUDA(cudaMemcpyAsync(k->d_flag.p, flag.data(), flag.size(), cudaMemcpyHostToDevice, k->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    k->has_filter = true;
    return VREC_OK;
}

namespace {

int knn_check_params(double pw, double cw, int32_t K) {
    // knn/KnnRecommender.scala:17-20
    if (!(pw > 0 && pw < 1.0)) {
        vrec_set_error("requirement failed: Place weight must be in the interval (0; 1): %g", pw);
        return VREC_EINVAL;
    }
    if (!(cw > 0 && cw < 1.0)) {
        vrec_set_error("requirement failed: Category weight must be in the interval (0; 1): %g", cw);
        return VREC_EINVAL;
    }
    if (!(pw + cw == 1.0)) {
        vrec_set_error("requirement failed: Sum of weights must be 1.0: place: %g, category: %g", pw, cw);
        return VREC_EINVAL;
    }
    if (K <= 0) {
        vrec_set_error("requirement failed: K nearest must be positive");
        return VREC_EINVAL;
    }
    return VREC_OK;
}

bool use_gather_path(const vrec_knn *k, int K) {
    if (K > TOPK_MAX_K) return false;
    if (k->opt_rating_path == 1) return true;
    if (k->opt_rating_path == 2) return false;
    return (int64_t)K * std::max(1, k->max_rating_row) <= (int64_t)(1 << 20);
}

// neighbours of a tile of targets by the fused top-K kernels (K <= 1024)
int knn_run_topk(vrec_knn *k, int tn, double pw, double cw, int K) {
    vrec_ctx *ctx = k->ctx;
    // 0 = automatic (tensor cores > CUDA-core tile > exact scan), 1 = exact scan, 2 = tile,
    // 3 = tensor cores (all warps do everything), 4 = tensor cores, warp-specialised (the automatic choice)
    const bool use_tc = k->tc_ok && K <= TC_MAX_K && (k->opt_kernel == 0 || k->opt_kernel >= 3);
    const bool use_ws = use_tc && (k->opt_kernel != 3 || !VREC_WITH_TC_BASELINE);
    const bool tiled = !use_tc && k->tile_ok && k->opt_kernel != 1;
    const bool filtered = use_tc || tiled;
    // filtered kernels: slot S of every target's partial lists belongs to the postings kernel
    // partial lists of the postings kernel per target: heavy targets are split when the merge buffer has room
    const int post_parts = filtered ? (TOPK_BUF / K >= 2 * POST_MAX_PARTS ? POST_MAX_PARTS : 1) : 0;
    int smax = std::max(1, std::min(32, TOPK_BUF / K - post_parts));
    int S = (int)k->opt_splits;
    int T = 1;
    size_t smem = 0;
    int ws_pool_ints = 0;
    if (use_tc) {
        T = TC_M;
        smem = 1024 + (use_ws ? WS_STAGES : 1 + TC_STAGES) * (size_t)TC_TILE_BYTES + 12 * (size_t)TC_M * K +
               sizeof(unsigned long long) * (use_ws ? WS_QCAP : TC_QCAP) + sizeof(float) * TC_M +
               sizeof(int) * 3 * TC_M + 16 + sizeof(unsigned long long) * TC_M;
        if (use_ws) {
            // what is left of the 227 KB holds the targets' place columns (knn_tc_ws_kernel)
            const size_t limit = 227 * 1024 - 2048;
            ws_pool_ints = smem < limit ? (int)((limit - smem) / sizeof(int)) : 0;
            smem += sizeof(int) * (size_t)ws_pool_ints;
        }
    } else if (tiled) {
        // targets per block: heaps must fit next to the queue and the target vectors
        T = (int)std::max<int64_t>(1, std::min<int64_t>(64, (48 * 1024) / (12 * (int64_t)K)));
        T = std::min(T, std::max(1, tn));
        smem = sizeof(double) * (size_t)T * K + sizeof(unsigned long long) * TILE_QCAP +
               sizeof(float) * (size_t)T * TILE_TVEC_STRIDE + sizeof(int) * (size_t)T * K + sizeof(int) * 3 * T + 16;
    }
    const int tiles = (tn + T - 1) / T;
    if (S <= 0) {
        S = 1;
        int want_blocks = ctx->sm_count * (use_tc ? 1 : tiled ? 4 : 8);
        while (S < smax && tiles * S < want_blocks) S <<= 1;
    }
    S = std::min(S, smax);
    if ((int64_t)S > std::max<int64_t>(1, k->P)) S = 1;
    const int SP = S + post_parts;                       // partial lists per target
    VREC_TRY(k->d_part.ensure((size_t)tn * SP * K));
    VREC_TRY(k->d_part_cnt.ensure((size_t)tn * SP));
    VREC_TRY(k->d_nb_rank.ensure((size_t)tn * K));
    VREC_TRY(k->d_nb_idx.ensure((size_t)tn * K));
    VREC_TRY(k->d_nb_cnt.ensure((size_t)tn));
    if (filtered) {
        VREC_TRY(k->d_seed_thr.ensure((size_t)tn));
        VREC_TRY(k->d_work.ensure(1));
        // seed pass: exact top-K of a strided sample -> lower bound of every target's K-th best
        long long sample = std::min<long long>(k->P, std::max<long long>(4096, std::min<long long>(65536, k->P / 16)));
        long long stride = std::max<long long>(1, k->P / sample);
        sample = (k->P + stride - 1) / stride;
        // main pass: dense filter only (mode 2); the tail-place pairs are the postings kernel's
        const int main_mode = 2;
        TileAux aux;
        // Optional ("post_first"): postings FIRST.  The pairs that share a tail place are
        // usually a target's strongest candidates, so the K-th best of its postings lists is a much better
        // start threshold for the dense filter than the bootstrap alone (fewer survivors, fewer drains).
        const bool post_first = use_ws && !k->opt_tc_seed && k->opt_post_first && !k->opt_debug_skip_postings;
        // the postings kernel (below, as a lambda: it runs before the dense kernel on the default path)
        auto run_postings = [&]() -> int {
            // postings kernel: one warp per work item (a target or a part of a heavy one), heaviest first
            VREC_TRY(k->d_post_bin.ensure((size_t)tn * 2));
            VREC_TRY(k->d_post_order.ensure((size_t)tn * POST_MAX_PARTS));
            VREC_TRY(k->d_post_bins.ensure(80));
            VREC_CUDA(cudaMemsetAsync(k->d_post_bins.p, 0, sizeof(int) * 80, ctx->stream));
            knn_post_work_kernel<<<(tn + 255) / 256, 256, 0, ctx->stream>>>(k->dev(), aux, k->d_tidx.p, tn, post_parts,
                                                                            k->d_post_bin.p, k->d_post_bin.p + tn, k->d_post_bins.p);
            VREC_LAUNCHED(ctx);
            knn_post_order_kernel<<<1, 32, 0, ctx->stream>>>(k->d_post_bins.p, k->d_post_bins.p + 64);
            VREC_LAUNCHED(ctx);
            knn_post_scatter_kernel<<<(tn + 255) / 256, 256, 0, ctx->stream>>>(tn, k->d_post_bin.p, k->d_post_bin.p + tn,
                                                                               k->d_post_bins.p, k->d_post_bins.p + 32, k->d_post_order.p);
            VREC_LAUNCHED(ctx);
            for (int p_ = 0; p_ < post_parts; ++p_) {          // slots of parts that do not exist stay empty
                knn_fill_int_stride_kernel<<<(tn + 255) / 256, 256, 0, ctx->stream>>>(k->d_part_cnt.p, tn, SP, S + p_, 0);
                VREC_LAUNCHED(ctx);
            }
            VREC_CUDA(cudaMemsetAsync(k->d_work.p, 0, sizeof(int), ctx->stream));
            const size_t psmem = (size_t)POST_WARPS * post_warp_bytes(K);
            // 8 warps x (12 K + 1728) bytes: 100 KB covers K <= 922; K up to TOPK_MAX_K = 1024 needs 110 KB
            // (the launch used to fail with "invalid argument" for 922 < K <= 1024)
            size_t &attr_post = ctx->attr_knn_post;
            if (psmem > attr_post) {
                const size_t want = std::max(psmem, (size_t)100 * 1024);
                VREC_CUDA(cudaFuncSetAttribute(knn_postings_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
                attr_post = want;
            }
            int pblocks = std::max(1, std::min(ctx->sm_count * 4, (tn + POST_WARPS - 1) / POST_WARPS));
            knn_postings_kernel<<<pblocks, POST_WARPS * 32, psmem, ctx->stream>>>(
                k->dev(), aux, k->d_tidx.p, k->opt_debug_skip_postings ? 0 : tn, K, S, SP, k->cat_dim, pw, cw, k->d_part.p,
                k->d_part_cnt.p, k->d_seed_thr.p, k->d_work.p, k->opt_debug_skip_postings ? nullptr : k->d_post_order.p,
                k->d_post_bins.p + 64);
            VREC_LAUNCHED(ctx);
            return VREC_OK;
        };
        if (use_tc) {
#if VREC_WITH_TC_BASELINE
            bool &attr_tc = ctx->attr_knn_tc;
            if (!attr_tc) {
                VREC_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                attr_tc = true;
            }
#endif
            aux = TileAux{nullptr, 0, k->d_head_slot_tc.p, k->d_pcp.p, k->d_pper.p, k->d_meta.p, k->d_rec.p, REC_TAIL_TC};
            aux.vals_f32 = k->rec_f32;
            k->attach_compact(aux);
            // No seed pass here: it cost more (heap warm-up on a sample, ~9 ms per 19K targets) than the
            // ~250 extra survivors per target it saves the main pass (~2 ms).  Optional via "tc_seed".
#if VREC_WITH_TC_BASELINE
            if (k->opt_tc_seed) {
                long long smp = std::min<long long>(k->P, std::max<long long>(TC_N, k->opt_tc_seed));
                long long str = std::max<long long>(1, k->P / smp);
                smp = (k->P + str - 1) / str;
                knn_tc_kernel<<<dim3(tiles, 1), TC_THREADS, smem, ctx->stream>>>(
                    k->dev(), aux, k->d_feat16.p, k->d_tidx.p, tn, K, 1, k->cat_dim, pw, cw, k->d_part.p,
                    k->d_part_cnt.p, str, smp, 1, k->d_seed_thr.p, SP);
                VREC_LAUNCHED(ctx);
            } else
#endif
            {
                VREC_CUDA(cudaMemsetAsync(k->d_seed_thr.p, 0, sizeof(double) * (size_t)tn, ctx->stream));
            }
            if (post_first) {
                for (int sp = 0; sp < S; ++sp) {               // the dense lists do not exist yet
                    knn_fill_int_stride_kernel<<<(tn + 255) / 256, 256, 0, ctx->stream>>>(k->d_part_cnt.p, tn, SP, sp, 0);
                    VREC_LAUNCHED(ctx);
                }
                VREC_TRY(run_postings());
                knn_seed_from_post_kernel<<<(int)(((long long)tn * 32 + 255) / 256), 256, 0, ctx->stream>>>(
                    tn, K, S, post_parts, SP, k->d_part.p, k->d_part_cnt.p, k->d_seed_thr.p);
                VREC_LAUNCHED(ctx);
            }
            if (use_ws) {
                VREC_TRY(k->d_tdense.ensure((size_t)tn * (size_t)std::max(1, (int)k->cat_dim)));
                aux.tdense = k->d_tdense.p;
                aux.cat_dim = (int)k->cat_dim;
                bool &attr_ws = ctx->attr_knn_ws;
                if (!attr_ws) {
                    VREC_CUDA(cudaFuncSetAttribute(knn_tc_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                    attr_ws = true;
                }
                if (!k->ev_dense[0]) {
                    VREC_CUDA(cudaEventCreate(&k->ev_dense[0]));
                    VREC_CUDA(cudaEventCreate(&k->ev_dense[1]));
                }
                VREC_CUDA(cudaEventRecord(k->ev_dense[0], ctx->stream));
                knn_tc_ws_kernel<<<dim3(tiles, S), WS_THREADS, smem, ctx->stream>>>(
                    k->dev(), aux, (const __half *)k->d_featsw.p, k->d_tidx.p, tn, K, S, k->cat_dim, pw, cw, k->d_part.p,
                    k->d_part_cnt.p, k->d_seed_thr.p, SP, k->tc_tiles, ws_pool_ints);
                VREC_LAUNCHED(ctx);
                VREC_CUDA(cudaEventRecord(k->ev_dense[1], ctx->stream));
            } else {
#if VREC_WITH_TC_BASELINE
                knn_tc_kernel<<<dim3(tiles, S), TC_THREADS, smem, ctx->stream>>>(
                    k->dev(), aux, k->d_feat16.p, k->d_tidx.p, tn, K, S, k->cat_dim, pw, cw, k->d_part.p,
                    k->d_part_cnt.p, 1, k->P, main_mode, k->d_seed_thr.p, SP);
                VREC_LAUNCHED(ctx);
#endif
            }
        } else {
            bool &attr_set = ctx->attr_knn_tile;
            if (!attr_set) {
                VREC_CUDA(cudaFuncSetAttribute(knn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set = true;
            }
            aux = TileAux{k->d_feat.p, k->fstride, k->d_head_slot.p, k->d_pcp.p, k->d_pper.p, k->d_meta.p, k->d_rec.p,
                          REC_TAIL_TILE};
            aux.vals_f32 = k->rec_f32;
            k->attach_compact(aux);
            knn_tile_kernel<<<dim3(tiles, 1), TILE_THREADS, smem, ctx->stream>>>(
                k->dev(), aux, k->d_tidx.p, tn, T, K, 1, k->cat_dim, pw, cw, k->d_part.p, k->d_part_cnt.p, stride, sample,
                1, k->d_seed_thr.p, SP);
            VREC_LAUNCHED(ctx);
            knn_tile_kernel<<<dim3(tiles, S), TILE_THREADS, smem, ctx->stream>>>(
                k->dev(), aux, k->d_tidx.p, tn, T, K, S, k->cat_dim, pw, cw, k->d_part.p, k->d_part_cnt.p, 1, k->P,
                main_mode, k->d_seed_thr.p, SP);
            VREC_LAUNCHED(ctx);
        }
        if (!post_first) VREC_TRY(run_postings());
    } else {
        dim3 grid(tn, S);
        knn_topk_kernel<<<grid, TOPK_THREADS, 0, ctx->stream>>>(k->dev(), k->d_tidx.p, K, S, pw, cw, k->d_part.p,
                                                               k->d_part_cnt.p);
        VREC_LAUNCHED(ctx);
    }
    if (SP * K <= MERGE_WARP_MAX && K <= 64) {
        knn_merge_warp_kernel<<<(int)(((long long)tn * 32 + 127) / 128), 128, 0, ctx->stream>>>(
            k->d_part.p, k->d_part_cnt.p, tn, K, SP, k->d_nb_rank.p, k->d_nb_idx.p, k->d_nb_cnt.p);
    } else {
        knn_merge_kernel<<<tn, MERGE_THREADS, 0, ctx->stream>>>(k->d_part.p, k->d_part_cnt.p, K, SP, k->d_nb_rank.p,
                                                               k->d_nb_idx.p, k->d_nb_cnt.p);
    }
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}

// dense masked similarity rows of a tile of targets (any K)
int knn_run_dense(vrec_knn *k, int tn, double pw, double cw, int K) {
    vrec_ctx *ctx = k->ctx;
    VREC_TRY(k->d_sim.ensure((size_t)tn * (size_t)std::max<int64_t>(1, k->P)));
    if (k->P == 0) return VREC_OK;
    dim3 grid((unsigned)((k->P + 255) / 256), tn);
    knn_sim_rows_kernel<<<grid, 256, 0, ctx->stream>>>(k->dev(), k->d_tidx.p, pw, cw, k->d_sim.p);
    VREC_LAUNCHED(ctx);
    knn_select_mask_kernel<<<tn, 1024, 0, ctx->stream>>>(k->d_sim.p, k->P, K);
    VREC_LAUNCHED(ctx);
    return VREC_OK;
}

int knn_lookup(vrec_knn *k, const long long *d_targets, int tn, int *d_status) {
    VREC_TRY(k->d_tidx.ensure((size_t)tn));
    knn_lookup_kernel<<<(tn + 127) / 128, 128, 0, k->ctx->stream>>>(k->dev(), d_targets, tn, k->d_tidx.p, d_status);
    VREC_LAUNCHED(k->ctx);
    return VREC_OK;
}

int knn_ensure_rate_scratch(vrec_knn *k, int K) {
    int touch_cap = (int)std::min<int64_t>((int64_t)K * std::max(1, k->max_rating_row), k->rdim);
    touch_cap = std::max(touch_cap, 1);
    if (k->rate_slots > 0 && touch_cap <= k->rate_touch_cap) return VREC_OK;
    int64_t per_slot = (int64_t)k->rdim * 16 + (int64_t)touch_cap * 4;
    int64_t slots = std::min<int64_t>(k->ctx->sm_count * 16, ((int64_t)4 << 30) / std::max<int64_t>(1, per_slot));
    slots = std::max<int64_t>(RATE_WARPS, slots / RATE_WARPS * RATE_WARPS);
    VREC_TRY(k->d_num.alloc((size_t)slots * k->rdim));
    VREC_TRY(k->d_den.alloc((size_t)slots * k->rdim));
    VREC_TRY(k->d_touched.alloc((size_t)slots * touch_cap));
    VREC_CUDA(cudaMemsetAsync(k->d_num.p, 0, k->d_num.bytes(), k->ctx->stream));
    VREC_CUDA(cudaMemsetAsync(k->d_den.p, 0, k->d_den.bytes(), k->ctx->stream));
    k->rate_slots = (int)slots;
    k->rate_touch_cap = touch_cap;
    return VREC_OK;
}

}  // namespace

extern "C" int vrec_knn_query_device(vrec_knn *k, const int64_t *d_targets, int32_t n_targets, double pw,
                                     double cw, int32_t K, int32_t max_recs, int64_t *d_out_place,
                                     double *d_out_rating, int32_t *d_out_count, int32_t *d_out_status) {
    if (!k || n_targets < 0 || (n_targets > 0 && (!d_targets || !d_out_count || !d_out_status))) {
        vrec_set_error("vrec_knn_query: NULL argument");
        return VREC_EINVAL;
    }
    VREC_TRY(knn_check_params(pw, cw, K));
    if (max_recs < 0) {
        vrec_set_error("Maximum recommendations number must be non-negative");
        return VREC_EINVAL;
    }
    if (n_targets == 0) return VREC_OK;
    vrec_ctx *ctx = k->ctx;
    VREC_CUDA(cudaSetDevice(ctx->device));
    const bool gather = use_gather_path(k, K);
    int64_t tile = k->opt_tile;
    // gather path: as many targets per pass as possible (more M-tiles -> no candidate splits, one wave)
    if (tile <= 0) tile = gather ? 65536 : std::max<int64_t>(1, ((int64_t)1 << 30) / (8 * std::max<int64_t>(1, k->P)));
    tile = std::min<int64_t>(tile, n_targets);
    const unsigned char *flag = k->has_filter ? k->d_flag.p : nullptr;
    if (gather) VREC_TRY(knn_ensure_rate_scratch(k, K));
    if (!gather) VREC_TRY(k->d_est.ensure((size_t)k->rdim));
    for (int64_t t0 = 0; t0 < n_targets; t0 += tile) {
        const int tn = (int)std::min<int64_t>(tile, n_targets - t0);
        VREC_TRY(knn_lookup(k, (const long long *)d_targets + t0, tn, d_out_status + t0));
        if (max_recs == 0) {
            knn_fill_int_kernel<<<(tn + 255) / 256, 256, 0, ctx->stream>>>(d_out_count + t0, tn, 0);
            VREC_LAUNCHED(ctx);
            continue;
        }
        if (gather) {
            VREC_TRY(knn_run_topk(k, tn, pw, cw, K));
            int blocks = std::min(k->rate_slots / RATE_WARPS, (tn + RATE_WARPS - 1) / RATE_WARPS);
            knn_rate_gather_kernel<<<blocks, RATE_WARPS * 32, 0, ctx->stream>>>(
                k->d_rrp.p, k->d_rpl.p, k->d_rv.p, k->d_nb_idx.p, k->d_nb_cnt.p, tn, K, k->rdim, flag, k->d_num.p,
                k->d_den.p, k->d_touched.p, k->rate_touch_cap, max_recs, (long long *)d_out_place + t0 * max_recs,
                d_out_rating + t0 * max_recs, d_out_count + t0);
            VREC_LAUNCHED(ctx);
        } else {
            VREC_TRY(knn_run_dense(k, tn, pw, cw, K));
            for (int tt = 0; tt < tn; ++tt) {
                knn_rate_cols_kernel<<<(int)(((long long)k->rdim * 32 + 127) / 128), 128, 0, ctx->stream>>>(
                    k->d_ccp.p, k->d_cper.p, k->d_crv.p, k->rdim, k->d_sim.p + (size_t)tt * k->P, k->d_est.p);
                VREC_LAUNCHED(ctx);
                VREC_TRY(vrec_launch_select_topn(ctx, k->d_est.p, nullptr, flag, k->rdim, 0, 1, max_recs,
                                                 (long long *)d_out_place + (t0 + tt) * max_recs,
                                                 d_out_rating + (t0 + tt) * max_recs, d_out_count + t0 + tt));
            }
        }
    }
    return VREC_OK;
}

extern "C" int vrec_knn_query(vrec_knn *k, const int64_t *targets, int32_t n_targets, double pw, double cw,
                              int32_t K, const int64_t *place_filter, int64_t n_filter, int32_t max_recs,
                              int64_t *out_place_id, double *out_rating, int32_t *out_count, int32_t *out_status) {
    if (!k || n_targets < 0 || (n_targets > 0 && (!targets || !out_count || !out_status))) {
        vrec_set_error("vrec_knn_query: NULL argument");
        return VREC_EINVAL;
    }
    VREC_TRY(knn_check_params(pw, cw, K));
    if (max_recs < 0 || n_filter < 0) {
        vrec_set_error("Maximum recommendations number must be non-negative");
        return VREC_EINVAL;
    }
    if (n_targets == 0) return VREC_OK;
    vrec_ctx *ctx = k->ctx;
    VREC_CUDA(cudaSetDevice(ctx->device));
    VREC_TRY(vrec_knn_set_filter(k, place_filter, n_filter));
    const size_t m = (size_t)std::max(1, (int)max_recs);
    VREC_TRY(k->d_targets.ensure((size_t)n_targets));
    VREC_TRY(k->d_out_place.ensure((size_t)n_targets * m));
    VREC_TRY(k->d_out_rating.ensure((size_t)n_targets * m));
    VREC_TRY(k->d_out_count.ensure((size_t)n_targets));
    VREC_TRY(k->d_status.ensure((size_t)n_targets));
    VREC_CUDA(cudaMemcpyAsync(k->d_targets.p, targets, sizeof(int64_t) * (size_t)n_targets, cudaMemcpyHostToDevice,
                              ctx->stream));
    VREC_TRY(vrec_knn_query_device(k, (const int64_t *)k->d_targets.p, n_targets, pw, cw, K, max_recs,
                                   (int64_t *)k->d_out_place.p, k->d_out_rating.p, k->d_out_count.p,
                                   k->d_status.p));
    if (max_recs > 0 && out_place_id && out_rating) {
        VREC_CUDA(cudaMemcpyAsync(out_place_id, k->d_out_place.p, sizeof(int64_t) * (size_t)n_targets * max_recs,
                                  cudaMemcpyDeviceToHost, ctx->stream));
        VREC_CUDA(cudaMemcpyAsync(out_rating, k->d_out_rating.p, sizeof(double) * (size_t)n_targets * max_recs,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    }
    VREC_CUDA(cudaMemcpyAsync(out_count, k->d_out_count.p, sizeof(int32_t) * (size_t)n_targets,
                              cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaMemcpyAsync(out_status, k->d_status.p, sizeof(int32_t) * (size_t)n_targets,
                              cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int32_t t = 0; t < n_targets; ++t)
        if (out_status[t] == VREC_ENOENT) {
            vrec_set_error("No such person: %lld", (long long)targets[t]);     // knn/KnnRecommender.scala:83
            break;
        }
    return VREC_OK;
}

namespace {

// one target -> d_tidx / status on the host
int knn_single_target(vrec_knn *k, int64_t target, int *status) {
    vrec_ctx *ctx = k->ctx;
    VREC_TRY(k->d_targets.ensure(1));
    VREC_TRY(k->d_status.ensure(1));
    VREC_CUDA(cudaMemcpyAsync(k->d_targets.p, &target, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    VREC_TRY(knn_lookup(k, k->d_targets.p, 1, k->d_status.p));
    VREC_CUDA(cudaMemcpyAsync(status, k->d_status.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*status == VREC_ENOENT) vrec_set_error("No such person: %lld", (long long)target);
    return VREC_OK;
}

}  // namespace

extern "C" int vrec_knn_neighbours(vrec_knn *k, int64_t target, double pw, double cw, int32_t K,
                                   int64_t *out_person_id, double *out_similarity, int32_t capacity,
                                   int32_t *out_count) {
    if (!k || !out_person_id || !out_similarity || !out_count || capacity < 0) return VREC_EINVAL;
    VREC_TRY(knn_check_params(pw, cw, K));
    if (K > TOPK_MAX_K) {
        vrec_set_error("vrec_knn_neighbours supports k_nearest <= %d; use vrec_knn_similarities", TOPK_MAX_K);
        return VREC_EINVAL;
    }
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    int status = 0;
    VREC_TRY(knn_single_target(k, target, &status));
    if (status != VREC_OK) return status;
    VREC_TRY(knn_run_topk(k, 1, pw, cw, K));
    int cnt = 0;
    std::vector<Nb> h((size_t)K);
    VREC_CUDA(cudaMemcpyAsync(&cnt, k->d_nb_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, k->ctx->stream));
    VREC_CUDA(cudaMemcpyAsync(h.data(), k->d_nb_rank.p, sizeof(Nb) * (size_t)K, cudaMemcpyDeviceToHost,
                              k->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    cnt = std::min(cnt, (int)capacity);
    for (int i = 0; i < cnt; ++i) {
        out_person_id[i] = k->h_person[(size_t)h[i].idx];
        out_similarity[i] = h[i].sim;
    }
    *out_count = cnt;
    return VREC_OK;
}

// Dense similarity vector of one target after orderBy(similarity desc).limit(K): out_sim[P] in
// ascending person_id order (vrec_knn_person_ids), 0.0 for persons that are not neighbours.
extern "C" int vrec_knn_similarities(vrec_knn *k, int64_t target, double pw, double cw, int32_t K,
                                     double *out_sim) {
    if (!k || !out_sim) return VREC_EINVAL;
    VREC_TRY(knn_check_params(pw, cw, K));
    VREC_CUDA(cudaSetDevice(k->ctx->device));
    int status = 0;
    VREC_TRY(knn_single_target(k, target, &status));
    if (status != VREC_OK) return status;
    VREC_TRY(knn_run_dense(k, 1, pw, cw, K));
    VREC_CUDA(cudaMemcpyAsync(out_sim, k->d_sim.p, sizeof(double) * (size_t)k->P, cudaMemcpyDeviceToHost,
                              k->ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(k->ctx->stream));
    return VREC_OK;
}

extern "C" int vrec_knn_estimates(vrec_knn *k, int64_t target, double pw, double cw, int32_t K,
                                  int64_t *out_place_id, double *out_rating, int64_t capacity,
                                  int64_t *out_count) {
    if (!k || !out_place_id || !out_rating || !out_count || capacity < 0) return VREC_EINVAL;
    VREC_TRY(knn_check_params(pw, cw, K));
    vrec_ctx *ctx = k->ctx;
    VREC_CUDA(cudaSetDevice(ctx->device));
    int status = 0;
    VREC_TRY(knn_single_target(k, target, &status));
    if (status != VREC_OK) return status;
    VREC_TRY(knn_run_dense(k, 1, pw, cw, K));
    VREC_TRY(k->d_est.ensure((size_t)k->rdim));
    knn_rate_cols_kernel<<<(int)(((long long)k->rdim * 32 + 127) / 128), 128, 0, ctx->stream>>>(k->d_ccp.p, k->d_cper.p, k->d_crv.p,
                                                                        k->rdim, k->d_sim.p, k->d_est.p);
    VREC_LAUNCHED(ctx);
    std::vector<double> est((size_t)k->rdim);
    VREC_CUDA(cudaMemcpyAsync(est.data(), k->d_est.p, sizeof(double) * (size_t)k->rdim, cudaMemcpyDeviceToHost,
                              ctx->stream));
    VREC_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t n = 0;
    for (int pl = 0; pl < k->rdim; ++pl) {
        if (est[pl] != est[pl]) continue;
        if (n < capacity) {
            out_place_id[n] = pl;
            out_rating[n] = est[pl];
        }
        ++n;
    }
    *out_count = std::min(n, capacity);
    return VREC_OK;
}
