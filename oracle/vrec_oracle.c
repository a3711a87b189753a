/*
 * oracle/vrec_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, fp64) of the two hot paths of
 * tashoyan/locations-recommender.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (libvrec.so) never links, imports or executes it.
 *
 * The reference is Scala on Spark; no JVM exists in this image, so the
 * reference itself cannot be run here ("oracle/_ref" is not buildable).  The
 * oracle is pinned instead against the reference's own known-answer tests:
 *   recommender/src/test/scala/.../knn/DistanceTest.scala:10-60        (8 KATs)
 *   recommender/src/test/scala/.../stochastic/StochasticRecommenderTest.scala:11-94 (3 KATs)
 * (see tests/test_oracle_kat.py).  KnnRecommender's combine / top-K / rating
 * reduction has NO test in the reference: that part is "parity unpinned"
 * (restated from the source only).
 *
 * Paths below are relative to
 * /root/reference/recommender/src/main/scala/com/github/tashoyan/recommender/.
 *
 * Third-party arithmetic restated here (not vendored in the reference):
 *   org.apache.spark:spark-mllib-local_2.12:3.1.2  BLAS.dot(SparseVector,SparseVector)
 *   -- two-pointer merge over ascending indices, `sum += x(kx) * y(ky)`, fp64,
 *   called from knn/Distance.scala:8.
 *
 * Determinism rules added on top of Spark (which leaves them open):
 *   - ties: neighbours by (similarity desc, person_id asc); places by
 *     (value desc, id asc)   [north_star: "ties broken by ID"]
 *   - KNN rating sums: neighbours visited in ascending person_id, both sums
 *     accumulated in the same pass, left to right.
 *   - SG sigma sums: in-edges of a vertex in ascending source id (stable for
 *     duplicates), summed in the CANONICAL ORDER `canon_sum` below -- a fixed
 *     32-lane strided + xor-butterfly order (what a warp does), segmented
 *     every 1024 terms.  For <= 3 terms it is identical to left-to-right
 *     summation, which is what the reference KATs pin.  In graphs with more
 *     than 3 * 2^21 vertices the in-edges of a vertex are first grouped by source
 *     block (source index / (3 * 2^21)); every group is summed in the canonical
 *     order and the group sums are added left to right.
 *   - SG residual: sum of squared differences left to right over ascending
 *     vertex id.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VRO_OK 0
#define VRO_ENOENT (-2)
#define VRO_EINVAL (-22)
#define VRO_ENOMEM (-12)

#define CANON_LANES 32
#define CANON_SEG 1024
#define CANON_SRC_BLOCK 6291456      /* graphs with more vertices than this: sources grouped in blocks of 3 * 2^21 */

/* ------------------------------------------------------------------ */
/* knn/Distance.scala                                                  */
/* ------------------------------------------------------------------ */

/* knn/Distance.scala:11-16  vectorLength: values.map(v => v*v).sum, math.sqrt */
double vro_vector_length(const double *values, int32_t n)
{
    double s = 0.0;
    for (int32_t i = 0; i < n; ++i) {
        double sq = values[i] * values[i];
        s = s + sq;
    }
    return sqrt(s);
}

/* spark-mllib-local 3.1.2 BLAS.dot(sparse, sparse): "y catching x" merge. */
double vro_sparse_dot(const int32_t *xi, const double *xv, int32_t nx,
                      const int32_t *yi, const double *yv, int32_t ny)
{
    int32_t kx = 0, ky = 0;
    double sum = 0.0;
    while (kx < nx && ky < ny) {
        int32_t ix = xi[kx];
        while (ky < ny && yi[ky] < ix) ky++;
        if (ky < ny && yi[ky] == ix) {
            double p = xv[kx] * yv[ky];
            sum = sum + p;
            ky++;
        }
        kx++;
    }
    return sum;
}

/* knn/Distance.scala:7-9  (v1 dot v2) / (vectorLength(v1) * vectorLength(v2)) */
double vro_cosine(const int32_t *xi, const double *xv, int32_t nx,
                  const int32_t *yi, const double *yv, int32_t ny)
{
    double d = vro_sparse_dot(xi, xv, nx, yi, yv, ny);
    double l1 = vro_vector_length(xv, nx);
    double l2 = vro_vector_length(yv, ny);
    double den = l1 * l2;
    return d / den;
}

/* ------------------------------------------------------------------ */
/* knn/KnnRecommender.scala                                            */
/* ------------------------------------------------------------------ */

typedef struct {
    int64_t P;
    const int64_t *person_id;   /* ascending */
    const int64_t *prp; const int32_t *pci; const double *pv;   /* place vectors, CSR    */
    const int64_t *crp; const int32_t *cci; const double *cv;   /* category vectors, CSR */
    const int64_t *rrp; const int64_t *rpl; const int64_t *rv;  /* place_ratings rows grouped by person */
} vro_knn_data;

static int64_t find_person(const vro_knn_data *d, int64_t id)
{
    int64_t lo = 0, hi = d->P;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (d->person_id[mid] < id) lo = mid + 1; else hi = mid;
    }
    return (lo < d->P && d->person_id[lo] == id) ? lo : -1;
}

/* knn/KnnRecommender.scala:76-96 (findSimilarPersons0, both kinds) +
 * :38-45 (outer join, na.fill(0), weighted combine).  A person absent from a
 * table is an empty CSR row.  out_sim[i] = combined similarity, or 0.0 when
 * the person is the target or is in neither filtered set.                   */
static int knn_similarities(const vro_knn_data *d, double pw, double cw,
                            int64_t target_id, double *out_sim)
{
    int64_t t = find_person(d, target_id);
    if (t < 0) return VRO_ENOENT;
    int64_t tps = d->prp[t], tpn = d->prp[t + 1] - tps;
    int64_t tcs = d->crp[t], tcn = d->crp[t + 1] - tcs;
    /* :77-83 "No such person" when the target has no row in either table */
    if (tpn == 0 || tcn == 0) return VRO_ENOENT;
    for (int64_t i = 0; i < d->P; ++i) {
        double ps = 0.0, cs = 0.0;
        int keep = 0;
        if (i != t) {
            int64_t s = d->prp[i], n = d->prp[i + 1] - s;
            if (n > 0) {
                /* cosineSimilarity(vector = row, personRatingVector = target) :86-88 */
                double v = vro_cosine(d->pci + s, d->pv + s, (int32_t)n,
                                      d->pci + tps, d->pv + tps, (int32_t)tpn);
                if (v > 0) { ps = v; keep = 1; }       /* :91 where(sim > 0) */
            }
            s = d->crp[i]; n = d->crp[i + 1] - s;
            if (n > 0) {
                double v = vro_cosine(d->cci + s, d->cv + s, (int32_t)n,
                                      d->cci + tcs, d->cv + tcs, (int32_t)tcn);
                if (v > 0) { cs = v; keep = 1; }
            }
        }
        if (keep) {
            double a = ps * pw;          /* :42-44, no FMA on the JVM */
            double b = cs * cw;
            out_sim[i] = a + b;
        } else {
            out_sim[i] = 0.0;
        }
    }
    return VRO_OK;
}

typedef struct { double sim; int64_t idx; } nb_t;

/* order: similarity desc, person index (== person_id order) asc */
static int nb_before(const nb_t *a, const nb_t *b)
{
    if (a->sim != b->sim) return a->sim > b->sim;
    return a->idx < b->idx;
}
static int nb_cmp_qsort(const void *pa, const void *pb)
{
    const nb_t *a = (const nb_t *)pa, *b = (const nb_t *)pb;
    if (nb_before(a, b)) return -1;
    if (nb_before(b, a)) return 1;
    return 0;
}
static int idx_cmp_qsort(const void *pa, const void *pb)
{
    const nb_t *a = (const nb_t *)pa, *b = (const nb_t *)pb;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/* binary heap whose root is the WORST kept neighbour */
static void heap_sift_down(nb_t *h, int64_t n, int64_t i)
{
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && nb_before(&h[w], &h[l])) w = l;
        if (r < n && nb_before(&h[w], &h[r])) w = r;
        if (w == i) return;
        nb_t tmp = h[i]; h[i] = h[w]; h[w] = tmp;
        i = w;
    }
}

/* knn/KnnRecommender.scala:46-48  orderBy(similarity desc).limit(kNearest).
 * Returns the kept neighbours sorted (sim desc, idx asc). */
static int knn_top_k(const double *sim, int64_t P, int64_t K, nb_t **out, int64_t *count)
{
    int64_t cap = K < P ? K : P;
    nb_t *h = (nb_t *)malloc(sizeof(nb_t) * (size_t)(cap > 0 ? cap : 1));
    if (!h) return VRO_ENOMEM;
    int64_t n = 0;
    for (int64_t i = 0; i < P; ++i) {
        if (!(sim[i] > 0)) continue;
        nb_t e = { sim[i], i };
        if (n < cap) {
            h[n++] = e;
            if (n == cap) for (int64_t j = n / 2 - 1; j >= 0; --j) heap_sift_down(h, n, j);
        } else if (nb_before(&e, &h[0])) {
            h[0] = e;
            heap_sift_down(h, n, 0);
        }
    }
    qsort(h, (size_t)n, sizeof(nb_t), nb_cmp_qsort);
    *out = h; *count = n;
    return VRO_OK;
}

static int knn_check_params(double pw, double cw, int64_t K)
{
    /* knn/KnnRecommender.scala:17-20 */
    if (!(pw > 0 && pw < 1.0)) return VRO_EINVAL;
    if (!(cw > 0 && cw < 1.0)) return VRO_EINVAL;
    if (!(pw + cw == 1.0)) return VRO_EINVAL;
    if (K <= 0) return VRO_EINVAL;
    return VRO_OK;
}

int vro_knn_similarities(int64_t P, const int64_t *person_id,
                         const int64_t *prp, const int32_t *pci, const double *pv,
                         const int64_t *crp, const int32_t *cci, const double *cv,
                         double pw, double cw, int64_t target_id, double *out_sim)
{
    vro_knn_data d = { P, person_id, prp, pci, pv, crp, cci, cv, 0, 0, 0 };
    return knn_similarities(&d, pw, cw, target_id, out_sim);
}

/* findSimilarPersons, knn/KnnRecommender.scala:27-49 */
int vro_knn_neighbours(int64_t P, const int64_t *person_id,
                       const int64_t *prp, const int32_t *pci, const double *pv,
                       const int64_t *crp, const int32_t *cci, const double *cv,
                       double pw, double cw, int32_t K, int64_t target_id,
                       int64_t *out_person_id, double *out_sim, int32_t *out_count)
{
    int rc = knn_check_params(pw, cw, K);
    if (rc) return rc;
    vro_knn_data d = { P, person_id, prp, pci, pv, crp, cci, cv, 0, 0, 0 };
    double *sim = (double *)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
    if (!sim) return VRO_ENOMEM;
    rc = knn_similarities(&d, pw, cw, target_id, sim);
    if (rc) { free(sim); return rc; }
    nb_t *nb; int64_t n;
    rc = knn_top_k(sim, P, K, &nb, &n);
    free(sim);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) { out_person_id[i] = person_id[nb[i].idx]; out_sim[i] = nb[i].sim; }
    *out_count = (int32_t)n;
    free(nb);
    return VRO_OK;
}

typedef struct { double est; int64_t place; } rec_t;
static int rec_cmp_rank(const void *pa, const void *pb)
{
    const rec_t *a = (const rec_t *)pa, *b = (const rec_t *)pb;
    if (a->est != b->est) return a->est > b->est ? -1 : 1;
    return (a->place > b->place) - (a->place < b->place);
}
static int rec_cmp_place(const void *pa, const void *pb)
{
    const rec_t *a = (const rec_t *)pa, *b = (const rec_t *)pb;
    return (a->place > b->place) - (a->place < b->place);
}

/* scratch for one query thread */
typedef struct {
    double *sim, *num, *den;
    int64_t *touched;
    unsigned char *seen;
    const unsigned char *flag;    /* place filter over [0, place_dim) or NULL */
    int64_t place_dim;
} knn_scratch;

static int scratch_alloc(knn_scratch *s, int64_t P, int64_t place_dim)
{
    memset(s, 0, sizeof(*s));
    s->place_dim = place_dim;
    s->sim = (double *)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
    s->num = (double *)calloc((size_t)(place_dim > 0 ? place_dim : 1), sizeof(double));
    s->den = (double *)calloc((size_t)(place_dim > 0 ? place_dim : 1), sizeof(double));
    s->touched = (int64_t *)malloc(sizeof(int64_t) * (size_t)(place_dim > 0 ? place_dim : 1));
    s->seen = (unsigned char *)calloc((size_t)(place_dim > 0 ? place_dim : 1), 1);
    return (s->sim && s->num && s->den && s->touched && s->seen) ? VRO_OK : VRO_ENOMEM;
}
static void scratch_free(knn_scratch *s)
{
    free(s->sim); free(s->num); free(s->den); free(s->touched); free(s->seen);
}

/* One full query = KnnRecommender.makeRecommendations (knn/KnnRecommender.scala:22-25,51-70)
 * followed, in mode 0, by KnnRecommenderMain.printRecommendations' region
 * filter + orderBy(estimated_rating desc).limit(N) (knn/KnnRecommenderMain.scala:96-102).
 * mode 1 returns the raw makeRecommendations rows (all places, sorted by place id). */
static int knn_query_one(const vro_knn_data *d, double pw, double cw, int64_t K,
                         int64_t target_id, knn_scratch *s, int mode, int64_t max_recs,
                         int64_t *out_place, double *out_rating, int64_t *out_count)
{
    *out_count = 0;
    int rc = knn_similarities(d, pw, cw, target_id, s->sim);
    if (rc) return rc;
    nb_t *nb; int64_t n;
    rc = knn_top_k(s->sim, d->P, K, &nb, &n);
    if (rc) return rc;
    /* join placeRatings(person != target) with similarPersons, groupBy(place_id):
     * neighbours in ascending person_id; num and den in the same pass */
    qsort(nb, (size_t)n, sizeof(nb_t), idx_cmp_qsort);
    int64_t nt = 0;
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = nb[k].idx;
        double sim = nb[k].sim;
        for (int64_t e = d->rrp[i]; e < d->rrp[i + 1]; ++e) {
            int64_t pl = d->rpl[e];
            if (pl < 0 || pl >= s->place_dim) continue;
            double w = (double)d->rv[e] * sim;          /* :60 rating * similarity */
            if (!s->seen[pl]) { s->seen[pl] = 1; s->touched[nt++] = pl; }
            s->num[pl] = s->num[pl] + w;                /* :63 */
            s->den[pl] = s->den[pl] + sim;              /* :64 */
        }
    }
    free(nb);
    rec_t *recs = (rec_t *)malloc(sizeof(rec_t) * (size_t)(nt > 0 ? nt : 1));
    if (!recs) return VRO_ENOMEM;
    int64_t nr = 0;
    for (int64_t k = 0; k < nt; ++k) {
        int64_t pl = s->touched[k];
        if (mode == 1 || !s->flag || s->flag[pl]) {
            recs[nr].est = s->num[pl] / s->den[pl];     /* :68 */
            recs[nr].place = pl;
            nr++;
        }
        s->num[pl] = 0.0; s->den[pl] = 0.0; s->seen[pl] = 0;
    }
    if (mode == 1) {
        qsort(recs, (size_t)nr, sizeof(rec_t), rec_cmp_place);
    } else {
        qsort(recs, (size_t)nr, sizeof(rec_t), rec_cmp_rank);
        if (nr > max_recs) nr = max_recs;
    }
    for (int64_t k = 0; k < nr; ++k) { out_place[k] = recs[k].place; out_rating[k] = recs[k].est; }
    *out_count = nr;
    free(recs);
    return VRO_OK;
}

static unsigned char *build_flag(const int64_t *filter, int64_t n_filter, int64_t place_dim)
{
    unsigned char *f = (unsigned char *)calloc((size_t)(place_dim > 0 ? place_dim : 1), 1);
    if (!f) return 0;
    for (int64_t i = 0; i < n_filter; ++i)
        if (filter[i] >= 0 && filter[i] < place_dim) f[filter[i]] = 1;
    return f;
}

/* Batch of queries; parallel over targets with OpenMP (the CPU baseline of
 * bench.py).  out_* are [n_targets x max_recs]; status per target. */
int vro_knn_query_batch(int64_t P, const int64_t *person_id,
                        const int64_t *prp, const int32_t *pci, const double *pv,
                        const int64_t *crp, const int32_t *cci, const double *cv,
                        const int64_t *rrp, const int64_t *rpl, const int64_t *rv,
                        int64_t place_dim,
                        const int64_t *targets, int64_t n_targets,
                        double pw, double cw, int32_t K,
                        const int64_t *filter, int64_t n_filter, int32_t max_recs,
                        int64_t *out_place, double *out_rating, int32_t *out_count,
                        int32_t *out_status, int32_t n_threads)
{
    int rc = knn_check_params(pw, cw, K);
    if (rc) return rc;
    if (max_recs < 0) return VRO_EINVAL;
    vro_knn_data d = { P, person_id, prp, pci, pv, crp, cci, cv, rrp, rpl, rv };
    unsigned char *flag = filter ? build_flag(filter, n_filter, place_dim) : 0;
    if (filter && !flag) return VRO_ENOMEM;
    int fail = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        knn_scratch s;
        int ok = scratch_alloc(&s, P, place_dim) == VRO_OK;
        s.flag = flag;
        int64_t *pl = (int64_t *)malloc(sizeof(int64_t) * (size_t)(place_dim > 0 ? place_dim : 1));
        double *rt = (double *)malloc(sizeof(double) * (size_t)(place_dim > 0 ? place_dim : 1));
        if (!ok || !pl || !rt) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t t = 0; t < n_targets; ++t) {
                int64_t cnt = 0;
                int st = knn_query_one(&d, pw, cw, K, targets[t], &s, 0, max_recs, pl, rt, &cnt);
                out_status[t] = st;
                out_count[t] = (int32_t)cnt;
                for (int64_t k = 0; k < cnt; ++k) {
                    out_place[t * max_recs + k] = pl[k];
                    out_rating[t * max_recs + k] = rt[k];
                }
            }
        }
        scratch_free(&s); free(pl); free(rt);
    }
    free(flag);
    return fail ? VRO_ENOMEM : VRO_OK;
}

/* Raw makeRecommendations rows for one target: every place rated by >= 1
 * neighbour, sorted by place id.  out arrays sized place_dim. */
int vro_knn_estimates(int64_t P, const int64_t *person_id,
                      const int64_t *prp, const int32_t *pci, const double *pv,
                      const int64_t *crp, const int32_t *cci, const double *cv,
                      const int64_t *rrp, const int64_t *rpl, const int64_t *rv,
                      int64_t place_dim, int64_t target, double pw, double cw, int32_t K,
                      int64_t *out_place, double *out_rating, int64_t *out_count)
{
    int rc = knn_check_params(pw, cw, K);
    if (rc) return rc;
    vro_knn_data d = { P, person_id, prp, pci, pv, crp, cci, cv, rrp, rpl, rv };
    knn_scratch s;
    if (scratch_alloc(&s, P, place_dim) != VRO_OK) { scratch_free(&s); return VRO_ENOMEM; }
    rc = knn_query_one(&d, pw, cw, K, target, &s, 1, 0, out_place, out_rating, out_count);
    scratch_free(&s);
    return rc;
}

/* ------------------------------------------------------------------ */
/* stochastic/StochasticRecommender.scala                              */
/* ------------------------------------------------------------------ */

/* Canonical summation order shared with the CUDA engine (see header). */
static double warp_sum(const double *t, int64_t n)
{
    double lane[CANON_LANES];
    for (int l = 0; l < CANON_LANES; ++l) lane[l] = 0.0;
    for (int64_t i = 0; i < n; ++i) lane[i % CANON_LANES] = lane[i % CANON_LANES] + t[i];
    for (int off = 1; off < CANON_LANES; off <<= 1) {
        double nx[CANON_LANES];
        for (int l = 0; l < CANON_LANES; ++l) nx[l] = lane[l] + lane[l ^ off];
        for (int l = 0; l < CANON_LANES; ++l) lane[l] = nx[l];
    }
    return lane[0];
}

typedef struct {
    int64_t N, nnz;
    int64_t *ids;       /* ascending vertex ids                       */
    int64_t *rowptr;    /* CSR of P^T: row = target vertex            */
    int32_t *src;       /* source vertex index, ascending in each row */
    double *w;
} vro_sg;

static int i64_cmp(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
static int64_t sg_find(const vro_sg *g, int64_t id)
{
    int64_t lo = 0, hi = g->N;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (g->ids[mid] < id) lo = mid + 1; else hi = mid;
    }
    return (lo < g->N && g->ids[lo] == id) ? lo : -1;
}

void vro_sg_free(vro_sg *g)
{
    if (!g) return;
    free(g->ids); free(g->rowptr); free(g->src); free(g->w);
    free(g);
}

/* stochastic/StochasticRecommender.scala:42-49: vertexes = distinct(source ∪ target);
 * the edge list becomes the CSR of P^T with in-edges in ascending source
 * order (stable: duplicate (s,t) edges keep file order).                      */
int vro_sg_build(int64_t nnz, const int64_t *source, const int64_t *target, const double *weight,
                 vro_sg **out)
{
    vro_sg *g = (vro_sg *)calloc(1, sizeof(vro_sg));
    if (!g) return VRO_ENOMEM;
    int64_t *all = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * nnz + 1));
    if (!all) { free(g); return VRO_ENOMEM; }
    for (int64_t e = 0; e < nnz; ++e) { all[2 * e] = source[e]; all[2 * e + 1] = target[e]; }
    qsort(all, (size_t)(2 * nnz), sizeof(int64_t), i64_cmp);
    int64_t N = 0;
    for (int64_t i = 0; i < 2 * nnz; ++i) if (i == 0 || all[i] != all[i - 1]) all[N++] = all[i];
    g->N = N; g->nnz = nnz; g->ids = all;
    g->rowptr = (int64_t *)calloc((size_t)(N + 1), sizeof(int64_t));
    g->src = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
    g->w = (double *)malloc(sizeof(double) * (size_t)(nnz + 1));
    int32_t *si = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
    int32_t *ti = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
    int64_t *cnt_s = (int64_t *)calloc((size_t)(N + 1), sizeof(int64_t));
    int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nnz + 1));
    if (!g->rowptr || !g->src || !g->w || !si || !ti || !cnt_s || !ord) {
        free(si); free(ti); free(cnt_s); free(ord); vro_sg_free(g); return VRO_ENOMEM;
    }
    for (int64_t e = 0; e < nnz; ++e) {
        si[e] = (int32_t)sg_find(g, source[e]);
        ti[e] = (int32_t)sg_find(g, target[e]);
    }
    /* stable counting sort by source, then stable counting sort by target */
    for (int64_t e = 0; e < nnz; ++e) cnt_s[si[e] + 1]++;
    for (int64_t i = 0; i < N; ++i) cnt_s[i + 1] += cnt_s[i];
    for (int64_t e = 0; e < nnz; ++e) ord[cnt_s[si[e]]++] = e;
    for (int64_t e = 0; e < nnz; ++e) g->rowptr[ti[e] + 1]++;
    for (int64_t i = 0; i < N; ++i) g->rowptr[i + 1] += g->rowptr[i];
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N + 1));
    if (!pos) { free(si); free(ti); free(cnt_s); free(ord); vro_sg_free(g); return VRO_ENOMEM; }
    memcpy(pos, g->rowptr, sizeof(int64_t) * (size_t)(N + 1));
    int64_t maxrow = 0;
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t e = ord[k];
        int64_t p = pos[ti[e]]++;
        g->src[p] = si[e];
        g->w[p] = weight[e];
    }
    for (int64_t i = 0; i < N; ++i) {
        int64_t n = g->rowptr[i + 1] - g->rowptr[i];
        if (n > maxrow) maxrow = n;
    }
    (void)maxrow;
    free(si); free(ti); free(cnt_s); free(ord); free(pos);
    *out = g;
    return VRO_OK;
}

/* Graph given directly as the CSR of P^T over vertex ids 0..N-1 (rows = targets, sources ascending
 * per row): skips the id sort of vro_sg_build.  Used by bench.py's CPU baseline. */
int vro_sg_from_csr(int64_t N, const int64_t *rowptr, const int32_t *src, const double *w, vro_sg **out)
{
    vro_sg *g = (vro_sg *)calloc(1, sizeof(vro_sg));
    if (!g) return VRO_ENOMEM;
    int64_t nnz = rowptr[N];
    g->N = N; g->nnz = nnz;
    g->ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N + 1));
    g->rowptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N + 1));
    g->src = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
    g->w = (double *)malloc(sizeof(double) * (size_t)(nnz + 1));
    if (!g->ids || !g->rowptr || !g->src || !g->w) { vro_sg_free(g); return VRO_ENOMEM; }
    for (int64_t i = 0; i < N; ++i) g->ids[i] = i;
    memcpy(g->rowptr, rowptr, sizeof(int64_t) * (size_t)(N + 1));
    memcpy(g->src, src, sizeof(int32_t) * (size_t)nnz);
    memcpy(g->w, w, sizeof(double) * (size_t)nnz);
    *out = g;
    return VRO_OK;
}

int64_t vro_sg_vertex_count(const vro_sg *g) { return g->N; }
void vro_sg_vertex_ids(const vro_sg *g, int64_t *out) { memcpy(out, g->ids, sizeof(int64_t) * (size_t)g->N); }

/* canonical sum of the terms x[src] * w of the in-edges [s, s + n) */
static double sg_canon_range(const vro_sg *g, const double *x, int64_t s, int64_t n, double *terms)
{
    if (n <= CANON_SEG) {
        for (int64_t k = 0; k < n; ++k) terms[k] = x[g->src[s + k]] * g->w[s + k];   /* :112 */
        return warp_sum(terms, n);
    }
    int64_t m = (n + CANON_SEG - 1) / CANON_SEG;
    double *part = (double *)malloc(sizeof(double) * (size_t)m);
    for (int64_t j = 0; j < m; ++j) {
        int64_t len = n - j * CANON_SEG; if (len > CANON_SEG) len = CANON_SEG;
        for (int64_t k = 0; k < len; ++k)
            terms[k] = x[g->src[s + j * CANON_SEG + k]] * g->w[s + j * CANON_SEG + k];
        part[j] = warp_sum(terms, len);
    }
    double r = warp_sum(part, m);
    free(part);
    return r;
}

/* sigma of one vertex in the canonical order */
static double sg_sigma(const vro_sg *g, const double *x, int64_t i, double *terms)
{
    int64_t s = g->rowptr[i], e = g->rowptr[i + 1];
    if (g->N <= (int64_t)CANON_SRC_BLOCK) return sg_canon_range(g, x, s, e - s, terms);
    /* large graph: one canonical sum per source block, added left to right */
    double total = 0.0;
    while (s < e) {
        int32_t blk = g->src[s] / CANON_SRC_BLOCK;
        int64_t t = s;
        while (t < e && g->src[t] / CANON_SRC_BLOCK == blk) t++;
        double sb = sg_canon_range(g, x, s, t - s, terms);
        total = total + sb;
        s = t;
    }
    return total;
}

/* calcNextX, stochastic/StochasticRecommender.scala:108-128.  Rows are
 * independent, so the OpenMP split changes no result. */
static void sg_next(vro_sg *g, const double *x, int64_t uidx, double alpha, double *nx)
{
    double one_minus = 1 - alpha;                               /* :121 (1 - alpha) */
#pragma omp parallel
    {
        double terms[CANON_SEG];
#pragma omp for schedule(dynamic, 2048)
        for (int64_t i = 0; i < g->N; ++i) {
            double sigma = sg_sigma(g, x, i, terms);            /* :113-114 groupBy(target).sum */
            double u = (i == uidx) ? 1.0 : 0.0;                 /* :81 */
            double a = u * alpha;                               /* :120 */
            double b = sigma * one_minus;                       /* :121 */
            nx[i] = a + b;
        }
    }
}

/* step / isConverged, stochastic/StochasticRecommender.scala:92-106,130-141.
 * out_x[N]; *iterations = the `iteration` value of the printed message;
 * *converged = 1 for "Converged in ...", 0 for "... reached the maximum".    */
int vro_sg_run(vro_sg *g, int64_t vertex_id, double epsilon, int32_t max_iterations,
               double *out_x, int32_t *iterations, int32_t *converged, double *last_residual)
{
    if (!(epsilon >= 0) || max_iterations < 0) return VRO_EINVAL;    /* :33-34 */
    int64_t v = sg_find(g, vertex_id);
    if (v < 0) return VRO_ENOENT;                                     /* :70 */
    const double alpha = 0.15;                                        /* :38 */
    double eps2 = epsilon * epsilon;                                  /* :40 */
    int64_t N = g->N;
    double *x = (double *)malloc(sizeof(double) * (size_t)N);
    double *nx = (double *)malloc(sizeof(double) * (size_t)N);
    if (!x || !nx) { free(x); free(nx); return VRO_ENOMEM; }
    double x0 = 1.0 / (double)N;                                      /* :53-54 */
    for (int64_t i = 0; i < N; ++i) x[i] = x0;
    int32_t it = 0; int conv = 0; double res = -1.0;
    for (;;) {
        if (it >= max_iterations) { conv = 0; break; }                /* :93-95 */
        sg_next(g, x, v, alpha, nx);
        res = 0.0;
        for (int64_t i = 0; i < N; ++i) {                             /* :131-139 */
            double dlt = nx[i] - x[i];
            double sq = dlt * dlt;
            res = res + sq;
        }
        double *tmp = x; x = nx; nx = tmp;
        if (res <= eps2) { conv = 1; break; }                         /* :140, returns nextX */
        it++;
    }
    memcpy(out_x, x, sizeof(double) * (size_t)N);
    *iterations = it; *converged = conv;
    if (last_residual) *last_residual = res;
    free(x); free(nx);
    return VRO_OK;
}

/* makeRecommendations0's filter (id != vertex and probability > 0, :85-88) and
 * StochasticRecommenderMain.printRecommendations' join with the target
 * region's places + orderBy(probability desc).limit(N)
 * (stochastic/StochasticRecommenderMain.scala:69-73).  filter == NULL keeps
 * every vertex.                                                               */
int vro_sg_query(vro_sg *g, int64_t vertex_id, double epsilon, int32_t max_iterations,
                 const int64_t *filter, int64_t n_filter, int32_t max_recs,
                 int64_t *out_id, double *out_prob, int32_t *out_count,
                 int32_t *iterations, int32_t *converged)
{
    if (max_recs < 0) return VRO_EINVAL;
    double *x = (double *)malloc(sizeof(double) * (size_t)(g->N > 0 ? g->N : 1));
    if (!x) return VRO_ENOMEM;
    int rc = vro_sg_run(g, vertex_id, epsilon, max_iterations, x, iterations, converged, 0);
    if (rc) { free(x); return rc; }
    rec_t *recs = (rec_t *)malloc(sizeof(rec_t) * (size_t)(g->N > 0 ? g->N : 1));
    if (!recs) { free(x); return VRO_ENOMEM; }
    int64_t nr = 0;
    if (filter) {
        for (int64_t k = 0; k < n_filter; ++k) {
            int64_t i = sg_find(g, filter[k]);
            if (i < 0 || g->ids[i] == vertex_id || !(x[i] > 0)) continue;
            recs[nr].est = x[i]; recs[nr].place = g->ids[i]; nr++;
        }
        /* a filter may list an id twice: keep one */
        qsort(recs, (size_t)nr, sizeof(rec_t), rec_cmp_rank);
        int64_t m = 0;
        for (int64_t k = 0; k < nr; ++k)
            if (k == 0 || recs[k].place != recs[k - 1].place || recs[k].est != recs[k - 1].est) recs[m++] = recs[k];
        nr = m;
    } else {
        for (int64_t i = 0; i < g->N; ++i) {
            if (g->ids[i] == vertex_id || !(x[i] > 0)) continue;
            recs[nr].est = x[i]; recs[nr].place = g->ids[i]; nr++;
        }
        qsort(recs, (size_t)nr, sizeof(rec_t), rec_cmp_rank);
    }
    if (nr > max_recs) nr = max_recs;
    for (int64_t k = 0; k < nr; ++k) { out_id[k] = recs[k].place; out_prob[k] = recs[k].est; }
    *out_count = (int32_t)nr;
    free(recs); free(x);
    return VRO_OK;
}

void vro_set_threads(int32_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int vro_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* knn/RatingsBuilder.scala + knn/RatingVectorsBuilder.scala           */
/* (the step in front of the KNN path)          */
/* ------------------------------------------------------------------ */

typedef struct { int64_t person, entity, weight; } visit_t;

static int visit_cmp(const void *a, const void *b)
{
    const visit_t *x = (const visit_t *)a, *y = (const visit_t *)b;
    if (x->person != y->person) return x->person < y->person ? -1 : 1;
    if (x->entity != y->entity) return x->entity < y->entity ? -1 : 1;
    return 0;
}

/* RatingsBuilder.calcRatings (knn/RatingsBuilder.scala:32-48): count(*) per (person_id, entity) (:38-40; a
 * row may stand for `weight` visits), rank() over (partition by person_id order by count desc) (:42-45),
 * keep rank <= top_n (:46) -- rank() = 1 + number of rows of the person with a larger count, so ties
 * all stay.  Then RatingVectorsBuilder.calcRatingVectors (knn/RatingVectorsBuilder.scala:12-83): size =
 * max(entity id) + 1 (:26-34), ids must fit an Int (:36-41, VRO_EINVAL), indices ascending (:45-50), values
 * = counts as doubles (:69).  Outputs in CSR form, persons ascending.                                      */
int vro_build_rating_vectors(int64_t n_rows, const int64_t *person_id, const int64_t *entity_id,
                             const int64_t *weight, int32_t top_n, int64_t *out_n_persons, int64_t *out_nnz,
                             int64_t *out_person_id, int64_t *out_rowptr, int32_t *out_col, double *out_val,
                             int32_t *out_dim)
{
    if (n_rows < 0 || top_n <= 0) return VRO_EINVAL;
    *out_n_persons = 0; *out_nnz = 0; *out_dim = 1; out_rowptr[0] = 0;
    if (n_rows == 0) return VRO_OK;
    visit_t *v = (visit_t *)malloc(sizeof(visit_t) * (size_t)n_rows);
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_rows);
    if (!v || !cnt) { free(v); free(cnt); return VRO_ENOMEM; }
    for (int64_t i = 0; i < n_rows; ++i) {
        v[i].person = person_id[i]; v[i].entity = entity_id[i]; v[i].weight = weight ? weight[i] : 1;
    }
    qsort(v, (size_t)n_rows, sizeof(visit_t), visit_cmp);
    /* runs of equal (person, entity): v[0..m) keeps one row per run, cnt[] its count */
    int64_t m = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        if (m > 0 && v[m - 1].person == v[i].person && v[m - 1].entity == v[i].entity) {
            cnt[m - 1] += v[i].weight;
        } else {
            v[m] = v[i]; cnt[m] = v[i].weight; m++;
        }
    }
    int64_t P = 0, nnz = 0, maxid = -1;
    int rc = VRO_OK;
    for (int64_t s = 0; s < m && rc == VRO_OK; ) {
        int64_t e = s;
        while (e < m && v[e].person == v[s].person) e++;
        out_person_id[P] = v[s].person;
        for (int64_t r = s; r < e; ++r) {
            int64_t larger = 0;
            for (int64_t q = s; q < e; ++q) larger += cnt[q] > cnt[r];
            if (larger + 1 <= top_n) {
                if (v[r].entity < 0 || v[r].entity > 0x7fffffffLL) { rc = VRO_EINVAL; break; }
                out_col[nnz] = (int32_t)v[r].entity;
                out_val[nnz] = (double)cnt[r];
                if (v[r].entity > maxid) maxid = v[r].entity;
                nnz++;
            }
        }
        out_rowptr[++P] = nnz;
        s = e;
    }
    free(v); free(cnt);
    if (rc != VRO_OK) return rc;
    *out_n_persons = P; *out_nnz = nnz; *out_dim = (int32_t)(maxid + 1);
    return VRO_OK;
}

/* ------------------------------------------------------------------ */
/* the four edge families + StochasticGraphBuilder */
/* ------------------------------------------------------------------ */

/* One edge family, e.g. PersonLikesPlace.calcPersonLikesPlaceEdges (stochastic/PersonLikesPlace.scala:12-41):
 * count(*) per (source, target) (:13-15), rank() <= top_n per source by count desc (:17-23, ties stay),
 * weight = count / (sum of the kept counts of the source) as doubles (:25-32), then
 * StochasticGraphBuilder.buildWithBalancedWeights' `weight * beta` (stochastic/StochasticGraphBuilder.scala:8-28).
 * Output sorted by (source, target).  Returns the number of edges in *out_n; fills the arrays if it fits.   */
int vro_build_edge_family(int64_t n_rows, const int64_t *source_id, const int64_t *target_id, const int64_t *weight,
                          int32_t top_n, double beta, int64_t capacity, int64_t *out_n, int64_t *out_source,
                          int64_t *out_target, double *out_weight)
{
    if (n_rows < 0 || top_n <= 0) return VRO_EINVAL;
    *out_n = 0;
    if (n_rows == 0) return VRO_OK;
    visit_t *v = (visit_t *)malloc(sizeof(visit_t) * (size_t)n_rows);
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_rows);
    if (!v || !cnt) { free(v); free(cnt); return VRO_ENOMEM; }
    for (int64_t i = 0; i < n_rows; ++i) {
        v[i].person = source_id[i]; v[i].entity = target_id[i]; v[i].weight = weight ? weight[i] : 1;
    }
    qsort(v, (size_t)n_rows, sizeof(visit_t), visit_cmp);
    int64_t m = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        if (m > 0 && v[m - 1].person == v[i].person && v[m - 1].entity == v[i].entity) cnt[m - 1] += v[i].weight;
        else { v[m] = v[i]; cnt[m] = v[i].weight; m++; }
    }
    int64_t ne = 0;
    for (int64_t s = 0; s < m; ) {
        int64_t e = s;
        while (e < m && v[e].person == v[s].person) e++;
        int64_t total = 0;
        for (int64_t r = s; r < e; ++r) {                     /* first pass: which rows stay, and their total */
            int64_t larger = 0;
            for (int64_t q = s; q < e; ++q) larger += cnt[q] > cnt[r];
            v[r].weight = larger + 1 <= top_n;                /* reuse the field as the keep flag */
            if (v[r].weight) total += cnt[r];
        }
        for (int64_t r = s; r < e; ++r) {
            if (!v[r].weight) continue;
            if (ne < capacity && out_source) {
                double w = (double)cnt[r] / (double)total;
                out_source[ne] = v[r].person;
                out_target[ne] = v[r].entity;
                out_weight[ne] = w * beta;
            }
            ne++;
        }
        s = e;
    }
    free(v); free(cnt);
    *out_n = ne;
    return ne > capacity ? VRO_ENOMEM : VRO_OK;
}

typedef struct { int64_t person, place, ts; } pvisit_t;
static int pvisit_cmp(const void *a, const void *b)
{
    const pvisit_t *x = (const pvisit_t *)a, *y = (const pvisit_t *)b;
    return x->person < y->person ? -1 : x->person > y->person;
}

/* StochasticGraphBuilderMain.generateStochasticGraph (stochastic/StochasticGraphBuilderMain.scala:47-66): the
 * union, in this order, of PlaceSimilarPlace (the visits self-joined on person_id, different places, timestamps
 * at most 7 days apart, one row per pair of visits; top 50; stochastic/PlaceSimilarPlace.scala:13-63),
 * CategorySelectedPlace (top 100), PersonLikesPlace (top 100, beta_person_place) and PersonLikesCategory
 * (top 100, beta_person_category); the first two have beta = 1 (:8-9).                                      */
int vro_build_stochastic_graph(int64_t n, const int64_t *person_id, const int64_t *place_id,
                               const int64_t *category_id, const int64_t *timestamp_ms, double beta_person_place,
                               double beta_person_category, int64_t capacity, int64_t *out_n, int64_t *out_source,
                               int64_t *out_target, double *out_weight)
{
    if (n < 0) return VRO_EINVAL;
    *out_n = 0;
    if (n == 0) return VRO_OK;
    const int64_t interval = 7LL * 24 * 3600 * 1000;
    pvisit_t *pv = (pvisit_t *)malloc(sizeof(pvisit_t) * (size_t)n);
    if (!pv) return VRO_ENOMEM;
    for (int64_t i = 0; i < n; ++i) { pv[i].person = person_id[i]; pv[i].place = place_id[i]; pv[i].ts = timestamp_ms[i]; }
    qsort(pv, (size_t)n, sizeof(pvisit_t), pvisit_cmp);
    int64_t np = 0;
    for (int64_t s = 0; s < n; ) {
        int64_t e = s;
        while (e < n && pv[e].person == pv[s].person) e++;
        for (int64_t i = s; i < e; ++i)
            for (int64_t j = s; j < e; ++j) {
                int64_t dt = pv[i].ts - pv[j].ts; if (dt < 0) dt = -dt;
                np += pv[i].place != pv[j].place && dt <= interval;
            }
        s = e;
    }
    int64_t *pa = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np > 0 ? np : 1));
    int64_t *pb = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np > 0 ? np : 1));
    if (!pa || !pb) { free(pv); free(pa); free(pb); return VRO_ENOMEM; }
    int64_t k = 0;
    for (int64_t s = 0; s < n; ) {
        int64_t e = s;
        while (e < n && pv[e].person == pv[s].person) e++;
        for (int64_t i = s; i < e; ++i)
            for (int64_t j = s; j < e; ++j) {
                int64_t dt = pv[i].ts - pv[j].ts; if (dt < 0) dt = -dt;
                if (pv[i].place != pv[j].place && dt <= interval) { pa[k] = pv[i].place; pb[k] = pv[j].place; k++; }
            }
        s = e;
    }
    free(pv);
    int64_t at = 0, got = 0;
    int rc = VRO_OK, r;
    const int64_t *fs[4] = {pa, category_id, person_id, person_id};
    const int64_t *fd[4] = {pb, place_id, place_id, category_id};
    const int64_t fn[4] = {np, n, n, n};
    const int32_t ft[4] = {50, 100, 100, 100};
    const double fb[4] = {1.0, 1.0, beta_person_place, beta_person_category};
    for (int f = 0; f < 4; ++f) {
        int64_t room = capacity > at ? capacity - at : 0;
        r = vro_build_edge_family(fn[f], fs[f], fd[f], 0, ft[f], fb[f], room, &got,
                                  out_source ? out_source + (at < capacity ? at : 0) : 0,
                                  out_target ? out_target + (at < capacity ? at : 0) : 0,
                                  out_weight ? out_weight + (at < capacity ? at : 0) : 0);
        if (r != VRO_OK && r != VRO_ENOMEM) { rc = r; break; }
        if (r == VRO_ENOMEM) rc = VRO_ENOMEM;
        at += got;
    }
    free(pa); free(pb);
    *out_n = at;
    return rc;
}

/* ------------------------------------------------------------------ */
/* PlaceVisits.calcPlaceVisits                  */
/* ------------------------------------------------------------------ */

/* Location.distanceMeters, Location.scala:30-43 (haversine; libm here, FastMath in the reference) */
double vro_distance_meters(double lat1d, double lon1d, double lat2d, double lon2d)
{
    const double k = 0.017453292519943295;
    double lat1 = lat1d * k, lat2 = lat2d * k, lon1 = lon1d * k, lon2 = lon2d * k;
    double s1 = sin((lat2 - lat1) / 2), s2 = sin((lon2 - lon1) / 2);
    double hav = s1 * s1 + cos(lat1) * cos(lat2) * (s2 * s2);
    return 6371000.0 * 2 * asin(sqrt(hav));
}

/* PlaceVisits.calcPlaceVisits (PlaceVisits.scala:11-48): the visits with timestamp >= max(timestamp) -
 * last_days_count days (:50-61, whole UTC days here) joined with the places of the same region (:33), kept where
 * the distance is <= accuracy_m (:13-22).  This is the reference's "almost cross-join", row by row.  Rows in
 * visit order, a visit's places in ascending place id; out_margin[k] = |distance - accuracy| of row k and
 * *out_closest_miss = the smallest such margin among the rejected pairs (so a caller comparing another
 * implementation can tell boundary cases apart).                                                               */
int vro_build_place_visits(int64_t n_visits, const int64_t *person_id, const double *latitude, const double *longitude,
                           const int64_t *timestamp_ms, const int64_t *region_id, int64_t n_places,
                           const int64_t *place_id, const double *place_latitude, const double *place_longitude,
                           const int64_t *place_category, const int64_t *place_region, int32_t last_days_count,
                           double accuracy_m, int64_t capacity, int64_t *out_n, int64_t *out_person,
                           int64_t *out_timestamp_ms, int64_t *out_place, int64_t *out_region, int64_t *out_category,
                           double *out_margin, double *out_closest_miss)
{
    if (n_visits < 0 || n_places < 0 || !(accuracy_m > 0) || last_days_count < 0) return VRO_EINVAL;
    *out_n = 0;
    if (out_closest_miss) *out_closest_miss = 1e300;
    if (n_visits == 0 || n_places == 0) return VRO_OK;
    int64_t ts_max = timestamp_ms[0];
    for (int64_t i = 1; i < n_visits; ++i) if (timestamp_ms[i] > ts_max) ts_max = timestamp_ms[i];
    const int64_t ts_from = ts_max - (int64_t)last_days_count * 86400000LL;
    /* places ordered by id so that a visit's rows come out in ascending place id */
    int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_places);
    if (!ord) return VRO_ENOMEM;
    for (int64_t p = 0; p < n_places; ++p) ord[p] = p;
    for (int64_t p = 1; p < n_places; ++p) {               /* insertion sort is enough for the oracle's sizes */
        int64_t v = ord[p], q = p;
        while (q > 0 && place_id[ord[q - 1]] > place_id[v]) { ord[q] = ord[q - 1]; q--; }
        ord[q] = v;
    }
    int64_t n = 0;
    double miss = 1e300;
    for (int64_t i = 0; i < n_visits; ++i) {
        if (timestamp_ms[i] < ts_from) continue;
        for (int64_t k = 0; k < n_places; ++k) {
            int64_t p = ord[k];
            if (place_region[p] != region_id[i]) continue;
            double d = vro_distance_meters(latitude[i], longitude[i], place_latitude[p], place_longitude[p]);
            if (d <= accuracy_m) {
                if (n < capacity && out_person) {
                    out_person[n] = person_id[i]; out_timestamp_ms[n] = timestamp_ms[i]; out_place[n] = place_id[p];
                    out_region[n] = region_id[i]; out_category[n] = place_category[p];
                    if (out_margin) out_margin[n] = accuracy_m - d;
                }
                n++;
            } else if (d - accuracy_m < miss) {
                miss = d - accuracy_m;
            }
        }
    }
    free(ord);
    *out_n = n;
    if (out_closest_miss) *out_closest_miss = miss;
    return n > capacity ? VRO_ENOMEM : VRO_OK;
}
