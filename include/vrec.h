/*
 * vrec.h -- C ABI of libvrec.so, the B200-native engine for the two hot paths
 * of the Visit Recommender (tashoyan/locations-recommender).
 *
 * The reference has NO native/FFI seam: the seam is the pair of Scala classes
 *   recommender/src/main/scala/com/github/tashoyan/recommender/knn/KnnRecommender.scala:9-25
 *   recommender/src/main/scala/com/github/tashoyan/recommender/stochastic/StochasticRecommender.scala:28-71
 * and their only callers (knn/KnnRecommenderMain.scala:53-107,
 * stochastic/StochasticRecommenderMain.scala:53-81).  A JVM-side drop-in
 * `KnnRecommender` / `StochasticRecommender` with the same constructors binds
 * exactly the entry points below through JNI.
 *
 * Conventions: plain C, opaque handles, caller-owned host buffers (may be freed
 * as soon as a call returns), library-owned device memory.  Every call returns
 * VREC_OK (0) or a negative VREC_E* code; vrec_last_error() gives the message
 * of the last failure on the calling thread.  A context may be used by one
 * thread at a time (the reference has one REPL thread).  There is NO CPU
 * fallback: without an sm_100 device vrec_init fails with VREC_ENODEV.
 *
 * Paths cited below are relative to
 * /root/reference/recommender/src/main/scala/com/github/tashoyan/recommender/.
 */
#ifndef VREC_H
#define VREC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VREC_OK       0
#define VREC_ENOENT  (-2)    /* "No such person: <id>" / "No such vertex in the graph: <id>"      */
#define VREC_ENOMEM  (-12)
#define VREC_ENODEV  (-19)   /* no sm_100 CUDA device                                              */
#define VREC_EINVAL  (-22)   /* a Scala `require(...)` of the reference would have failed          */
#define VREC_ECUDA   (-100)
#define VREC_ENCCL   (-101)

#define VREC_ABI_VERSION 1

typedef struct vrec_ctx vrec_ctx;
typedef struct vrec_knn vrec_knn;
typedef struct vrec_sg  vrec_sg;

/* ---------------------------------------------------------------- context */

/* One context per process and GPU (one process per GPU).  device < 0 picks
 * the current device.  Replaces `SparkSession.builder().getOrCreate()`
 * (knn/KnnRecommenderMain.scala:20-21) as the owner of execution resources. */
int vrec_init(int device, vrec_ctx **out);
void vrec_shutdown(vrec_ctx *ctx);
const char *vrec_last_error(void);
int vrec_abi_version(void);

/* cudaStream_t every kernel of this context is launched on (for CUDA-event
 * timing by the caller), and the number of kernels launched so far.         */
void *vrec_stream(vrec_ctx *ctx);
int64_t vrec_launch_count(vrec_ctx *ctx);
int vrec_synchronize(vrec_ctx *ctx);

/* ---------------------------------------------------------------- multi-GPU
 * One process per GPU.  Rank 0 calls vrec_comm_unique_id and ships the 128 bytes to the other
 * ranks by any means (bench.py uses torch.distributed); every rank then calls vrec_comm_init.
 * Region-sharded work (independent region-sets / target ranges) needs no communicator; it is
 * required only for the row-partitioned graph (vrec_sg_load_partitioned / vrec_sg_generate
 * with world > 1): the ranks exchange CUDA IPC handles of their x buffers over it once per graph;
 * the per-iteration exchange of x' is then fused into the sweep kernel (peer stores over NVLink,
 * one flag barrier per iteration, csrc/vrec_sg.cu) -- no collective runs inside the iteration.  */
int vrec_comm_unique_id(void *out128);
int vrec_comm_init(vrec_ctx *ctx, int rank, int world, const void *unique_id128);
int vrec_comm_rank(vrec_ctx *ctx);
int vrec_comm_world(vrec_ctx *ctx);

/* ---------------------------------------------------------------- KNN path */

/*
 * Loads one region-set: the three DataFrames handed to
 * `new KnnRecommender(placeRatingVectors, categoryRatingVectors, placeRatings, ...)`
 * (knn/KnnRecommender.scala:9-16; files of knn/KnnRecommenderMain.scala:69-88).
 *
 *  person_id[P]                    person ids (any order; sorted internally)
 *  place_rowptr[P+1], place_col, place_val, place_dim
 *                                  `place_rating_vectors`: row i is the SparseVector of
 *                                  person i (indices ascending, < place_dim = SparseVector.size;
 *                                  knn/RatingVectorsBuilder.scala:74-83).  An empty row means
 *                                  "person has no row in this table".
 *  cat_rowptr[P+1], cat_col, cat_val, cat_dim      `category_rating_vectors`, same layout
 *  n_ratings, rating_person, rating_place, rating_value
 *                                  `place_ratings` rows (person_id, place_id, rating)
 *                                  (knn/RatingsBuilder.scala:32-48).  rating_person == NULL
 *                                  derives them from the place vectors (they hold the same
 *                                  non-zeros, knn/RatingVectorsBuilderMain.scala:41-73).
 */
int vrec_knn_load(vrec_ctx *ctx, int64_t P, const int64_t *person_id,
                  const int64_t *place_rowptr, const int32_t *place_col, const double *place_val,
                  int32_t place_dim,
                  const int64_t *cat_rowptr, const int32_t *cat_col, const double *cat_val,
                  int32_t cat_dim,
                  int64_t n_ratings, const int64_t *rating_person, const int64_t *rating_place,
                  const int64_t *rating_value,
                  vrec_knn **out);
void vrec_knn_free(vrec_knn *knn);

/*
 * KnnRecommender.makeRecommendations (knn/KnnRecommender.scala:22-25) for a batch of
 * targets, followed by KnnRecommenderMain.printRecommendations' join with the target
 * region's places and `orderBy(estimated_rating desc).limit(max_recs)`
 * (knn/KnnRecommenderMain.scala:96-102).
 *
 *  place_weight, category_weight, k_nearest   ctor arguments; the `require`s of
 *                                  knn/KnnRecommender.scala:17-20 give VREC_EINVAL
 *  place_filter[n_filter]          ids of the places of the target region
 *                                  (`places.where(region_id === target)`); NULL = all places
 *  out_place_id / out_rating       [n_targets x max_recs], row t holds out_count[t] rows ordered
 *                                  by (estimated_rating desc, place_id asc)
 *  out_status[t]                   VREC_OK, or VREC_ENOENT = IllegalArgumentException
 *                                  "No such person: <id>" (knn/KnnRecommender.scala:83)
 */
int vrec_knn_query(vrec_knn *knn, const int64_t *targets, int32_t n_targets,
                   double place_weight, double category_weight, int32_t k_nearest,
                   const int64_t *place_filter, int64_t n_filter, int32_t max_recs,
                   int64_t *out_place_id, double *out_rating, int32_t *out_count,
                   int32_t *out_status);

/* Same, with targets and outputs already resident in device memory (the place filter
 * is installed beforehand with vrec_knn_set_filter).  Asynchronous on vrec_stream(). */
int vrec_knn_set_filter(vrec_knn *knn, const int64_t *place_filter, int64_t n_filter);
int vrec_knn_query_device(vrec_knn *knn, const int64_t *d_targets, int32_t n_targets,
                          double place_weight, double category_weight, int32_t k_nearest,
                          int32_t max_recs,
                          int64_t *d_out_place_id, double *d_out_rating, int32_t *d_out_count,
                          int32_t *d_out_status);

/* findSimilarPersons (knn/KnnRecommender.scala:27-49): the k_nearest most similar persons,
 * ordered by (similarity desc, person_id asc).  capacity = size of the out arrays.          */
int vrec_knn_neighbours(vrec_knn *knn, int64_t target, double place_weight, double category_weight,
                        int32_t k_nearest, int64_t *out_person_id, double *out_similarity,
                        int32_t capacity, int32_t *out_count);

/* The raw DataFrame of makeRecommendations (place_id, estimated_rating), every place rated by
 * at least one neighbour, sorted by place_id (knn/KnnRecommender.scala:51-70).              */
int vrec_knn_estimates(vrec_knn *knn, int64_t target, double place_weight, double category_weight,
                       int32_t k_nearest, int64_t *out_place_id, double *out_rating,
                       int64_t capacity, int64_t *out_count);

/* Person ids in the engine's (ascending) order: out[P]. */
int vrec_knn_person_ids(vrec_knn *knn, int64_t *out);

/* Dense similarity vector of one target after `orderBy(similarity desc).limit(k_nearest)`
 * (knn/KnnRecommender.scala:38-48): out_sim[P] in vrec_knn_person_ids order, 0.0 for persons
 * that are not among the k nearest.  Works for any k_nearest.                               */
int vrec_knn_similarities(vrec_knn *knn, int64_t target, double place_weight, double category_weight,
                          int32_t k_nearest, double *out_sim);

/* Tuning / test knobs.  "rating_path": 0 = automatic, 1 = neighbour-row gather,
 * 2 = column scan.  "tile": targets per pass; "splits": candidate ranges per target in the fused
 * top-K kernel (0 = automatic).  Unknown names give VREC_EINVAL.                */
int vrec_knn_set_option(vrec_knn *knn, const char *name, int64_t value);
int64_t vrec_knn_resident_bytes(vrec_knn *knn);
/* Debug counters of the tiled similarity kernel since the last call: out4 = {exact evaluations
 * from the postings pass, exact evaluations of filter survivors, heap inserts, queue overflows}. */
int vrec_knn_debug_stats(vrec_knn *knn, uint64_t *out4);
/* Debug: SM cycles block 0 of the tensor-core kernel spent per phase since the last call: out8 =
 * {B-tile wait, MMA, TMEM epilogue, barriers + queue drain, postings pass, tiles, -, -}.          */
int vrec_knn_debug_tc_cycles(vrec_knn *knn, uint64_t *out8);
/* Debug: cycles per block of the last tensor-core main pass: out[0..n) dense phase, out[n..2n) postings. */
int vrec_knn_debug_tc_block_cycles(vrec_knn *knn, uint64_t *out, int n);
/* Duration in milliseconds of the last dense-filter kernel launch (knn_tc_ws_kernel), from CUDA events on
 * vrec_stream(); 0 if that kernel has not run.  bench.py's roofline uses it.                          */
int vrec_knn_last_dense_ms(vrec_knn *knn, double *out_ms);
/* Debug: neighbours (person_id ascending, similarity) of target index t of the last vrec_knn_query pass that ran
 * the fused top-K kernels with this K; what makeRecommendations0 (knn/KnnRecommender.scala:51-70) consumed. */
int vrec_knn_debug_last_neighbours(vrec_knn *knn, int32_t t, int32_t K, int64_t *out_person_id,
                                   double *out_similarity, int32_t *out_count);
/* Debug: exact-evaluation probe of block 0 / thread 0 of the dense kernel, cycles summed over its
 * evaluations: out8 = {meta words, record headers, place matching, evaluations, heap inserts (cycles),
 * heap inserts (count), place section load, category section + dense row}.                            */
int vrec_knn_debug_probe(vrec_knn *knn, uint64_t *out8);

/* ---------------------------------------------------------------- rating vectors builder */

/*
 * The step in front of the KNN path, on the device: RatingsBuilder.calcRatings
 * (knn/RatingsBuilder.scala:32-48: count(*) per (person_id, entity), keep rank() <= top_n per person
 * ordered by count desc -- rank keeps ties) followed by RatingVectorsBuilder.calcRatingVectors
 * (knn/RatingVectorsBuilder.scala:12-83: vector size = max(entity id) + 1, indices ascending, values =
 * counts as doubles), for ONE entity column; call it with place_id and with category_id of the same
 * rows (RatingVectorsBuilderMain.scala:41-63) to get the two tables vrec_knn_load takes.
 *
 *  person_id / entity_id [n_rows]  one row per visit (any order)
 *  weight [n_rows] or NULL         visits the row stands for (NULL = 1: count(*))
 *  out_person_id [<= n_rows]       ascending;  out_rowptr [<= n_rows + 1];  out_col / out_val [<= n_rows]
 *  out_dim                         max kept entity id + 1
 * An entity id outside [0, 2^31) gives VREC_EINVAL ("Index out of Int range", :36-41).
 */
int vrec_build_rating_vectors(vrec_ctx *ctx, int64_t n_rows, const int64_t *person_id, const int64_t *entity_id,
                              const int64_t *weight, int32_t top_n, int64_t *out_n_persons, int64_t *out_nnz,
                              int64_t *out_person_id, int64_t *out_rowptr, int32_t *out_col, double *out_val,
                              int32_t *out_dim);

/*
 * PlaceVisits.calcPlaceVisits (PlaceVisits.scala:11-48) with the spatial grid its authors ask for (":30 TODO Very
 * inefficient almost cross-join"): every location visit of the last `last_days_count` days (counted in whole UTC
 * days from the latest timestamp, :50-61) becomes a place visit of every place of its region within `accuracy_m`
 * metres (haversine, Location.scala:30-43; 100 m in the reference, PlaceVisits.scala:127).  Output rows
 * (person_id, timestamp, place_id, region_id, category_id) in visit order, a visit's places in ascending id --
 * the input of the two builders above / below.  Distances are fp64 with CUDA's libm: a pair whose distance is
 * within a few ulps of `accuracy_m` may fall on the other side than with the reference's FastMath.
 * *out_n = rows; VREC_ENOMEM (after setting *out_n) if capacity is too small or the outputs are NULL.
 */
int vrec_build_place_visits(vrec_ctx *ctx, int64_t n_visits, const int64_t *person_id, const double *latitude,
                            const double *longitude, const int64_t *timestamp_ms, const int64_t *region_id,
                            int64_t n_places, const int64_t *place_id, const double *place_latitude,
                            const double *place_longitude, const int64_t *place_category, const int64_t *place_region,
                            int32_t last_days_count, double accuracy_m, int64_t capacity, int64_t *out_n,
                            int64_t *out_person, int64_t *out_timestamp_ms, int64_t *out_place, int64_t *out_region,
                            int64_t *out_category);

/*
 * The step in front of the SG path, on the device.
 *  vrec_build_edge_family        one edge family of the stochastic graph, e.g. PersonLikesPlace
 *      (stochastic/PersonLikesPlace.scala:12-41): count(*) per (source, target) (a row may stand for `weight`
 *      rows), rank() <= top_n per source by count desc (ties stay), weight = count / (sum of the source's kept
 *      counts), times beta (stochastic/StochasticGraphBuilder.scala:8-28).  Edges sorted by (source, target).
 *  vrec_build_stochastic_graph   StochasticGraphBuilderMain.generateStochasticGraph
 *      (stochastic/StochasticGraphBuilderMain.scala:47-66) from place visits (person, place, category,
 *      timestamp in ms): PlaceSimilarPlace (visits of one person at different places at most 7 days apart,
 *      top 50), CategorySelectedPlace (top 100), PersonLikesPlace (top 100, beta_person_place),
 *      PersonLikesCategory (top 100, beta_person_category), in this order -- the input of vrec_sg_load.
 * *out_n = number of edges; with capacity too small (or NULL outputs) the call returns VREC_ENOMEM after
 * setting *out_n, so a first call can size the arrays.
 */
int vrec_build_edge_family(vrec_ctx *ctx, int64_t n_rows, const int64_t *source_id, const int64_t *target_id,
                           const int64_t *weight, int32_t top_n, double beta, int64_t capacity, int64_t *out_n,
                           int64_t *out_source, int64_t *out_target, double *out_weight);
int vrec_build_stochastic_graph(vrec_ctx *ctx, int64_t n_visits, const int64_t *person_id, const int64_t *place_id,
                                const int64_t *category_id, const int64_t *timestamp_ms, double beta_person_place,
                                double beta_person_category, int64_t capacity, int64_t *out_n, int64_t *out_source,
                                int64_t *out_target, double *out_weight);

/* ---------------------------------------------------------------- SG path */

/*
 * Loads one stochastic graph: the DataFrame (source_id, target_id, balanced_weight) handed to
 * `new StochasticRecommender(stochasticEdges, epsilon, maxIterations)`
 * (stochastic/StochasticRecommender.scala:28-32; file of
 * stochastic/StochasticRecommenderMain.scala:83-89).  Builds the vertex set
 * distinct(source ∪ target) (:42-49) and the CSR of P^T on the device.
 */
int vrec_sg_load(vrec_ctx *ctx, int64_t nnz, const int64_t *source_id, const int64_t *target_id,
                 const double *balanced_weight, vrec_sg **out);
/* Same, for ONE oversized graph on several GPUs: every rank passes the whole edge list and keeps a
 * contiguous block of rows of P^T, balanced by in-edges (vrec_sg_row_range); queries must then be
 * issued by all ranks together (same arguments); every rank gets the full result.  The sweep of
 * calcNextX (stochastic/StochasticRecommender.scala:108-128) stores every rank's rows of x' straight
 * into the peers' buffers; the residual partials ride in a slot per rank and one flag barrier per
 * iteration replaces the reference's per-iteration Spark action (:130-141).                      */
int vrec_sg_load_partitioned(vrec_ctx *ctx, int64_t nnz, const int64_t *source_id, const int64_t *target_id,
                             const double *balanced_weight, vrec_sg **out);
int vrec_sg_row_range(vrec_sg *sg, int64_t *out_lo, int64_t *out_hi);   /* rows of P^T this process owns */
/* Single-process group: the `world` parts of one row-partitioned graph held by the calling process
 * (same kernels, peer stores and residual slots as the one-process-per-GPU path; the exchange
 * barrier is the stream order).  Serves hosts that drive the parts from one process and is how the
 * partition logic is verified where a single GPU is visible.  out_parts[world];
 * vrec_sg_group_stationary: out_x[world x vertex_count] (every part's copy of the result),
 * out_iterations / out_converged / out_residual [world].  Free every part with vrec_sg_free.     */
int vrec_sg_group_load(vrec_ctx *ctx, int32_t world, int64_t nnz, const int64_t *source_id,
                       const int64_t *target_id, const double *balanced_weight, vrec_sg **out_parts);
int vrec_sg_group_stationary(vrec_sg **parts, int32_t world, int64_t vertex, double epsilon,
                             int32_t max_iterations, double *out_x, int32_t *out_iterations,
                             int32_t *out_converged, double *out_residual);
void vrec_sg_free(vrec_sg *sg);
int64_t vrec_sg_vertex_count(vrec_sg *sg);
int64_t vrec_sg_edge_count(vrec_sg *sg);
int vrec_sg_vertex_ids(vrec_sg *sg, int64_t *out_ids);   /* [vertex_count], ascending */

/*
 * StochasticRecommender.makeRecommendations (stochastic/StochasticRecommender.scala:66-90) for a
 * batch of vertices, followed by StochasticRecommenderMain.printRecommendations' join with the
 * target region's places and `orderBy(probability desc).limit(max_recs)`
 * (stochastic/StochasticRecommenderMain.scala:69-73).
 *
 *  epsilon, max_iterations         ctor arguments (`require`s of :33-34 give VREC_EINVAL)
 *  place_filter[n_filter]          ids of the target region's places; NULL = every vertex
 *  out_id / out_prob               [n x max_recs], ordered by (probability desc, id asc)
 *  out_iterations[q]               the `iteration` printed by step() (:94,:100)
 *  out_converged[q]                1 = "Converged in ...", 0 = "... reached the maximum ..."
 *  out_status[q]                   VREC_OK or VREC_ENOENT ("No such vertex in the graph: <id>", :70)
 */
int vrec_sg_query(vrec_sg *sg, const int64_t *vertices, int32_t n,
                  double epsilon, int32_t max_iterations,
                  const int64_t *place_filter, int64_t n_filter, int32_t max_recs,
                  int64_t *out_id, double *out_prob, int32_t *out_count,
                  int32_t *out_iterations, int32_t *out_converged, int32_t *out_status);

/* Tuning / inspection of the batch kernel that serves start vertices without in-edges (every
 * person vertex of the reference's graphs; csrc/vrec_sg_batch.cu).  Results are identical either way.
 *  vrec_sg_set_option  "batch": 0 = per-query kernels only, 1 = batch kernel when a call has >= 4
 *                      eligible start vertices (default), 2 = for every eligible start vertex;
 *                      "batch_targets_per_cta": 0 = auto, 1 or 2;
 *                      "rows_kernel": 1 = the round-1 half-warp-per-row sweep instead of the flat-window
 *                      sweep (A/B measurements); "flat_variant": 0..2 = launch shape of the latter;
 *                      "graph": 0 = queue all maxIterations sweeps instead of replaying one sweep from a
 *                      CUDA-graph `while` node until step() returns (default 1; row-partitioned graphs always
 *                      queue: their exchange barrier waits on peers)
 *  vrec_sg_batch_info  what = 0: start vertices the last vrec_sg_query served with the batch kernel;
 *                      1: batch path available for this graph; 2: vertices with in-edges; 3: edges
 *                      between them; 4: the same with the slice padding;
 *                      5 / 6 / 7: host-clock microseconds of the last batch call (uploads / kernel /
 *                      copy-back)                                                            */
int vrec_sg_set_option(vrec_sg *sg, const char *name, int32_t value);
int64_t vrec_sg_batch_info(vrec_sg *sg, int32_t what);

/* The full stationary vector of one vertex (the DataFrame step() returns, before the
 * `id != vertex and probability > 0` filter): out_x[vertex_count], in vertex_ids order.   */
int vrec_sg_stationary(vrec_sg *sg, int64_t vertex, double epsilon, int32_t max_iterations,
                       double *out_x, int32_t *out_iterations, int32_t *out_converged,
                       double *out_residual);

/* Device-resident power iteration for throughput measurement: runs `iterations` SpMV
 * passes x' = 0.15 u + 0.85 P^T x from x0 = 1/N for vertex index 0, no convergence stop,
 * asynchronously on vrec_stream().                                                          */
int vrec_sg_iterate_device(vrec_sg *sg, int32_t iterations);
int64_t vrec_sg_resident_bytes(vrec_sg *sg);

/* Synthetic graph built directly on the device (bench config "oversized graph"): n_vertices
 * vertices, out_degree edges per vertex, row-stochastic weights; ids are 0..n_vertices-1.
 * rank/world select the contiguous row range [lo, hi) of P^T this process owns.            */
int vrec_sg_generate(vrec_ctx *ctx, int64_t n_vertices, int32_t out_degree, uint64_t seed,
                     int32_t rank, int32_t world, vrec_sg **out);

/* ---------------------------------------------------------------- host-only helpers (tests) */

/* The vertex table and CSR of P^T exactly as vrec_sg_load builds them, on the host: out_ids
 * [<= 2*nnz], out_rowptr [<= 2*nnz+1], out_src / out_w [nnz].  No device is touched.        */
int vrec_host_sg_csr(int64_t nnz, const int64_t *source_id, const int64_t *target_id,
                     const double *balanced_weight, int64_t *out_n, int64_t *out_ids,
                     int32_t *out_rowptr, int32_t *out_src, double *out_w);
/* Row ranges of a row-partitioned graph for `world` ranks, from the CSR's rowptr [n_rows+1]:
 * out_bounds[world+1], rank k owns rows [out_bounds[k], out_bounds[k+1]) (balanced by in-edges). */
int vrec_host_sg_partition(int64_t n_rows, const int32_t *rowptr, int32_t world, int64_t *out_bounds);
/* Device CSR of P^T (rows owned by this process) copied back: out_rowptr [rows+1], out_src /
 * out_w [edge_count].                                                                        */
int vrec_sg_export_csr(vrec_sg *sg, int32_t *out_rowptr, int32_t *out_src, double *out_w);

/* Debug: runs one 128x128x128 fp16 tile through tcgen05.mma + TMEM and returns the largest
 * absolute deviation from a double-precision reference (validates the tensor-core plumbing).   */
int vrec_debug_tc_selftest(vrec_ctx *ctx, double *out_max_abs_err);

#ifdef __cplusplus
}
#endif
#endif /* VREC_H */
