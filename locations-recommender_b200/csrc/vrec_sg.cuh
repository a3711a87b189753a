// Types shared by the SG translation units (vrec_sg.cu, vrec_sg_batch.cu); not part of the ABI.
#pragma once
#include <memory>

#include "vrec_internal.cuh"

constexpr double kAlpha = 0.15;                 // stochastic/StochasticRecommender.scala:38
constexpr int SPMV_THREADS = 256;
constexpr int SPMV_WARPS = SPMV_THREADS / 32;
constexpr int64_t SRC_BLOCK = 6291456;          // canonical source block, 3 * 2^21 vertices = 48 MB of x

// Device-side loop control (one per query slot).
struct SgState {
    int done;          // 1 once the iteration stopped
    int iterations;    // the `iteration` of step()'s message (:94,:100)
    int converged;     // 1 = "Converged in ...", 0 = "... reached the maximum ..."
    unsigned int ticket;
    double residual;   // last sum of squared differences (:131-139)
    // the query's parameters and the loop counter live on the device: every sweep reads them here, so the same
    // launch can be replayed by a CUDA-graph `while` node until `done` (no host round trip per iteration)
    int cur;           // `iteration` of the sweep that runs next (step()'s argument, :92)
    int max_it;
    int check;         // 1: isConverged after every sweep; 0: fixed number of sweeps (vrec_sg_iterate_device)
    int pad;
    long long uidx;    // index of the start vertex (u = 1 there), -1: none
    double eps2;       // epsilon * epsilon (:40)
    unsigned long long flag_base;   // row-partitioned graphs: epoch << 32 (see SgExchange)
};

// Batch path (vrec_sg_batch.cu): many personalised queries whose start vertex has no in-edge
// (every person vertex of the reference's graphs).  See the comment at the top of that file.
struct SgBatch {
    bool analysed = false, ok = false;
    int n_a = 0;                              // active vertices (in-degree > 0)
    int n_fast_chunks = 0, n_slow = 0;        // slices of 32 rows summed one row per thread / rows behind them
    int64_t r_nnz = 0;                        // edges between active vertices
    std::vector<int> h_act_of;                // [N] active index or -1
    DevBuf<int> r_rowptr, r_src, full_len;    // reduced CSR of P^T over the active vertices
    DevBuf<double> r_w;
    DevBuf<int> chunk_r;                      // the same graph in slices of 32 rows (see vrec_sg_batch.cu)
    DevBuf<long long> chunk_ptr, act_id;
    DevBuf<unsigned short> sell_src;
    DevBuf<double> sell_w;
    DevBuf<unsigned> cand_mask;
    long long sell_nnz = 0;
    DevBuf<int> z_rowptr, z_row, z_pos;       // out-edges of the in-degree-0 vertices: (active row, position in its full row)
    DevBuf<double> z_w;
    DevBuf<double> x1a;                       // x after the first (start-vertex independent) iteration, active part
    bool x1_ready = false;
    double r1_base = 0.0;                     // residual of that iteration without the start vertex's own term
    DevBuf<double> scratch;
    DevBuf<int> counter, q_vertex, out_count, out_it, out_conv;
    DevBuf<long long> out_id;
    DevBuf<double> out_prob;
    int mode = 1;                             // 0 never, 1 auto (>= 4 eligible queries), 2 whenever eligible
    int force_t = 0;                          // debug: targets per CTA (0 = largest that fits)
    long long us_prepare = 0, us_kernel = 0, us_results = 0;   // host-clock split of the last batch call
    int64_t last_batched = 0;                 // queries the last vrec_sg_query served by the batch kernel
};

// ---- row-partitioned graphs: exchange of x' and of the residual partials over peer memory ----
constexpr int VREC_MAX_WORLD = 16;

// One per graph and rank, in device memory, written by every peer (NVLink P2P stores).
struct SgExchange {
    double res[2][VREC_MAX_WORLD];              // residual partial of rank r for iteration parity p
    unsigned long long flag[VREC_MAX_WORLD];    // flag[r] = last step rank r has completed: (epoch << 32) | step
};
// Where every rank keeps its two x buffers and its exchange block (device copy, one per graph).
struct SgPeers {
    double *x[2][VREC_MAX_WORLD];
    SgExchange *xchg[VREC_MAX_WORLD];
    int world, rank;
};

struct vrec_sg {
    vrec_ctx *ctx = nullptr;
    int64_t N = 0, nnz = 0;
    int64_t row_lo = 0, row_hi = 0;           // rows of P^T owned by this process (balanced by nnz)
    bool partitioned = false;
    std::vector<int64_t> h_ids;               // ascending vertex ids (host copy for lookups)
    DevBuf<long long> d_ids;
    // Source blocks (canonical order): graphs with more than 3 * 2^21 vertices are
    // swept once per block of 3 * 2^21 sources, so that the gathered part of x (<= 48 MB) stays L2-resident.
    // Every block is a CSR of its own (block-major layout): the in-edges of a row that come from the block's
    // sources, rows adjacent in src / w, so that a sweep streams one contiguous array pair.  (Sub-)rows longer
    // than VREC_CANON_SEG are summed segment-wise through the per-block tables.
    struct Block {
        int64_t nnz = 0;
        DevBuf<int> rowptr;                   // [rows+1]
        DevBuf<int> src;                      // source vertex index per in-edge (global index)
        DevBuf<double> w;
        int n_long = 0, n_seg = 0;
        DevBuf<int> long_rows;                // [n_long] ascending local row
        DevBuf<int> long_segptr;              // [n_long+1] offsets into partials
        DevBuf<int> seg_row;                  // [n_seg] index into long_rows
        DevBuf<double> partials;              // [n_seg]
    };
    int nblocks = 1;
    int flat_variant = 0;                     // tuning: (CTAs per SM, windows per batch) of the flat kernel, see sg_run_device
    bool rows_kernel = false;                 // debug / A-B: the round-1 half-warp-per-row kernel
    std::vector<std::unique_ptr<Block>> blocks;
    DevBuf<double> d_xbuf;                    // x[0] | x[1] | SgExchange, one allocation (one IPC handle)
    double *d_x[2] = {nullptr, nullptr};
    int64_t xstride = 0;                      // doubles between x[0] and x[1]
    SgExchange *d_xchg = nullptr;
    DevBuf<SgPeers> d_peers;                  // device copy of `peers`
    SgPeers peers;
    bool peers_ready = false, peers_local = false;
    std::vector<void *> ipc_opened;           // peer mappings to close
    unsigned long long epoch = 0;             // query counter of a partitioned graph (same on every rank)
    DevBuf<double> d_block_partials;
    DevBuf<SgState> d_state;
    int grid = 0;
    // the step() loop as a CUDA graph: one conditional `while` node whose body is one sweep (all source blocks)
    // plus the kernel that sets the loop condition from SgState::done.  Built at the first query.
    void *loop_graph = nullptr, *loop_exec = nullptr;     // cudaGraph_t / cudaGraphExec_t
    int loop_launches = 0;                                // kernel launches of one pass of the body
    int use_graph = 1;                                    // option "graph": 0 = queue all maxIterations sweeps
    bool graph_pending = false;                           // a graph launch whose pass count is not yet in ctx->launches
    // top-N scratch
    DevBuf<long long> d_filter_ids;
    DevBuf<double> d_cand_val;
    DevBuf<long long> d_cand_key;
    DevBuf<long long> d_out_key;
    DevBuf<double> d_out_val;
    DevBuf<int> d_out_count;
    SgBatch batch;
    ~vrec_sg();
};

int sg_run_device(vrec_sg *g, long long uidx, double epsilon, int max_it, bool check_convergence);
int sg_fetch_state(vrec_sg *g, int max_it, SgState *h, int *result_buf);
int64_t sg_lookup(const vrec_sg *g, int64_t id);
// vrec_sg_batch.cu
int sg_batch_analyse(vrec_sg *g, const std::vector<int> &rowptr, const int *h_src, const double *h_w);
int sg_batch_prepare(vrec_sg *g);
int sg_batch_query(vrec_sg *g, const std::vector<int> &qidx, const std::vector<int> &qvertex,
                   double epsilon, int max_it, const int64_t *place_filter, int64_t n_filter, int max_recs,
                   int64_t *out_id, double *out_prob, int32_t *out_count, int32_t *out_iterations,
                   int32_t *out_converged);
