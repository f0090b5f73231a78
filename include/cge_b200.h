/*
 * cge_b200.h -- C ABI of libcge_b200.so: the B200-native scoring path of CGE.jl.
 *
 * The reference has no FFI for this path; the seam is the Julia call
 *     wGCL(edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
 *          init_edges, init_eweights, init_embed, split, seed, auc_samples, verbose)
 * (/root/reference/src/divergence.jl:27-31) and wGCL_directed (divergence.jl:282-286), made from
 * example/CGE_CLI.jl:18-24.  Each entry point below is what a `ccall` from a Julia wrapper of
 * those two functions binds (see INTEGRATION.md and julia/CGEB200.jl); plain pointers and
 * sizes only, the caller owns every buffer, the library only reads inputs during the call.
 *
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails
 * with CGE_B200_ERR_CUDA.
 */
#ifndef CGE_B200_H
#define CGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGE_B200_VERSION_MAJOR 0
#define CGE_B200_VERSION_MINOR 1
#define CGE_B200_VERSION_PATCH 0

/* alpha grid 0.25:0.25:10.0 (divergence.jl:36-37,139) */
#define CGE_B200_N_ALPHA 40

/* status codes (0 = success).  The Julia/Python wrappers turn ERR_ASSERT_* into the same
 * AssertionError messages the reference raises (divergence.jl:50,81,303,363). */
enum {
    CGE_B200_OK = 0,
    CGE_B200_ERR_ARG = -1,            /* NULL pointer / bad size / bad struct_size */
    CGE_B200_ERR_ASSERT_COMM = -2,    /* "No. communities not matching no. vertices" */
    CGE_B200_ERR_ASSERT_DIST = -3,    /* "Distances vector length is not equal to no. vertices" */
    CGE_B200_ERR_OOM = -4,            /* device or host allocation failed */
    CGE_B200_ERR_CUDA = -5,           /* no device, launch or runtime failure */
    CGE_B200_ERR_NCCL = -6,
    CGE_B200_ERR_STATE = -7           /* run() before upload(), comm not initialised, ... */
};

/* how the alpha loop is driven on the device */
enum {
    CGE_B200_DRIVER_AUTO = 0,
    CGE_B200_DRIVER_HOSTLOOP = 1,     /* one launch per pass, host reads the residual */
    CGE_B200_DRIVER_PERSISTENT = 2,   /* one cooperative launch per alpha runs all its passes */
    CGE_B200_DRIVER_RING = 3          /* same, matrix streamed by cp.async.bulk through a smem ring */
};

/* where the pair matrix lives */
enum {
    CGE_B200_REGIME_AUTO = 0,          /* stored when it fits in free HBM, else recompute */
    CGE_B200_REGIME_STORED = 1,        /* q = (1-D)^(1/4) kept in HBM: 8 B per pair and pass, HBM bound */
    CGE_B200_REGIME_RECOMPUTE = 2,     /* distances re-derived from the embedding every pass (FP64 bound),
                                          from d^2 = n_i + n_j - 2 x_i.x_j on the centred embedding: one
                                          FMA per dimension and pair; pairs under cancellation use the
                                          difference form; extrema and sampled pairs share the arithmetic */
    CGE_B200_REGIME_RECOMPUTE_DOT = 3, /* the same (the name of round 1, when it was opt-in); reported back
                                          as 3 when asked for as 3 */
    CGE_B200_REGIME_RECOMPUTE_DIFF = 4 /* recompute with the reference's difference form sum (x_i - x_j)^2
                                          (auxilary.jl:14-20): two FP64 instructions per dimension and
                                          pair; the cross-check of the default.  CGE_B200_RC_FORM=diff
                                          selects it whenever the regime resolves to recompute */
};

/*
 * One scoring problem = the argument list of wGCL / wGCL_directed.
 * Index arrays hold `index_base`-based ids (1 when passed straight from Julia).
 * Matrices are addressed as base[row*row_stride + col*col_stride] (in elements), so Julia's
 * column-major Matrix{Float64} (row_stride 1, col_stride = size(A,1)) and C row-major
 * arrays are both accepted without a copy.
 */
typedef struct cge_b200_problem {
    int32_t struct_size;          /* sizeof(cge_b200_problem), ABI check */
    int32_t index_base;           /* 0 or 1 */
    int32_t directed;             /* 0: wGCL, 1: wGCL_directed */
    int32_t split;                /* --split-global (divergence.jl:235-241) */

    /* scored graph (the landmark graph in landmark mode) */
    int64_t m;                    /* size(edges,1) */
    const int64_t *edge_src;      /* edges[:,1] */
    const int64_t *edge_dst;      /* edges[:,2] */
    const double *eweights;       /* m */
    int64_t n_comm;               /* size(comm,1) */
    const int64_t *comm;          /* comm[:,1], index_base-based community ids */
    const double *embed;          /* size(embed) = embed_rows x d */
    int64_t embed_rows, d, embed_row_stride, embed_col_stride;
    int64_t n_distances;          /* length(distances) */
    const double *distances;      /* diagonal of the distance matrix */
    const double *vweights;       /* vertex (landmark) weights, >= n entries */

    /* landmark mode iff n_full > 0 (= length(v_to_l), divergence.jl:44) */
    int64_t n_full;
    const double *init_vweights;  /* n_full */
    const int64_t *v_to_l;        /* n_full, index_base-based landmark id per vertex */
    const double *init_embed;     /* n_full x d */
    int64_t init_row_stride, init_col_stride;

    /* sampled pairs for the local score, drawn by the host wrapper exactly where the reference
     * draws them (divergence.jl:184-210, 484-513).  n_sets = 1 when seeded (the same sets for
     * every alpha), CGE_B200_N_ALPHA when seed = -1.  Arrays are n_sets x n_samples, set-major;
     * ids refer to the ORIGINAL graph.  n_samples = 0 skips the local score. */
    int64_t n_samples, n_sets;
    const int64_t *pos_i, *pos_j; /* sampled edges */
    const double *pos_w;          /* their weights (auc_weights) */
    const int64_t *neg_i, *neg_j; /* sampled non-edges */

    /* tuning; 0 = default */
    int32_t max_alphas;           /* evaluate only the first max_alphas grid points (<= 40) */
    int32_t driver;               /* CGE_B200_DRIVER_* */
    int32_t regime;               /* CGE_B200_REGIME_* */
    int32_t reserved;
} cge_b200_problem;

/* filled by run()/score(); everything a roofline computation or a parity test needs */
typedef struct cge_b200_stats {
    int32_t struct_size;
    int32_t n_alpha_run;                    /* alphas for which the fixed point ran */
    int32_t iters[CGE_B200_N_ALPHA];        /* fixed-point passes per alpha */
    double div[CGE_B200_N_ALPHA];           /* global score per alpha (NaN when skipped) */
    double auc[CGE_B200_N_ALPHA];           /* local score per alpha (NaN when skipped) */
    double lo, hi;                          /* extrema of the raw distances (divergence.jl:92) */
    double hi_full;                         /* landmark mode: max full-graph distance (:113) */
    int64_t n, n_pairs;                     /* scored vertices; n(n+1)/2 */
    int64_t fp_sweeps, b_sweeps;            /* passes over the pair matrix: fixed point / B */
    int64_t matrix_bytes;                   /* bytes of the stored q matrix */
    int64_t launches;                       /* kernels launched by this library in the call */
    int32_t n_tiles, grid, driver, n_ranks;
    int32_t regime;                         /* regime actually used */
    int32_t diam_candidate_tiles;           /* landmark mode: tiles re-checked in FP64 after the
                                               tensor-core diameter filter; -1 = filter not used */
    float ms_upload;                        /* H2D + host preprocessing */
    float ms_build;                         /* distance tiles + normalisation (CUDA events) */
    float ms_solve;                         /* the alpha loop (CUDA events) */
    float ms_total;                         /* wall clock of the call */
    float ms_sweeps;                        /* sum of the fixed-point sweep kernels (CUDA events) */
    float ms_bsweeps;                       /* sum of the stand-alone B sweep kernels (CUDA events) */
    int32_t b_fused;                        /* of b_sweeps: B sweeps that rode on the first fixed-point
                                               pass of the next alpha (one matrix read for both; that
                                               pass is counted in fp_sweeps, its kernel time in ms_sweeps) */
    float ms_fused;                         /* of ms_sweeps: the fused B + first-pass kernels (k_bfp) */
} cge_b200_stats;

typedef struct cge_b200_handle cge_b200_handle;

/* library / device discovery */
void cge_b200_version(int *major, int *minor, int *patch);
int cge_b200_device_count(void);             /* 0 when no CUDA device is usable */
const char *cge_b200_last_error(void);       /* thread-local message of the last failure */

/* One-shot drop-in for a single wGCL / wGCL_directed call on device 0.
 * out has room for 7 doubles; *out_len is 7, or 6 for the directed star-graph early exit
 * (divergence.jl:332-334).  stats may be NULL. */
int cge_b200_score(const cge_b200_problem *p, double *out, int32_t *out_len,
                   cge_b200_stats *stats);

/* The same call sharded over n_gpus (2..8) GPUs of this box from ONE process: one host thread per
 * GPU, tiles split by cge_b200_shard_plan, per-pass exchange over NVLink peer memory inside the
 * kernel.  Exact mode only (landmark-mode problems stay on one GPU).  cge_b200_score itself does
 * this when the environment variable CGE_B200_GPUS=N is set, so hosts need no code change. */
int cge_b200_score_multi(const cge_b200_problem *p, int n_gpus, double *out, int32_t *out_len,
                         cge_b200_stats *stats);

/* Handle API: keeps device buffers between calls and separates the host->device stage from
 * the device-resident run (bench.py times them separately). */
int cge_b200_create(int device, cge_b200_handle **out);
void cge_b200_destroy(cge_b200_handle *h);
int cge_b200_upload(cge_b200_handle *h, const cge_b200_problem *p);
int cge_b200_run(cge_b200_handle *h, double *out, int32_t *out_len, cge_b200_stats *stats);

/* Multi-GPU exact mode, one rank per process: every rank uploads the same problem, owns a
 * contiguous range of pair-matrix tiles and all-reduces the n-length partial degree sums
 * once per pass (NCCL).  id_bytes = cge_b200_comm_id_size() bytes obtained on rank 0 from
 * cge_b200_comm_unique_id() and broadcast by the caller (e.g. torch.distributed). */
int cge_b200_comm_id_size(void);
int cge_b200_comm_unique_id(void *id_bytes);
int cge_b200_comm_init(cge_b200_handle *h, const void *id_bytes, int rank, int n_ranks);

/* Optional NVLink peer-memory exchange for the multi-GPU persistent driver: instead of one NCCL
 * all-reduce per pass issued by the host, the fixed-point kernel itself stores its partial degree
 * sums into every peer's exchange buffer as self-flagged 8-byte records (value half + pass
 * number) that the peers poll, so a whole alpha runs in one launch on every rank.  After comm_init: every rank calls p2p_export (allocates
 * its buffer for up to max_vertices vertices, returns a 64-byte CUDA IPC handle), the caller
 * all-gathers the handles, every rank calls p2p_import with the n_ranks x 64 bytes.  Without it
 * multi-rank runs fall back to the NCCL host loop. */
int cge_b200_p2p_handle_size(void);
int cge_b200_p2p_export(cge_b200_handle *h, int64_t max_vertices, void *handle_out);
int cge_b200_p2p_import(cge_b200_handle *h, const void *all_handles);

/* Measured FP64 FMA throughput of the handle's device in TFLOP/s (2 flop per FMA): the roofline
 * denominator of the recompute regime, which MEASURED_PEAKS.json does not provide. */
int cge_b200_measure_fp64_peak(cge_b200_handle *h, double *tflops);

/* The same for every way this device can issue FP64 multiply-adds (cge_microbench.cu): out[0] DFMA,
 * out[1] mma.sync m8n8k4 (DMMA), out[2] m16n8k16, out[3]/out[4] DMMA and DFMA shares of a 4:1 mix
 * (do the two share a pipe?), out[5]/out[6] the same for m16n8k16 with an 8:1 mix, out[7] m16n8k8;
 * all in TFLOP/s.  Design evidence for the Gram step of the recompute regime. */
int cge_b200_measure_fp64_pipes(cge_b200_handle *h, double out[8]);

/* SURVEY.md 8(f) F1 -- the negative pairs of the local score drawn on the device.  Replaces the
 * reference's NE construction (all pairs minus the edge Set, divergence.jl:121-137 / 405-421: n^2/2
 * tuples, infeasible above a few 10^4 vertices) and its sample(NE, auc_samples, replace=true)
 * (:193-194, 209 / 495-496, 513): the edges go into a hash set in HBM and every sample is an
 * independent uniform draw from the non-edges (i < j when undirected, ordered pairs i != j when
 * directed), n_sets x n_samples of them, set-major, ids index_base-based like the edges.
 * Deterministic in (seed, n, edges); a counter-based generator replaces Julia's stream, so the
 * sets are identically distributed, not identical -- hosts keep drawing with Julia wherever NE
 * fits.  A graph without non-edges fails like sample() on an empty collection.
 * draws_per_sample (may be NULL) returns the mean number of candidate pairs drawn per sample. */
int cge_b200_sample_non_edges(cge_b200_handle *h, int64_t n, int64_t m, const int64_t *edge_src,
                              const int64_t *edge_dst, int32_t index_base, int32_t directed,
                              int64_t n_samples, int64_t n_sets, uint64_t seed, int64_t *out_i,
                              int64_t *out_j, double *draws_per_sample);

/* SURVEY.md 8(f) F2 -- the aggregation half of landmarks() (landmarks.jl:387-463) on the device.  Landmark
 * selection (runsplit, landmarks.jl:155-345) stays on the host; given its assignment landmark[i] (index_base-
 * based, values < n_landmarks + index_base) this call produces every landmark-mode input of wGCL:
 *   out_embed    n_landmarks x d, row-major: weighted centroids (:387-404)
 *   out_lweight  n_landmarks: summed vertex weights
 *   out_dii      n_landmarks: sqrt(sum of UNWEIGHTED squared deviations / lweight) (:407-423)
 *   out_cluster  n_landmarks: community of the landmark's last member, as passed in comm (:426-430)
 *   out_edge_*   the weighted landmark edge list in idx order (row-major; upper triangle with self loops
 *                when undirected), cells with weight > 0 only (:433-463); capacity edge_cap entries
 *                (min(m, n_landmarks^2) always suffices), *out_n_edges = entries written.
 * Every sum runs in the reference's order (members in ascending vertex order, edges in file order, products
 * and sums rounded separately), so the outputs are bit-identical to the Julia loops.  comm may be NULL when
 * out_cluster is NULL.  embed is addressed as embed[i*row_stride + j*col_stride]. */
int cge_b200_landmarks_aggregate(cge_b200_handle *h, int64_t n, int64_t d, int64_t n_landmarks,
                                 const int64_t *landmark, int32_t index_base, const double *vweights,
                                 const int64_t *comm, const double *embed, int64_t embed_row_stride,
                                 int64_t embed_col_stride, int64_t m, const int64_t *edge_src,
                                 const int64_t *edge_dst, const double *eweights, int32_t directed,
                                 double *out_embed, double *out_lweight, double *out_dii,
                                 int64_t *out_cluster, int64_t *out_edge_src, int64_t *out_edge_dst,
                                 double *out_eweights, int64_t edge_cap, int64_t *out_n_edges);

/* SURVEY.md 8(f) F4 -- landmark SELECTION on the device.  Replaces runsplit(embedding, w, initial_clusters,
 * n, s, rule) (landmarks.jl:279-345, called at :379) with the split rules split_cluster_rss (:155-210),
 * split_cluster_size (:218-238) and split_cluster_diameter (:247-267): the embedding, the weights and the
 * member order of every cluster stay in HBM; each cut runs its O(s d^2) / O(s d) passes (weighted moments and
 * RSS, covariance, projection on the principal axis, sort, range moments, stable regrouping) as kernels and
 * leaves the queue (landmarks.jl:12-46), the principal axis of the d x d covariance and the rule's
 * comparisons to the host.
 *   clusters      CSR: cluster c holds cluster_members[cluster_ptr[c] .. cluster_ptr[c+1]) (index_base-based
 *                 vertex ids), clusters in the order of the reference's sort(initial_clusters)
 *   land, forced  number of landmarks wanted / pieces every initial cluster is first cut into (n, s)
 *   rule          CGE_B200_RULE_RSS (0), _SIZE (2), _DIAMETER (3); split_cluster_rss2 ("retained for testing
 *                 purposes", :83-147) is not offered: CGE_B200_ERR_ARG
 *   out_group     n entries: 0-based landmark id per vertex (runsplit's return value)
 *   out_cuts      optional: number of cluster cuts performed
 * The reference's ErrorExceptions ("Trying to split homogenous cluster", "Unexpected empty cluster generated")
 * come back as CGE_B200_ERR_STATE with that text.
 *   eig, eig_user optional callback for the one step the reference gives to LAPACK: the principal axis of a
 *                 cluster's d x d weighted covariance (eigvecs(...)[:, end], landmarks.jl:160-162).  c is the
 *                 full symmetric matrix, row-major; the callback writes the eigenvector of the largest
 *                 eigenvalue to v (any length) and returns 0.  A cut depends on the SIGN LAPACK happens to
 *                 return (which child is "low", where the median of an odd cluster goes), so a host that
 *                 wants the reference's very landmarks passes its own LAPACK here (Julia: eigvecs, Python:
 *                 numpy.linalg.eigh).  NULL: the library's own solver (Householder tridiagonalisation,
 *                 bisection, inverse iteration), sign fixed to "largest-magnitude component positive". */
#define CGE_B200_RULE_RSS 0
#define CGE_B200_RULE_SIZE 2
#define CGE_B200_RULE_DIAMETER 3
typedef int (*cge_b200_eigvec_fn)(const double *c, int64_t d, double *v, void *user);
int cge_b200_landmarks_select(cge_b200_handle *h, int64_t n, int64_t d, const double *embed,
                              int64_t embed_row_stride, int64_t embed_col_stride, const double *vweights,
                              int64_t n_clusters, const int64_t *cluster_ptr, const int64_t *cluster_members,
                              int32_t index_base, int64_t land, int64_t forced, int32_t rule,
                              cge_b200_eigvec_fn eig, void *eig_user, int64_t *out_group, int64_t *out_cuts);

/* landmarks.jl:369 -- size(unique(embedding, dims=1), 1), the row count landmarks() caps `land` with
 * (:371-374), on the device: rows are hashed, sorted by hash, and neighbours with equal hashes compared in
 * full (bit patterns: isequal semantics, like Julia's unique).  5 s of hashing on the host at 10^6 x 128. */
int cge_b200_unique_rows(cge_b200_handle *h, int64_t n, int64_t d, const double *embed,
                         int64_t embed_row_stride, int64_t embed_col_stride, int64_t *out_count);

/* Host-only helper of the above, exported for its unit test: unit-length eigenvector of the largest
 * eigenvalue of the symmetric d x d matrix a (row-major, upper triangle read), largest-magnitude component
 * positive; *lambda (optional) receives the eigenvalue. */
int cge_b200_sym_top_eigvec(const double *a, int64_t d, double *v_out, double *lambda);

/* SURVEY.md 8(f) F3 -- ingest of the whitespace-delimited numeric tables the reference reads with
 * readdlm(fn, Float64 | Int) (auxilary.jl:86 edgelist, :123 communities, :150-155 embedding), parsed
 * on all host cores straight into the caller's matrix.  Host-only, no GPU needed.
 *   table_dims: rows = non-blank lines after skip_rows lines, cols = cells of the first row.
 *   read_table: fills out[r*row_stride + c*col_stride] (elements) for a rows x cols table; fails
 *   with CGE_B200_ERR_ARG when a row has another column count or a cell is not a number in full --
 *   the failure the reference uses to detect a node2vec header line (retry with skip_rows = 1).
 * n_threads = 0 uses every hardware thread. */
int cge_b200_table_dims(const char *path, int64_t skip_rows, int32_t n_threads, int64_t *rows,
                        int64_t *cols);
int cge_b200_read_table(const char *path, int64_t skip_rows, int32_t n_threads, int64_t rows,
                        int64_t cols, int64_t row_stride, int64_t col_stride, double *out);

/* Self-test of the recompute regime's arithmetic: its epilogue evaluates the square roots and the
 * normalisation 1 - (D - lo)/(hi - lo) with short branch-free instruction sequences (so that 8 pairs
 * interleave in the FP64 pipe) instead of the correctly rounded sqrt and divide the stored regime
 * and the reference use (Julia sqrt and /, divergence.jl:92, auxilary.jl:19).  Runs n_samples
 * pseudo-random operands in the epilogue's ranges through both and counts the results farther
 * than 2 ulp from the correctly rounded ones (square roots; normalisations, which must also be
 * exactly 0 at D = hi). */
int cge_b200_selftest_math(cge_b200_handle *h, int64_t n_samples, uint64_t seed,
                           int64_t *sqrt_mismatches, int64_t *div_mismatches);

/* Host-only: the tile range [*tile_begin, *tile_end) of the upper-triangular tile sequence
 * that `rank` of `n_ranks` owns for an n-vertex problem, and the tile count. No GPU needed. */
int cge_b200_shard_plan(int64_t n, int rank, int n_ranks, int64_t *n_tiles,
                        int64_t *tile_begin, int64_t *tile_end);

/* Debug / parity probes (used by tests): copy device state of the last upload()/run() back.
 * what: 0 = dense n x n row-major matrix of q = (1 - D)^(1/4) in the caller's vertex order,
 *       1 = T (or Tin) in the caller's order, 2 = Tout, 3 = last S (or Sin), 4 = last Sout.
 * n_elems is the capacity of buf in doubles. */
int cge_b200_debug_read(cge_b200_handle *h, int what, double *buf, int64_t n_elems);

#ifdef __cplusplus
}
#endif
#endif /* CGE_B200_H */
