/*
 * annembed_cuda.h -- C ABI of the B200-native (sm_100a) cross-entropy embedding optimizer.
 *
 * Drop-in boundary for ONE hot path of jean-pierreBoth/annembed: what sits behind
 *   Embedder::new(&kgraph, EmbedderParams) / embed() / get_embedded_reindexed()
 * i.e. to_proba_edges (src/tools/kdumap.rs:26-235) + entropy_optimize (src/embedder.rs:794-904).
 * The reference is pure Rust with no FFI of its own for this path; these are the entry points a
 * `cuda` cargo feature would bind with `extern "C"` (INTEGRATION.md shows the Rust side).
 * Citations below are relative to /root/reference/src.
 *
 * Conventions
 *  - every function returns an int status (ANNEMBED_OK == 0); nothing aborts or throws across the ABI
 *    (the reference exits the process on an empty neighbourhood, kdumap.rs:75-85; here it is a status);
 *  - all pointer arguments are HOST pointers owned by the caller and are copied before return;
 *  - the opaque context owns every device allocation, its stream and (optionally) its NCCL communicator;
 *  - a context is not thread-safe; distinct contexts are independent; calls block until done;
 *  - there is no CPU fallback: without a CUDA device annembed_cuda_create fails with ANNEMBED_ERR_CUDA.
 *  - one context drives one GPU; multi-GPU = one context per process/GPU + annembed_cuda_comm_init.
 */
#ifndef ANNEMBED_CUDA_H
#define ANNEMBED_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANNEMBED_CUDA_ABI_VERSION 1

enum annembed_status {
    ANNEMBED_OK = 0,
    ANNEMBED_ERR_INVALID_ARG = 1,   /* null pointer, bad size, col >= n, self edge (embedder.rs:1201 asserts) */
    ANNEMBED_ERR_CUDA = 2,          /* CUDA runtime error / no device */
    ANNEMBED_ERR_EMPTY_ROW = 3,     /* node without neighbour: kdumap.rs:75-85 (process::exit(1) there) */
    ANNEMBED_ERR_UNSORTED_ROW = 4,  /* row distances not ascending: kgraph.rs:508-509 invariant */
    ANNEMBED_ERR_STATE = 5,         /* call order: graph / weights / embedding not set (embedder.rs:802-808) */
    ANNEMBED_ERR_UNSUPPORTED = 6,   /* asked_dim > 32, E >= 2^32, n >= 2^32 */
    ANNEMBED_ERR_COMM = 7,          /* NCCL not loadable / communicator failure */
    ANNEMBED_ERR_NO_NEGATIVE = 8    /* n too small for 5 accepted negatives (embedder.rs:1241-1252 would spin) */
};

/* Mirror of EmbedderParams (embedparams.rs:76-103; defaults :107-132) + device-side knobs.
 * dmap_init / grad_factor / hierarchy_layer are carried for fidelity; the initial layout is an input
 * (annembed_cuda_set_embedding), so they do not change what this library computes. */
typedef struct annembed_cuda_params {
    uint32_t asked_dim;            /* embedparams.rs:78  default 2   (1..32 supported on device) */
    uint32_t dmap_init;            /* :80  default 1 */
    double   beta;                 /* :82  default 1.   exponent inside the exponential, kdumap.rs:172-174 */
    double   b;                    /* :85  default 1.   Cauchy exponent, embedder.rs:1216-1222 */
    double   scale_rho;            /* :88  default 1. */
    double   grad_step;            /* :90  default 2. */
    uint32_t nb_sampling_by_edge;  /* :92  default 10 */
    uint32_t nb_grad_batch;        /* :94  default 20 */
    uint32_t grad_factor;          /* :98  default 4 */
    uint32_t hierarchy_layer;      /* :100 default 0 */
    uint32_t hubness_weighting;    /* :102 default 0: negatives uniform; 1: alias over set_neg_weights */
    /* ---- device-side additions (no reference counterpart: its RNG is unseeded, embedder.rs:1182) ---- */
    uint32_t mini_epochs_per_batch;/* sweeps over the nodes (asynchronous form) / bulk-synchronous mini-epochs per
                                      reference batch; 0 -> the default schedule, see eff_mini_epochs /
                                      mini_epochs_of_batch in annembed_cuda.cu */
    uint64_t seed;                 /* Philox4x32-10 key */
    uint32_t flags;                /* ANNEMBED_FLAG_* */
    uint32_t cell_substeps;        /* cell-resident epoch kernel: mini-epochs per launch (partners in other cells and the
                                      negatives are read from the layout as of the start of the launch);
                                      0 -> chosen from the fraction of edges that cross cells */
} annembed_cuda_params;

#define ANNEMBED_FLAG_NONE 0u
#define ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL 1u   /* force the thread-per-node epoch kernel (cross-check of the tiled one) */
#define ANNEMBED_FLAG_NO_L2_PERSIST 2u
#define ANNEMBED_FLAG_NO_RELABEL 4u             /* keep the caller's node order inside the optimizer (no locality relabelling) */
#define ANNEMBED_FLAG_REPLAY_IN_EDGES 8u        /* single rank: replay the sources' decisions in the in-edge kernel instead of
                                                   consuming the firing counts pushed by the out-edge kernel (cross-check) */
#define ANNEMBED_FLAG_LEGACY_EPOCH_KERNELS 16u  /* one launch of k_epoch_out + k_epoch_in per mini-epoch instead of the
                                                   cell-resident kernel (cross-check; also what dimensions > 4 use);
                                                   implies the bulk-synchronous form */
#define ANNEMBED_FLAG_NODE_ALIAS 64u            /* hubness sampler: node-level alias table for every draw (cross-check of the
                                                   sector-level table) */
#define ANNEMBED_FLAG_CP_ASYNC_PIPELINE 128u    /* asynchronous form, dimension <= 4, uniform negatives: k_sweep_events_cp (visits
                                                   pipelined through cp.async groups in shared memory) instead of the
                                                   register-staged k_sweep_events; same speed, kept for the A/B.
                                                   The hubness sampler always runs k_sweep_events_cp in these dimensions. */
#define ANNEMBED_FLAG_SECTOR_NEGATIVES 256u     /* asynchronous event kernels, uniform sampler: negatives shared by the 4 nodes of a
                                                   32-byte sector of the layout instead of the nodes of a 128-byte line (A/B) */
#define ANNEMBED_FLAG_BULK_SYNCHRONOUS 32u      /* one rank: the deterministic snapshot kernels (bit-reproducible for a seed)
                                                   instead of the asynchronous sweep (async_sweep.cuh), which like the
                                                   reference's threaded loop depends on the interleaving of the threads */

typedef struct annembed_cuda_stats {
    double   edge_weights_ms;      /* K0+K1 device time, last call */
    double   build_ms;             /* device context build (scales, transposed index), last call */
    double   optimize_ms;          /* device time of the last optimize loop (events on the library stream) */
    double   epoch_kernel_ms;      /* sum of epoch-kernel (K4) durations inside optimize_ms */
    double   exchange_ms;          /* all-gather time inside optimize_ms (0 on one GPU) */
    double   cross_entropy_ms;     /* K5 time, last call */
    uint64_t epoch_launches;       /* K4 launches in the last optimize */
    uint64_t kernel_launches;      /* all kernels of this library launched since create / reset_stats */
    uint64_t positive_samples;     /* positive-edge samples applied by this rank in the last optimize */
    uint64_t edge_updates;         /* = 6 * positive_samples (1 attraction + 5 repulsions, embedder.rs:1241) */
    uint64_t h2d_bytes;            /* since create / reset_stats */
    uint64_t d2h_bytes;
    double   model_bytes;          /* positive_samples * (12 + 36 d): SURVEY.md 8(d) algorithmic bytes */
    uint64_t mini_epochs_per_batch;/* the value in effect (resolved default) */
    uint64_t l2_persist_max_bytes; /* cudaDevAttrMaxPersistingL2CacheSize */
    uint64_t l2_window_max_bytes;  /* cudaDevAttrMaxAccessPolicyWindowSize */
    uint64_t n_cells;              /* cells of the internal numbering (0 before the first optimize / cross_entropy) */
    uint64_t cell_nodes;           /* largest cell size */
    uint64_t cell_substeps;        /* mini-epochs per launch of the cell kernel in the last optimize (0: not used) */
    uint64_t cross_cell_edges;     /* edges of the graph whose ends lie in different cells */
    uint64_t cross_rank_edges;     /* edges of the graph whose ends are owned by different ranks (0 on one rank) */
    uint64_t exchanges;            /* row exchanges between the ranks in the last optimize (0 on one rank) */
} annembed_cuda_stats;

typedef struct annembed_cuda_ctx annembed_cuda_ctx;

/* Fill `p` with EmbedderParams::default() (embedparams.rs:107-132) and the device-side defaults. */
int annembed_cuda_default_params(annembed_cuda_params *p);

/* ≙ Embedder::new (embedder.rs:107): copies params; `device` is the CUDA ordinal. */
int annembed_cuda_create(annembed_cuda_ctx **ctx, const annembed_cuda_params *params, int device);
int annembed_cuda_destroy(annembed_cuda_ctx *ctx);
/* NUL-terminated message of the last failing call on ctx (ctx == NULL: last create failure). */
const char *annembed_cuda_last_error(const annembed_cuda_ctx *ctx);

/* ---- multi-GPU (one context per rank). No reference counterpart (single process, rayon). ----
 * Every rank owns (fires, and holds the authoritative rows of) a part of the nodes; graph and n x d layout are replicated.
 * unique_id is the 128-byte ncclUniqueId made by rank 0. */
int annembed_cuda_comm_unique_id(uint8_t unique_id[128]);
int annembed_cuda_comm_init(annembed_cuda_ctx *ctx, int rank, int nranks, const uint8_t unique_id[128]);

/* Peer memory (optional; after set_graph_csr on every rank): opens every rank's layout buffers on every rank (CUDA IPC).
 * With it the asynchronous form runs on several ranks: moves of nodes owned elsewhere are reduced straight into the
 * owner's replica over NVLink, the owners' rows are all-gathered every few sweeps under the next launch.
 * (With ANNEMBED_FLAG_BULK_SYNCHRONOUS: the snapshot kernels store every owned row into all replicas, a 4-byte
 * all-reduce closes the mini-epoch.)  Without it: bulk-synchronous form, NCCL broadcast of the owned rows per mini-epoch.
 * export: 2 x 64-byte cudaIpcMemHandle_t of this rank's two layout buffers; import: the handles of all ranks, rank-major
 * (nranks x 128 bytes). */
int annembed_cuda_comm_export_layout(annembed_cuda_ctx *ctx, uint8_t handles[128]);
int annembed_cuda_comm_import_layouts(annembed_cuda_ctx *ctx, const uint8_t *all_handles);

/* ≙ the KGraph hand-off, kgraph.rs:108-120 + get_neighbours :157.  Rows sorted ascending by distance
 * (kgraph.rs:508-509), no self edges, every row non-empty.  row_ptr has n+1 entries.
 * After comm_init with several ranks the call is COLLECTIVE: every rank passes the same graph, uploads one nranks-th of
 * each array from its host and receives the other parts over NVLink (grouped ncclBroadcast). */
int annembed_cuda_set_graph_csr(annembed_cuda_ctx *ctx, uint64_t n, const uint64_t *row_ptr,
                                const uint32_t *col, const float *dist);

/* K1 ≙ to_proba_edges (kdumap.rs:26-116) / get_scale_from_proba_normalisation (:132-235).
 * Outputs are optional (NULL to keep results on the device only). */
int annembed_cuda_edge_weights(annembed_cuda_ctx *ctx, float *scale_out /*[n]*/, float *proba_out /*[E]*/);
/* K1b ≙ get_scale_from_umap (embedder.rs:760-783) + dichotomy_solver (tools/dichotomy.rs:4-65); dead code in
 * the reference.  Per row: bisection for beta' with sum_m exp(-(d_m-d_0) beta') = norm.  Outputs 1/beta' and
 * the un-normalised weights; does not replace the weights used by optimize. status_out[i]: 0 ok, 1 not converged. */
int annembed_cuda_edge_weights_umap(annembed_cuda_ctx *ctx, float norm, float *scale_out /*[n]*/,
                                    float *weight_out /*[E]*/, uint8_t *status_out /*[n], nullable*/);
/* Inject weights computed elsewhere (isolates the optimizer in tests). */
int annembed_cuda_set_edge_weights(annembed_cuda_ctx *ctx, const float *scale /*[n]*/, const float *proba /*[E]*/);
/* ≙ NodeParam::get_perplexity (nodeparam.rs:88-91) for every node. */
int annembed_cuda_get_perplexity(annembed_cuda_ctx *ctx, float *out /*[n]*/);

/* ≙ NodeSampler::new (embedder.rs:909-931): weights already clamp(in_degree,1,n) (embedder.rs:826-833).
 * NULL -> uniform negatives (embedder.rs:1121). */
int annembed_cuda_set_neg_weights(annembed_cuda_ctx *ctx, const float *w /*[n], nullable*/);
/* In-degree counts of the loaded graph ≙ Hubness::get_counts (fromhnsw/hubness.rs:39-79). */
int annembed_cuda_get_hubness_counts(annembed_cuda_ctx *ctx, uint32_t *counts /*[n]*/);

/* The initial layout, n x asked_dim row-major (≙ initial_embedding argument of entropy_optimize,
 * embedder.rs:794-798).  A device-resident copy is kept for annembed_cuda_reset_embedding. */
int annembed_cuda_set_embedding(annembed_cuda_ctx *ctx, const float *y);
int annembed_cuda_reset_embedding(annembed_cuda_ctx *ctx);
/* Diffusion-map initial layout computed on the device from the loaded graph and installed like set_embedding
 * ≙ the dmap_init branch of one_step_embed (embedder.rs:308-345): DiffusionMaps::embed_from_kgraph with
 * DiffusionParams::new(., Some(5.), Some(12)), alfa 0.5, beta -0.1 (diffmaps.rs:397-587,752-942: density-adapted
 * symmetrised kernel, sparse formulation), rank-20 / 5-iteration subspace SVD (graphlaplace.rs:97-134,
 * tools/svdapprox.rs:343-425,721-801), coordinates (lambda_j/lambda_0)^t U_ij / weight_i clipped to +-10
 * (diffmaps.rs:1145-1243), then set_data_box(., 10) (embedder.rs:345,1376-1408).  The layout has asked_dim columns
 * (<= 19).  gnbn = 0 -> 12, diffusion_time <= 0 -> 5.  The Gaussian test matrix is drawn from the context seed, so
 * the layout is deterministic.  y_out (n x asked_dim) may be NULL. */
int annembed_cuda_dmap_init(annembed_cuda_ctx *ctx, uint32_t gnbn, float diffusion_time, float *y_out);
/* The symmetric normalised kernel annembed_cuda_dmap_init decomposes ≙ GraphLaplacian{sym_kernel, normalizer}
 * (diffmaps.rs:504-587), as  diag + A + A^T  with A = val on the directed edges of the graph (CSR order);
 * sw = sqrt(degrees) (`normalizer`), normed_scale = first-pass scales / their mean (diffmaps.rs:806).  Any output
 * may be NULL.  For tests and diagnostics. */
int annembed_cuda_dmap_kernel(annembed_cuda_ctx *ctx, uint32_t gnbn, float *diag_out /*[n]*/, float *val_out /*[E]*/,
                              float *sw_out /*[n]*/, float *normed_scale_out /*[n]*/);
/* Replace the Gaussian test matrix of the range finder (svdapprox.rs:363) by the caller's (rows = n, 20 columns,
 * row-major) for the following annembed_cuda_dmap_init calls; NULL restores the seeded generator.  Lets a test feed
 * the CPU restatement and the device the same random matrix. */
int annembed_cuda_dmap_set_test_matrix(annembed_cuda_ctx *ctx, const float *omega /*[rows*20], nullable*/, uint64_t rows);
/* Singular values (decreasing) of the rank-20 approximation computed by the last annembed_cuda_dmap_init
 * ≙ the spectrum logged at diffmaps.rs:1182-1197. */
int annembed_cuda_dmap_singular_values(const annembed_cuda_ctx *ctx, double *sigma_out, uint32_t count);
/* Hierarchical embedding, second-step initial layout ≙ h_embed (embedder.rs:245-269): nodes < n_small keep `first`
 * (the first-step layout of the small graph, n_small x asked_dim); every other node i starts at
 * first[proj_node[i]] + clip(sqrt((proj_dist[i] / median_dist) / asked_dim) * N(0,1), 2) per coordinate, where
 * (proj_node, proj_dist) ≙ KGraphProjection::get_projection_by_nodeidx (fromhnsw/kgproj.rs:376) and median_dist
 * ≙ get_projection_distance_quant().query(0.5) (:403).  Entries of proj_* below n_small are ignored. */
int annembed_cuda_set_embedding_from_projection(annembed_cuda_ctx *ctx, uint64_t n_small, const float *first /*[n_small*d]*/,
                                                const uint32_t *proj_node /*[n]*/, const float *proj_dist /*[n]*/,
                                                float median_dist);
/* K2 ≙ estimate_embedded_scales_from_initial_scales (embedder.rs:1356-1373). */
int annembed_cuda_get_embedded_scales(annembed_cuda_ctx *ctx, float *out /*[n]*/);

/* K3 ≙ serial gradient_iteration (embedder.rs:1304-1308) over an explicit sample list, applied strictly in
 * list order: sample s = (flat edge index edge_idx[s], 5 accepted negatives neg_idx[5s..5s+5)). */
int annembed_cuda_step_fixed(annembed_cuda_ctx *ctx, uint64_t n_samples, const uint64_t *edge_idx,
                             const uint32_t *neg_idx, double grad_step);

/* K4(+K5) ≙ entropy_optimize (embedder.rs:794-904): CE, nb_grad_batch batches with the reference's
 * schedule grad_step*(1-iter/nb) (:873-876), CE.  ce_* nullable (skips that K5 pass). */
int annembed_cuda_optimize(annembed_cuda_ctx *ctx, double *ce_initial, double *ce_final);
/* Bounded slice of the schedule: batches iter = first_batch .. first_batch+n_batches-1 of 1..=nb_grad_batch. */
int annembed_cuda_optimize_batches(annembed_cuda_ctx *ctx, uint32_t first_batch, uint32_t n_batches);

/* K5 ≙ ce_compute_threaded (embedder.rs:1127-1163) with cauchy_edge_weight (:1322-1345). */
int annembed_cuda_cross_entropy(annembed_cuda_ctx *ctx, double *out);

/* ≙ the copy-out at embedder.rs:888-899: n x asked_dim, node-index order.  The caller re-indexes to
 * DataId order exactly as get_embedded_reindexed does (embedder.rs:384-405). */
int annembed_cuda_get_embedding(annembed_cuda_ctx *ctx, float *y_out);

/* N3 ≙ Embedder::get_quality_estimate_from_edge_length(nbng) (embedder.rs:620-753) on the current layout:
 * neighbourhood conservation statistics.  The radius (distance to the nbng-th nearest embedded neighbour) is exact
 * (uniform-grid search) where the reference rebuilds an HNSW on the embedded points (embedder.rs:527-554); quantiles
 * are exact where the reference uses CKMS(0.01).  Optional per-node outputs ≙ the csv dumps at :729-743. */
typedef struct annembed_cuda_quality {
    uint64_t nb_without_match;     /* neighbourhoods without a match, :676-678 */
    double   mean_nbmatch;         /* mean number of neighbours conserved when match, :679-680 */
    double   knn_preservation;     /* sum of matches / number of edges */
    double   mean_ratio;           /* mean of edge length / radius, :721-726 */
    double   radius_quantiles[6];  /* 0.05 0.25 0.5 0.75 0.85 0.95 of the embedded radii, :681-690 */
    double   ratio_quantiles[6];   /* same quantiles of edge length / radius, :695-714 */
} annembed_cuda_quality;
int annembed_cuda_quality_estimate(annembed_cuda_ctx *ctx, uint32_t nbng, annembed_cuda_quality *out,
                                   float *radius_out /*[n] nullable*/, float *first_dist_out /*[n] nullable*/,
                                   float *node_ratio_out /*[n] nullable*/);

int annembed_cuda_get_stats(annembed_cuda_ctx *ctx, annembed_cuda_stats *stats);
int annembed_cuda_reset_stats(annembed_cuda_ctx *ctx);

/* Test hook: what the counter-based sampler draws for `epoch` (global mini-epoch index) without applying it.
 * counts_out[e] = number of firings of edge e; neg_out[5e..5e+5) = accepted negatives of its first firing
 * (0xFFFFFFFF where the edge does not fire).  Same stream the epoch kernel consumes. */
int annembed_cuda_debug_draws(annembed_cuda_ctx *ctx, uint32_t epoch, uint32_t *counts_out /*[E]*/,
                              uint32_t *neg_out /*[5E], nullable*/);

#ifdef __cplusplus
}
#endif
#endif /* ANNEMBED_CUDA_H */
