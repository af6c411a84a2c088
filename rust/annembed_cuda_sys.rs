//! Rust side of the drop-in (SOURCE ONLY: there is no Rust toolchain in the build image).
//! `extern "C"` declarations matching include/annembed_cuda.h 1:1 -- every function, struct field, status and flag;
//! tests/test_rust_abi_drift.py parses both files and fails when they drift apart.
//! The safe wrappers that replace `to_proba_edges` + `entropy_optimize` inside `Embedder::one_step_embed` /
//! `h_embed` (annembed src/embedder.rs:351-356, 226-230, 245-276) live in rust/embedder_cuda.rs; the
//! `KGraph -> ANNKGCSR` writer (src/fromhnsw/kgraph.rs:157,335-348) in rust/kgraph_csr.rs.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const ANNEMBED_CUDA_ABI_VERSION: u32 = 1;

// enum annembed_status
pub const ANNEMBED_OK: c_int = 0;
pub const ANNEMBED_ERR_INVALID_ARG: c_int = 1;
pub const ANNEMBED_ERR_CUDA: c_int = 2;
pub const ANNEMBED_ERR_EMPTY_ROW: c_int = 3;
pub const ANNEMBED_ERR_UNSORTED_ROW: c_int = 4;
pub const ANNEMBED_ERR_STATE: c_int = 5;
pub const ANNEMBED_ERR_UNSUPPORTED: c_int = 6;
pub const ANNEMBED_ERR_COMM: c_int = 7;
pub const ANNEMBED_ERR_NO_NEGATIVE: c_int = 8;

pub const ANNEMBED_FLAG_NONE: u32 = 0;
pub const ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL: u32 = 1;
pub const ANNEMBED_FLAG_NO_L2_PERSIST: u32 = 2;
pub const ANNEMBED_FLAG_NO_RELABEL: u32 = 4;
pub const ANNEMBED_FLAG_REPLAY_IN_EDGES: u32 = 8;
pub const ANNEMBED_FLAG_LEGACY_EPOCH_KERNELS: u32 = 16;
pub const ANNEMBED_FLAG_BULK_SYNCHRONOUS: u32 = 32;
pub const ANNEMBED_FLAG_NODE_ALIAS: u32 = 64;
pub const ANNEMBED_FLAG_CP_ASYNC_PIPELINE: u32 = 128;
pub const ANNEMBED_FLAG_SECTOR_NEGATIVES: u32 = 256;

/// mirror of EmbedderParams (src/embedparams.rs:76-103) + the device-side knobs
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct annembed_cuda_params {
    pub asked_dim: u32,
    pub dmap_init: u32,
    pub beta: f64,
    pub b: f64,
    pub scale_rho: f64,
    pub grad_step: f64,
    pub nb_sampling_by_edge: u32,
    pub nb_grad_batch: u32,
    pub grad_factor: u32,
    pub hierarchy_layer: u32,
    pub hubness_weighting: u32,
    pub mini_epochs_per_batch: u32,
    pub seed: u64,
    pub flags: u32,
    pub cell_substeps: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct annembed_cuda_stats {
    pub edge_weights_ms: f64,
    pub build_ms: f64,
    pub optimize_ms: f64,
    pub epoch_kernel_ms: f64,
    pub exchange_ms: f64,
    pub cross_entropy_ms: f64,
    pub epoch_launches: u64,
    pub kernel_launches: u64,
    pub positive_samples: u64,
    pub edge_updates: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub model_bytes: f64,
    pub mini_epochs_per_batch: u64,
    pub l2_persist_max_bytes: u64,
    pub l2_window_max_bytes: u64,
    pub n_cells: u64,
    pub cell_nodes: u64,
    pub cell_substeps: u64,
    pub cross_cell_edges: u64,
    pub cross_rank_edges: u64,
    pub exchanges: u64,
}

/// ≙ the statistics logged by get_quality_estimate_from_edge_length (src/embedder.rs:620-753)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct annembed_cuda_quality {
    pub nb_without_match: u64,
    pub mean_nbmatch: f64,
    pub knn_preservation: f64,
    pub mean_ratio: f64,
    pub radius_quantiles: [f64; 6],
    pub ratio_quantiles: [f64; 6],
}

#[repr(C)]
pub struct annembed_cuda_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn annembed_cuda_default_params(p: *mut annembed_cuda_params) -> c_int;
    pub fn annembed_cuda_create(ctx: *mut *mut annembed_cuda_ctx, params: *const annembed_cuda_params, device: c_int) -> c_int;
    pub fn annembed_cuda_destroy(ctx: *mut annembed_cuda_ctx) -> c_int;
    pub fn annembed_cuda_last_error(ctx: *const annembed_cuda_ctx) -> *const c_char;
    // multi-GPU (one context per rank)
    pub fn annembed_cuda_comm_unique_id(unique_id: *mut u8) -> c_int;
    pub fn annembed_cuda_comm_init(ctx: *mut annembed_cuda_ctx, rank: c_int, nranks: c_int, unique_id: *const u8) -> c_int;
    pub fn annembed_cuda_comm_export_layout(ctx: *mut annembed_cuda_ctx, handles: *mut u8) -> c_int;
    pub fn annembed_cuda_comm_import_layouts(ctx: *mut annembed_cuda_ctx, all_handles: *const u8) -> c_int;
    // graph hand-off and edge weights (src/fromhnsw/kgraph.rs:157, src/tools/kdumap.rs:26-235)
    pub fn annembed_cuda_set_graph_csr(ctx: *mut annembed_cuda_ctx, n: u64, row_ptr: *const u64, col: *const u32, dist: *const f32) -> c_int;
    pub fn annembed_cuda_edge_weights(ctx: *mut annembed_cuda_ctx, scale_out: *mut f32, proba_out: *mut f32) -> c_int;
    pub fn annembed_cuda_edge_weights_umap(ctx: *mut annembed_cuda_ctx, norm: f32, scale_out: *mut f32, weight_out: *mut f32, status_out: *mut u8) -> c_int;
    pub fn annembed_cuda_set_edge_weights(ctx: *mut annembed_cuda_ctx, scale: *const f32, proba: *const f32) -> c_int;
    pub fn annembed_cuda_get_perplexity(ctx: *mut annembed_cuda_ctx, out: *mut f32) -> c_int;
    // negative sampler (src/embedder.rs:909-931) and hubness counts (src/fromhnsw/hubness.rs:39-79)
    pub fn annembed_cuda_set_neg_weights(ctx: *mut annembed_cuda_ctx, w: *const f32) -> c_int;
    pub fn annembed_cuda_get_hubness_counts(ctx: *mut annembed_cuda_ctx, counts: *mut u32) -> c_int;
    // initial layout
    pub fn annembed_cuda_set_embedding(ctx: *mut annembed_cuda_ctx, y: *const f32) -> c_int;
    pub fn annembed_cuda_reset_embedding(ctx: *mut annembed_cuda_ctx) -> c_int;
    /// the dmap_init branch of one_step_embed (src/embedder.rs:308-345) on the device; y_out may be null
    pub fn annembed_cuda_dmap_init(ctx: *mut annembed_cuda_ctx, gnbn: u32, diffusion_time: f32, y_out: *mut f32) -> c_int;
    pub fn annembed_cuda_dmap_kernel(ctx: *mut annembed_cuda_ctx, gnbn: u32, diag_out: *mut f32, val_out: *mut f32, sw_out: *mut f32, normed_scale_out: *mut f32) -> c_int;
    pub fn annembed_cuda_dmap_set_test_matrix(ctx: *mut annembed_cuda_ctx, omega: *const f32, rows: u64) -> c_int;
    pub fn annembed_cuda_dmap_singular_values(ctx: *const annembed_cuda_ctx, sigma_out: *mut f64, count: u32) -> c_int;
    /// second-step layout of h_embed (src/embedder.rs:245-269)
    pub fn annembed_cuda_set_embedding_from_projection(ctx: *mut annembed_cuda_ctx, n_small: u64, first: *const f32, proj_node: *const u32, proj_dist: *const f32, median_dist: f32) -> c_int;
    pub fn annembed_cuda_get_embedded_scales(ctx: *mut annembed_cuda_ctx, out: *mut f32) -> c_int;
    // the optimizer (src/embedder.rs:794-904,1167-1315)
    pub fn annembed_cuda_step_fixed(ctx: *mut annembed_cuda_ctx, n_samples: u64, edge_idx: *const u64, neg_idx: *const u32, grad_step: f64) -> c_int;
    pub fn annembed_cuda_optimize(ctx: *mut annembed_cuda_ctx, ce_initial: *mut f64, ce_final: *mut f64) -> c_int;
    pub fn annembed_cuda_optimize_batches(ctx: *mut annembed_cuda_ctx, first_batch: u32, n_batches: u32) -> c_int;
    pub fn annembed_cuda_cross_entropy(ctx: *mut annembed_cuda_ctx, out: *mut f64) -> c_int;
    pub fn annembed_cuda_get_embedding(ctx: *mut annembed_cuda_ctx, y_out: *mut f32) -> c_int;
    /// get_quality_estimate_from_edge_length (src/embedder.rs:620-753) on the device
    pub fn annembed_cuda_quality_estimate(ctx: *mut annembed_cuda_ctx, nbng: u32, out: *mut annembed_cuda_quality, radius_out: *mut f32, first_dist_out: *mut f32, node_ratio_out: *mut f32) -> c_int;
    pub fn annembed_cuda_get_stats(ctx: *mut annembed_cuda_ctx, stats: *mut annembed_cuda_stats) -> c_int;
    pub fn annembed_cuda_reset_stats(ctx: *mut annembed_cuda_ctx) -> c_int;
    pub fn annembed_cuda_debug_draws(ctx: *mut annembed_cuda_ctx, epoch: u32, counts_out: *mut u32, neg_out: *mut u32) -> c_int;
}
