//! Insertion points of the CUDA path inside annembed's `Embedder` (SOURCE ONLY; `--features cuda`).
//!
//! * `cuda_one_step`  replaces the pair `to_proba_edges` + `entropy_optimize` of `Embedder::one_step_embed`
//!   (src/embedder.rs:351-356), including the dmap_init branch (:308-345) when no layout is given;
//! * `cuda_h_embed`   replaces the same pair in both steps of `Embedder::h_embed` (src/embedder.rs:194-295): first step
//!   on the small graph with `grad_factor * nb_grad_batch` batches and grad_step 1 (:205-211), projection-noise initial
//!   layout of the large graph (:245-269) on the device, second step (:275-276);
//! * `cuda_quality`   replaces `get_quality_estimate_from_edge_length` (src/embedder.rs:620-753).
//!
//! All three return node-index order; `get_embedded_reindexed` (src/embedder.rs:384-405) stays as it is.
//! Error convention of `entropy_optimize` (src/embedder.rs:794-798): `Err(String)`; the callers map it to `Err(1)`.
use std::ffi::CStr;
use std::os::raw::c_int;

use super::annembed_cuda_sys::*;
use crate::fromhnsw::kgraph_csr::KGraphCsr;

/// RAII owner of an `annembed_cuda_ctx`
pub struct CudaCtx {
    raw: *mut annembed_cuda_ctx,
    n: usize,
    dim: usize,
}

impl Drop for CudaCtx {
    fn drop(&mut self) {
        unsafe {
            annembed_cuda_destroy(self.raw);
        }
    }
}

impl CudaCtx {
    fn check(&self, st: c_int) -> Result<(), String> {
        if st == ANNEMBED_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(annembed_cuda_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(format!("annembed_cuda status {} : {}", st, msg))
    }

    /// ≙ Embedder::new (src/embedder.rs:107) + the graph hand-off + to_proba_edges (src/tools/kdumap.rs:26-116)
    pub fn new(params: &annembed_cuda_params, graph: &KGraphCsr, device: c_int) -> Result<Self, String> {
        Self::new_with_comm(params, graph, device, None)
    }

    /// Multi-GPU (one process per GPU; the reference is single-process): `comm` = (rank, nranks, the 128-byte id made
    /// by rank 0 with `annembed_cuda_comm_unique_id` and exchanged by the launcher).  comm_init must precede the graph.
    /// The layout handles of every rank (`annembed_cuda_comm_export_layout`) are gathered by the same launcher and given
    /// to `annembed_cuda_comm_import_layouts` to enable the fused NVLink exchange.
    pub fn new_with_comm(
        params: &annembed_cuda_params,
        graph: &KGraphCsr,
        device: c_int,
        comm: Option<(i32, i32, &[u8; 128])>,
    ) -> Result<Self, String> {
        let mut raw: *mut annembed_cuda_ctx = std::ptr::null_mut();
        let st = unsafe { annembed_cuda_create(&mut raw, params, device) };
        if st != ANNEMBED_OK {
            let msg = unsafe { CStr::from_ptr(annembed_cuda_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(format!("annembed_cuda_create status {} : {}", st, msg));
        }
        let ctx = CudaCtx { raw, n: graph.nb_nodes(), dim: params.asked_dim as usize };
        if let Some((rank, nranks, id)) = comm {
            ctx.check(unsafe { annembed_cuda_comm_init(raw, rank, nranks, id.as_ptr()) })?;
        }
        ctx.check(unsafe {
            annembed_cuda_set_graph_csr(raw, graph.nb_nodes() as u64, graph.row_ptr.as_ptr(), graph.col.as_ptr(), graph.dist.as_ptr())
        })?;
        ctx.check(unsafe { annembed_cuda_edge_weights(raw, std::ptr::null_mut(), std::ptr::null_mut()) })?;
        if params.hubness_weighting != 0 {
            // src/embedder.rs:810-837: weights = clamp(in-degree, 1, n); counts ≙ Hubness::get_counts (hubness.rs:39-79)
            let mut counts = vec![0u32; ctx.n];
            ctx.check(unsafe { annembed_cuda_get_hubness_counts(raw, counts.as_mut_ptr()) })?;
            let w: Vec<f32> = counts.iter().map(|&c| (c as f32).max(1.0).min(ctx.n as f32)).collect();
            ctx.check(unsafe { annembed_cuda_set_neg_weights(raw, w.as_ptr()) })?;
        }
        Ok(ctx)
    }

    /// initial layout: explicit (row-major n x dim), or the device diffusion-map layout (src/embedder.rs:308-345)
    pub fn set_initial(&self, initial: Option<&[f32]>) -> Result<(), String> {
        match initial {
            Some(y) => {
                if y.len() != self.n * self.dim {
                    return Err("initial embedding must be n x asked_dim".to_string());
                }
                self.check(unsafe { annembed_cuda_set_embedding(self.raw, y.as_ptr()) })
            }
            None => self.check(unsafe { annembed_cuda_dmap_init(self.raw, 0, 0.0, std::ptr::null_mut()) }),
        }
    }

    /// ≙ entropy_optimize (src/embedder.rs:794-904); returns (layout, initial CE, final CE)
    pub fn optimize(&self) -> Result<(Vec<f32>, f64, f64), String> {
        let (mut ce0, mut ce1) = (0f64, 0f64);
        self.check(unsafe { annembed_cuda_optimize(self.raw, &mut ce0, &mut ce1) })?;
        log::info!(" initial cross entropy value {:.2e}, final {:.2e}", ce0, ce1); // :846-852,885-886
        let mut out = vec![0f32; self.n * self.dim];
        self.check(unsafe { annembed_cuda_get_embedding(self.raw, out.as_mut_ptr()) })?;
        Ok((out, ce0, ce1))
    }

    /// ≙ get_quality_estimate_from_edge_length (src/embedder.rs:620-753) on the current layout
    pub fn quality(&self, nbng: usize) -> Result<annembed_cuda_quality, String> {
        let mut q = annembed_cuda_quality::default();
        self.check(unsafe {
            annembed_cuda_quality_estimate(self.raw, nbng as u32, &mut q, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
        })?;
        log::info!("\n\n neighbourhood conservation: nodes without match {} , mean matches {:.3e}", q.nb_without_match, q.mean_nbmatch);
        Ok(q)
    }

    pub fn stats(&self) -> Result<annembed_cuda_stats, String> {
        let mut s = annembed_cuda_stats::default();
        self.check(unsafe { annembed_cuda_get_stats(self.raw, &mut s) })?;
        Ok(s)
    }
}

/// `Embedder::one_step_embed` with `--features cuda`: replaces src/embedder.rs:308-356.
/// `initial`: `Some` = the caller's layout (random init :348 or anything else), `None` = dmap_init on the device.
pub fn cuda_one_step(params: &annembed_cuda_params, graph: &KGraphCsr, initial: Option<&[f32]>) -> Result<Vec<f32>, String> {
    let ctx = CudaCtx::new(params, graph, 0)?;
    ctx.set_initial(initial)?;
    Ok(ctx.optimize()?.0)
}

/// `Embedder::h_embed` with `--features cuda`: replaces src/embedder.rs:205-230 (first step) and :245-276 (second step).
/// `proj_node[i]`, `proj_dist[i]` ≙ `graph_projection.get_projection_by_nodeidx(&i)` (src/fromhnsw/kgproj.rs:376) for
/// every node index of the large graph (entries below the small graph's size are ignored), `median_dist` ≙
/// `get_projection_distance_quant().query(0.5)` (:403, src/embedder.rs:254).
pub fn cuda_h_embed(
    params: &annembed_cuda_params,
    small: &KGraphCsr,
    large: &KGraphCsr,
    proj_node: &[u32],
    proj_dist: &[f32],
    median_dist: f32,
    first_initial: Option<&[f32]>,
) -> Result<Vec<f32>, String> {
    // first step: src/embedder.rs:205-211
    let mut p1 = *params;
    p1.nb_grad_batch = params.grad_factor * params.nb_grad_batch;
    p1.grad_step = 1.0;
    p1.hierarchy_layer = 0;
    let first = cuda_one_step(&p1, small, first_initial)?;
    // second step: projection-noise layout on the device (src/embedder.rs:245-269), then the optimizer (:275-276)
    if proj_node.len() != large.nb_nodes() || proj_dist.len() != large.nb_nodes() {
        return Err("projection arrays must have one entry per node of the large graph".to_string());
    }
    let ctx = CudaCtx::new(params, large, 0)?;
    ctx.check(unsafe {
        annembed_cuda_set_embedding_from_projection(
            ctx.raw,
            small.nb_nodes() as u64,
            first.as_ptr(),
            proj_node.as_ptr(),
            proj_dist.as_ptr(),
            median_dist,
        )
    })?;
    Ok(ctx.optimize()?.0)
}
