//! Rust side of the drop-in (SOURCE ONLY: there is no Rust toolchain in the build image).
//! `extern "C"` declarations matching include/annembed_cuda.h 1:1, plus the two functions that replace
//! `to_proba_edges` + `entropy_optimize` inside `Embedder::one_step_embed` (annembed src/embedder.rs:351-356)
//! when the crate is built with `--features cuda`.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

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
    pub reserved: u32,
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
    pub fn annembed_cuda_set_graph_csr(ctx: *mut annembed_cuda_ctx, n: u64, row_ptr: *const u64, col: *const u32, dist: *const f32) -> c_int;
    pub fn annembed_cuda_edge_weights(ctx: *mut annembed_cuda_ctx, scale_out: *mut f32, proba_out: *mut f32) -> c_int;
    pub fn annembed_cuda_set_neg_weights(ctx: *mut annembed_cuda_ctx, w: *const f32) -> c_int;
    pub fn annembed_cuda_set_embedding(ctx: *mut annembed_cuda_ctx, y: *const f32) -> c_int;
    /// the dmap_init branch of one_step_embed (src/embedder.rs:308-345) on the device; y_out may be null
    pub fn annembed_cuda_dmap_init(ctx: *mut annembed_cuda_ctx, gnbn: u32, diffusion_time: f32, y_out: *mut f32) -> c_int;
    pub fn annembed_cuda_optimize(ctx: *mut annembed_cuda_ctx, ce_initial: *mut f64, ce_final: *mut f64) -> c_int;
    pub fn annembed_cuda_get_embedding(ctx: *mut annembed_cuda_ctx, y_out: *mut f32) -> c_int;
}

/// What `Embedder::one_step_embed` calls instead of `to_proba_edges` + `entropy_optimize` (embedder.rs:351-356).
/// `neighbours` = `kgraph.get_neighbours()` (kgraph.rs:157) flattened by the caller: row_ptr / col / dist;
/// `initial` = the initial layout, row-major n x asked_dim; `hubness` = clamp(count,1,n) when hubness_weighting.
/// Returns the layout in node-index order (the caller re-indexes as embedder.rs:384-405 does), or Err(String)
/// with the same convention as entropy_optimize (embedder.rs:794-798).
pub fn cuda_entropy_optimize(
    params: &annembed_cuda_params,
    row_ptr: &[u64],
    col: &[u32],
    dist: &[f32],
    initial: &[f32],
    hubness: Option<&[f32]>,
) -> Result<Vec<f32>, String> {
    let n = row_ptr.len() - 1;
    let mut ctx: *mut annembed_cuda_ctx = std::ptr::null_mut();
    let check = |ctx: *mut annembed_cuda_ctx, st: c_int| -> Result<(), String> {
        if st == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(annembed_cuda_last_error(ctx)) }.to_string_lossy().into_owned();
        Err(format!("annembed_cuda status {} : {}", st, msg))
    };
    unsafe {
        check(std::ptr::null_mut(), annembed_cuda_create(&mut ctx, params, 0))?;
        let res = (|| {
            check(ctx, annembed_cuda_set_graph_csr(ctx, n as u64, row_ptr.as_ptr(), col.as_ptr(), dist.as_ptr()))?;
            check(ctx, annembed_cuda_edge_weights(ctx, std::ptr::null_mut(), std::ptr::null_mut()))?;
            if let Some(w) = hubness {
                check(ctx, annembed_cuda_set_neg_weights(ctx, w.as_ptr()))?;
            }
            check(ctx, annembed_cuda_set_embedding(ctx, initial.as_ptr()))?;
            let (mut ce0, mut ce1) = (0f64, 0f64);
            check(ctx, annembed_cuda_optimize(ctx, &mut ce0, &mut ce1))?;
            log::info!(" initial cross entropy value {:.2e}, final {:.2e}", ce0, ce1);
            let mut out = vec![0f32; n * params.asked_dim as usize];
            check(ctx, annembed_cuda_get_embedding(ctx, out.as_mut_ptr()))?;
            Ok(out)
        })();
        annembed_cuda_destroy(ctx);
        res
    }
}
