// dmap.cuh -- N2: the reference's diffusion-map initial layout on the device.
// Follows /root/reference/src: embedder.rs:308-345 (dmap_init branch), diffmaps.rs:397-587 (symmetrised kernel, sparse
// branch), :590-679 (node kernels), :752-849 (two scale passes), :852-942 (density), :1145-1243 (coordinates),
// graphlaplace.rs:97-134 + tools/svdapprox.rs:343-425,721-801 (rank-20 subspace iteration, 5 iterations, SVD of Q^T K).
// The symmetric kernel is never materialised: it is  diag + A + A^T  with A = one value per directed edge, applied with
// the graph's CSR (A) and its transposed index (A^T).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"

namespace annembed {

constexpr int DMAP_RANK = 20;           // graphlaplace.rs:113
constexpr int DMAP_ITERS = 5;           // graphlaplace.rs:114
constexpr float DMAP_PROBA_MIN = 1.0e-4f;

// diffmaps.rs:1020-1043: sqrt(sum of the first nbgh squared distances / row length)
__global__ void k_dmap_local_scale(uint64_t n, const uint64_t *__restrict__ row_ptr, const float *__restrict__ dist,
                                   uint32_t nbgh, float *__restrict__ scale)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    float acc = 0.0f;
    for (uint64_t m = r0; m < r1 && m < r0 + nbgh; m++) acc = __fmaf_rn(dist[m], dist[m], acc);
    scale[i] = __fsqrt_rn(acc / (float)(r1 - r0));
}

// diffmaps.rs:790-806: zero scales take the mean; normed = scale / mean
__global__ void k_dmap_fix_scale(uint64_t n, float mean, float *__restrict__ scale, float *__restrict__ normed)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = scale[i];
    if (!(s > 0.0f)) s = mean;
    scale[i] = s;
    normed[i] = s / mean;
}

// diffmaps.rs:590-679 (build_node_param) with remap_weight :815-818: self weight and one kernel weight per out-edge
__global__ void k_dmap_kernel_weights(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                      const float *__restrict__ dist, const float *__restrict__ scale, float sqrt_epsil,
                                      float *__restrict__ w_self, float *__restrict__ w)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    // rows are ascending: the last strictly positive distance is the last entry if it is positive at all
    const float first = dist[r0], last = dist[r1 - 1];
    const bool all_equal = !(last > 0.0f) || last <= first;                       // :614-627
    if (all_equal) {
        const float p = 1.0f / (float)(r1 - r0 + 1);
        w_self[i] = p;
        for (uint64_t m = r0; m < r1; m++) w[m] = p;
        return;
    }
    w_self[i] = 1.0f;
    const float from = scale[i];
    for (uint64_t m = r0; m < r1; m++) {
        const float ls = __fsqrt_rn(__fmul_rn(scale[col[m]], from));
        const float a = dist[m] / __fmul_rn(sqrt_epsil, ls);
        const float v = (float)exp(-(double)__fmul_rn(a, a));                      // f32 exp in the reference; correctly rounded here
        w[m] = fmaxf(v, DMAP_PROBA_MIN);
    }
}

// diffmaps.rs:522-539: sym_e = max(w_e, w of the reverse entry) when (j -> i) exists
__global__ void k_dmap_symmetrise(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                  const float *__restrict__ w, float *__restrict__ sym)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) {
        const uint32_t j = col[m];
        float v = w[m];
        for (uint64_t q = row_ptr[j]; q < row_ptr[j + 1]; q++)
            if (col[q] == (uint32_t)i) v = fmaxf(v, w[q]);
        sym[m] = v;
    }
}

// row sums of  diag + A + A^T : out-edges, in-edges (transposed index: in_ptr / in_eid), diagonal
__global__ void k_dmap_rowsum(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint64_t *__restrict__ in_ptr,
                              const uint32_t *__restrict__ in_eid, const float *__restrict__ val,
                              const float *__restrict__ diag, float diag_factor, float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = (double)diag_factor * (double)diag[i];
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) acc += (double)val[m];
    for (uint64_t q = in_ptr[i]; q < in_ptr[i + 1]; q++) acc += (double)val[in_eid[q]];
    out[i] = (float)acc;
}

// diffmaps.rs:928-935: q <- (q / max_nbng / mean)^beta * mean_scale, written as the second-pass scales
__global__ void k_dmap_beta_scales(uint64_t n, const float *__restrict__ q, double q_norm, float beta, float mean_scale,
                                   float *__restrict__ scale)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    scale[i] = (float)(pow((double)q[i] / q_norm, (double)beta) * (double)mean_scale);
}

// diffmaps.rs:551-556: v_e = sym_e / (q_i q_j)^alfa,  diagonal entry 2 w_self / (q_i^2)^alfa   (q already divided by q_norm)
__global__ void k_dmap_alpha(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                             const float *__restrict__ q, double q_norm, float alfa, float *__restrict__ val,
                             const float *__restrict__ w_self, float *__restrict__ diag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qi = (double)q[i] / q_norm;
    diag[i] = (float)(2.0 * (double)w_self[i] / pow(qi * qi, (double)alfa));
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++)
        val[m] = (float)((double)val[m] / pow(qi * ((double)q[col[m]] / q_norm), (double)alfa));
}

// diffmaps.rs:565-571: symetrization_weights = sqrt(degrees); entries divided by sw_i sw_j
__global__ void k_dmap_sqrt(uint64_t n, float *__restrict__ x)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fsqrt_rn(x[i]);
}
__global__ void k_dmap_normalise(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                 const float *__restrict__ sw, float *__restrict__ val, float *__restrict__ diag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float si = sw[i];
    diag[i] = diag[i] / __fmul_rn(si, si);
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) val[m] = val[m] / __fmul_rn(si, sw[col[m]]);
}

// Gaussian test matrix (svdapprox.rs:363: StandardNormal), Philox + Box-Muller, row-major n x DMAP_RANK
__global__ void k_dmap_gaussian(uint64_t n, uint32_t k0, uint32_t k1, float *__restrict__ omega)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < DMAP_RANK; c += 4) {
        const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)c, 0xD3A9u, k0, k1);
        const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = u01_24(r.y);
        const float u3 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u4 = u01_24(r.w);
        const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
        float s1, c1, s2, c2;
        sincospif(2.0f * u2, &s1, &c1);
        sincospif(2.0f * u4, &s2, &c2);
        float *o = omega + i * DMAP_RANK + c;
        o[0] = ra * c1; o[1] = ra * s1; o[2] = rb * c2; o[3] = rb * s2;
    }
}

// Y = (diag + A + A^T) X  for row-major n x DMAP_RANK blocks (svdapprox.rs:366,378,389: csr * dense); one thread per row
__global__ void __launch_bounds__(128)
k_dmap_spmm(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
            const uint64_t *__restrict__ in_ptr, const uint32_t *__restrict__ in_src, const uint32_t *__restrict__ in_eid,
            const float *__restrict__ val, const float *__restrict__ diag, const float *__restrict__ X, float *__restrict__ Y)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc[DMAP_RANK];
    {
        const float dgi = diag[i];
        const float4 *x = reinterpret_cast<const float4 *>(X + i * DMAP_RANK);
#pragma unroll
        for (int c = 0; c < DMAP_RANK / 4; c++) {
            const float4 t = x[c];
            acc[4 * c] = dgi * t.x; acc[4 * c + 1] = dgi * t.y; acc[4 * c + 2] = dgi * t.z; acc[4 * c + 3] = dgi * t.w;
        }
    }
    auto axpy = [&](float v, uint32_t j) {
        const float4 *x = reinterpret_cast<const float4 *>(X + (uint64_t)j * DMAP_RANK);
#pragma unroll
        for (int c = 0; c < DMAP_RANK / 4; c++) {
            const float4 t = x[c];
            acc[4 * c] = fmaf(v, t.x, acc[4 * c]); acc[4 * c + 1] = fmaf(v, t.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(v, t.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(v, t.w, acc[4 * c + 3]);
        }
    };
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) axpy(val[m], col[m]);
    for (uint64_t q = in_ptr[i]; q < in_ptr[i + 1]; q++) axpy(val[in_eid[q]], in_src[q]);
    float4 *y = reinterpret_cast<float4 *>(Y + i * DMAP_RANK);
#pragma unroll
    for (int c = 0; c < DMAP_RANK / 4; c++) y[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
}

// Gram matrix G = Y^T Y in fp64: each block sweeps row tiles staged in shared memory, thread (a,b) owns entry (a,b);
// per-block partials [gridDim.x][DMAP_RANK^2] are summed by k_dmap_gram_final (deterministic).
constexpr int DMAP_GRAM_ROWS = 64;
__global__ void __launch_bounds__(DMAP_RANK *DMAP_RANK)
k_dmap_gram(uint64_t n, const float *__restrict__ Y, double *__restrict__ partials)
{
    __shared__ float tile[DMAP_GRAM_ROWS][DMAP_RANK + 1];
    const int t = threadIdx.x, a = t / DMAP_RANK, b = t % DMAP_RANK;
    double acc = 0.0;
    for (uint64_t r0 = (uint64_t)blockIdx.x * DMAP_GRAM_ROWS; r0 < n; r0 += (uint64_t)gridDim.x * DMAP_GRAM_ROWS) {
        const int rows = (int)min((uint64_t)DMAP_GRAM_ROWS, n - r0);
        for (int e = t; e < rows * DMAP_RANK; e += DMAP_RANK * DMAP_RANK) tile[e / DMAP_RANK][e % DMAP_RANK] = Y[r0 * DMAP_RANK + e];
        __syncthreads();
        for (int r = 0; r < rows; r++) acc += (double)tile[r][a] * (double)tile[r][b];
        __syncthreads();
    }
    partials[(size_t)blockIdx.x * DMAP_RANK * DMAP_RANK + t] = acc;
}
__global__ void k_dmap_gram_final(unsigned int nblocks, const double *__restrict__ partials, double *__restrict__ G)
{
    const int t = threadIdx.x;
    double acc = 0.0;
    for (unsigned int b = 0; b < nblocks; b++) acc += partials[(size_t)b * DMAP_RANK * DMAP_RANK + t];
    G[t] = acc;
}

// Y <- Y M  (M: DMAP_RANK x ncols, fp64, row-major; the inverse Cholesky factor for the QR step, the eigenvectors of
// the projected problem at the end); one thread per row
__global__ void k_dmap_right_multiply(uint64_t n, int ncols, const double *__restrict__ M, const float *__restrict__ Y,
                                      float *__restrict__ out, int out_stride)
{
    __shared__ double sM[DMAP_RANK * DMAP_RANK];
    for (int e = threadIdx.x; e < DMAP_RANK * ncols; e += blockDim.x) sM[e] = M[e];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float y[DMAP_RANK];
#pragma unroll
    for (int c = 0; c < DMAP_RANK; c++) y[c] = Y[i * DMAP_RANK + c];
    for (int c = 0; c < ncols; c++) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < DMAP_RANK; k++) acc += (double)y[k] * sM[k * ncols + c];
        out[i * out_stride + c] = (float)acc;
    }
}

// diffmaps.rs:1219-1236: coordinate j = clip((lambda_{j+1}/lambda_0)^t U[i][j+1] / weight_i, 10),
// weight_i = normed_scale_i sqrt(sw_i / mean(sw)).  U holds columns 1..d of the singular vectors (n x d); written into the
// padded layout (row stride DP).
__global__ void k_dmap_coordinates(uint64_t n, int d, int DP, const float *__restrict__ U, const float *__restrict__ lam_t,
                                   const float *__restrict__ normed, const float *__restrict__ sw, float sw_mean,
                                   float *__restrict__ y)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float weight = normed[i] * sqrtf(sw[i] / sw_mean);
    for (int c = 0; c < DP; c++) {
        float v = 0.0f;
        if (c < d) v = fminf(fmaxf(lam_t[c] * U[i * d + c] / weight, -10.0f), 10.0f);
        y[i * DP + c] = v;
    }
}

// embedder.rs:1376-1408 (set_data_box): per-column sums, then max |y - mean|, then the scaling
__global__ void __launch_bounds__(256)
k_dmap_colsum(uint64_t n, int d, int DP, const float *__restrict__ y, double *__restrict__ partials /* [grid][32] */)
{
    __shared__ double red[256];
    for (int c = 0; c < d; c++) {
        double acc = 0.0;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += (double)y[i * DP + c];
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) partials[(size_t)blockIdx.x * 32 + c] = red[0];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
k_dmap_center_max(uint64_t n, int d, int DP, const float *__restrict__ means, float *__restrict__ y, float *__restrict__ block_max)
{
    __shared__ float red[256];
    float mx = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        for (int c = 0; c < d; c++) {
            const float v = y[i * DP + c] - means[c];
            y[i * DP + c] = v;
            mx = fmaxf(mx, fabsf(v));
        }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) block_max[blockIdx.x] = red[0];
}
__global__ void k_dmap_scale(uint64_t count, float inv, float *__restrict__ y)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) y[i] = y[i] * inv;
}

} // namespace annembed
