// quality.cuh -- N3: the reference's embedding-quality estimator on the device.
// Follows /root/reference/src/embedder.rs:478-753 (get_transformed_kgraph, get_max_edge_length_embedded_kgraph,
// get_quality_estimate_from_edge_length).  Difference from the reference (as in oracle/quality.py): the radius R_i =
// distance to the nbng-th nearest embedded neighbour is EXACT (uniform-grid search) instead of coming from an HNSW
// rebuilt on the embedded points (hnsw_rs, not vendored), and quantiles are exact instead of CKMS(0.01).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sgd_core.cuh"

namespace annembed {

// squared embedded distance with one fixed operation order (no fma): the transformed edges and the kNN radii must
// compare bit-for-bit equal when an original neighbour IS the nbng-th embedded neighbour (`<=` at embedder.rs:659)
template <int DP>
__device__ __forceinline__ float emb_dist2(const float (&a)[DP], const float (&b)[DP])
{
    float ds = 0.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) { const float d = __fsub_rn(a[c], b[c]); ds = __fadd_rn(ds, __fmul_rn(d, d)); }
    return ds;
}

// embedder.rs:494-516: per node, embedded L2 distance to every original neighbour kept as a RUNNING MINIMUM in graph
// order (:500-509), then sorted ascending.  One thread per node; rows are short (<= nbng of the graph).
template <int DP>
__global__ void k_transformed_kgraph(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                     const float *__restrict__ Y, float *__restrict__ t)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float yi[DP];
    load_row<DP>(Y, (uint32_t)i, yi);
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    float run = 3.402823466e+38f;
    for (uint64_t m = r0; m < r1; m++) {
        float yj[DP];
        load_row<DP>(Y, col[m], yj);
        run = fminf(__fsqrt_rn(emb_dist2<DP>(yi, yj)), run);
        // insertion into the sorted prefix (values are non-increasing in arrival order, so it goes to the front part)
        uint64_t p = m;
        while (p > r0 && t[p - 1] > run) { t[p] = t[p - 1]; p--; }
        t[p] = run;
    }
}

// ---- exact k-th nearest neighbour distance with a uniform grid on coordinates (0,1) -------------------------------
struct GridParams {
    float x0, y0, inv_h, h;
    int G;                      // cells per side
};

template <int DP>
__global__ void k_cell_of_point(uint64_t n, const float *__restrict__ Y, GridParams gp, uint32_t *__restrict__ cell,
                                uint32_t *__restrict__ idx)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = Y[i * DP], y = DP > 1 ? Y[i * DP + 1] : 0.0f;
    const int cx = min(gp.G - 1, max(0, (int)((x - gp.x0) * gp.inv_h)));
    const int cy = min(gp.G - 1, max(0, (int)((y - gp.y0) * gp.inv_h)));
    cell[i] = (uint32_t)cy * (uint32_t)gp.G + (uint32_t)cx;
    idx[i] = (uint32_t)i;
}

// cell_start[c] = first sorted position whose cell id is >= c  (c in 0..ncell)
__global__ void k_cell_start(uint64_t n, uint64_t ncell, const uint32_t *__restrict__ sorted_cell, uint32_t *__restrict__ cell_start)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > n) return;
    const uint64_t prev = (q == 0) ? 0 : (uint64_t)sorted_cell[q - 1] + 1;
    const uint64_t curr = (q == n) ? ncell + 1 : (uint64_t)sorted_cell[q] + 1;
    for (uint64_t c = prev; c < curr; c++) cell_start[c] = (uint32_t)q;
}

constexpr int KNN_CAP = 1024;   // candidate distances kept in shared memory per warp

// One warp per query (queries taken in cell order for locality).  Ring r grows until the ball of radius r*h around
// the query -- which is contained in the (2r+1)^2 block of cells around the query's cell -- holds at least k+1 points
// (the query itself included); the (k+1)-th smallest squared distance inside the ball is then exact.  Selection is a
// bisection on the float bit pattern (monotone for non-negative floats).
template <int DP>
__global__ void __launch_bounds__(128)
k_knn_radius(uint64_t n, uint32_t k, const float *__restrict__ Y, GridParams gp, const uint32_t *__restrict__ sorted_idx,
             const uint32_t *__restrict__ sorted_cell, const uint32_t *__restrict__ cell_start, float *__restrict__ radius)
{
    __shared__ float s_d2[4][KNN_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t qi = (uint64_t)blockIdx.x * 4 + w;
    if (qi >= n) return;
    const uint32_t p = sorted_idx[qi];
    const uint32_t cell = sorted_cell[qi];
    const int cx = (int)(cell % (uint32_t)gp.G), cy = (int)(cell / (uint32_t)gp.G);
    float yp[DP];
    load_row<DP>(Y, p, yp);
    int r = 1;
    unsigned cnt = 0;
    float lim2 = 0.0f;
    bool whole = false;
    for (;;) {
        const float lim = (float)r * gp.h;
        lim2 = lim * lim;
        const int x_lo = max(0, cx - r), x_hi = min(gp.G - 1, cx + r), y_lo = max(0, cy - r), y_hi = min(gp.G - 1, cy + r);
        whole = (x_lo == 0 && y_lo == 0 && x_hi == gp.G - 1 && y_hi == gp.G - 1);
        cnt = 0;
        for (int row = y_lo; row <= y_hi; row++) {
            const uint32_t b = cell_start[(uint32_t)row * gp.G + x_lo], e = cell_start[(uint32_t)row * gp.G + x_hi + 1];
            for (uint32_t tb = b; tb < e; tb += 32) {          // warp-uniform trip count
                const uint32_t t = tb + lane;
                float d2 = 0.0f;
                bool in = false;
                if (t < e) {
                    float yc[DP];
                    load_row<DP>(Y, sorted_idx[t], yc);
                    d2 = emb_dist2<DP>(yp, yc);
                    in = whole || d2 <= lim2;
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const unsigned slot = cnt + __popc(m & ((1u << lane) - 1u));
                    if (slot < KNN_CAP) s_d2[w][slot] = d2;
                }
                cnt += __popc(m);
            }
        }
        if (cnt >= k + 1 || whole) break;
        r += max(1, r >> 1);
    }
    __syncwarp();
    // (k+1)-th smallest d2 among the cnt in-ball candidates
    const unsigned kk = min(k, cnt - 1);
    uint32_t lo = 0u, hi = whole ? 0x7f7fffffu : __float_as_uint(lim2);
    if (cnt <= KNN_CAP) {
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            unsigned c = 0;
            for (unsigned t = lane; t < cnt; t += 32) c += (__float_as_uint(s_d2[w][t]) <= mid);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c >= kk + 1) hi = mid; else lo = mid + 1;
        }
    } else {
        // dense neighbourhood: bisection by re-scanning the block (rare)
        const int x_lo = max(0, cx - r), x_hi = min(gp.G - 1, cx + r), y_lo = max(0, cy - r), y_hi = min(gp.G - 1, cy + r);
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            unsigned c = 0;
            for (int row = y_lo; row <= y_hi; row++) {
                const uint32_t b = cell_start[(uint32_t)row * gp.G + x_lo], e = cell_start[(uint32_t)row * gp.G + x_hi + 1];
                for (uint32_t t = b + lane; t < e; t += 32) {
                    float yc[DP];
                    load_row<DP>(Y, sorted_idx[t], yc);
                    c += (__float_as_uint(emb_dist2<DP>(yp, yc)) <= mid);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c >= kk + 1) hi = mid; else lo = mid + 1;
        }
    }
    if (lane == 0) radius[p] = __fsqrt_rn(__uint_as_float(lo));
}

// embedder.rs:646-674: per node, number of transformed edges inside the radius, ratio edge / radius per edge,
// mean ratio per node, first (smallest) transformed distance.
__global__ void k_quality_per_node(uint64_t n, const uint64_t *__restrict__ row_ptr, const float *__restrict__ t,
                                   const float *__restrict__ radius, uint32_t *__restrict__ nodes_match,
                                   float *__restrict__ ratio, float *__restrict__ node_ratio, float *__restrict__ first_dist)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    const double R = (double)radius[i];
    uint32_t match = 0;
    double acc = 0.0;
    for (uint64_t m = r0; m < r1; m++) {
        const double e = (double)t[m];
        if (e <= R) match++;                                   // :659-661
        const double q = e / R;                                // :662-664
        ratio[m] = (float)q;
        acc += q;
    }
    nodes_match[i] = match;
    node_ratio[i] = (float)(acc / fmax(1.0, (double)(r1 - r0)));   // :667
    first_dist[i] = t[r0];                                     // :668
}

} // namespace annembed
