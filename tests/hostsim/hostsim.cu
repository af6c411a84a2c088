// hostsim.cu -- TEST TOOLING ONLY.  Compiles the product's __host__ __device__ mini-epoch body
// (annembed_b200/csrc/sgd_core.cuh) for the HOST so the bulk-synchronous owner-computes semantics of the
// epoch kernel can be replayed on a CPU-only box (design studies, draw-for-draw checks against the GPU).
// It is never loaded by the product package and is not a fallback: nothing in annembed_b200/ references it.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../annembed_b200/csrc/sgd_core.cuh"
#include "../../annembed_b200/csrc/alias_tables.hpp"

using namespace annembed;

extern "C" void hostsim_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

struct HostCtx {
    std::vector<uint64_t> in_ptr;
    std::vector<uint4> in_rec;
    std::vector<float> inv_s2;
    std::vector<float> cum;
};
extern "C" void hostsim_philox2(const uint32_t ctr[2], uint32_t key, uint32_t out[2])
{
    Philox2 r = philox2x32_10(ctr[0], ctr[1], key);
    out[0] = r.x; out[1] = r.y;
}

// the per-node uniforms of one mini-epoch (systematic-sampling offsets), as the kernels compute them
// the grouped alias tables of the hubness sampler (annembed_b200/csrc/alias_tables.hpp), identity numbering
extern "C" void hostsim_alias_tables(uint64_t n, uint32_t G, const float *w, uint32_t *sector_tab /*[8 * ceil(n/4)]*/,
                                     uint32_t *t1_out /*[2 * ceil(n/G)]*/, uint32_t *t2_out /*[2 G * ceil(n/G)]*/)
{
    auto weight = [&](uint64_t i) { return (double)w[i]; };
    std::vector<uint4> tab;
    annembed_host::build_sector_alias_table(n, weight, tab);
    memcpy(sector_tab, tab.data(), tab.size() * sizeof(uint4));
    std::vector<uint2> t1;
    std::vector<uint32_t> t2;
    annembed_host::build_line_alias_tables(n, G, weight, t1, t2);
    memcpy(t1_out, t1.data(), t1.size() * sizeof(uint2));
    memcpy(t2_out, t2.data(), t2.size() * sizeof(uint32_t));
}
// the node-level table (annembed_cuda_set_neg_weights builds it with the same function and the same sequential sum)
extern "C" void hostsim_node_alias_table(uint64_t n, const float *w, uint32_t *tab_out /*[2 n]*/)
{
    double tot = 0.0;
    for (uint64_t i = 0; i < n; i++) tot += w[i];
    std::vector<uint2> tab;
    annembed_host::build_node_alias_table(n, [&](uint64_t i) { return (double)w[i]; }, tot, tab);
    memcpy(tab_out, tab.data(), tab.size() * sizeof(uint2));
}

// multiply-shift maps of a random word to [0, n): the 32-bit one and the 40-bit one of large ranges (philox.cuh)
extern "C" void hostsim_below(uint64_t count, const uint32_t *w, const uint32_t *low8, uint32_t n, uint32_t *out32, uint32_t *out40,
                              uint32_t *out_auto)
{
    for (uint64_t i = 0; i < count; i++) {
        out32[i] = below32(w[i], n);
        out40[i] = below40(w[i], low8[i], n);
        out_auto[i] = below_auto(w[i], low8[i], n);
    }
}

extern "C" void hostsim_node_uniforms(uint32_t node0, uint32_t count, uint32_t epoch, uint64_t seed, float *out)
{
    const uint32_t k2 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
    const uint32_t ukey = epoch_ukey(epoch, k2);
    for (uint32_t i = 0; i < count; i++) out[i] = node_uniform(node0 + i, ukey);
}

static void build(HostCtx &h, uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *p, const float *emb_scale)
{
    const uint64_t E = row_ptr[n];
    h.inv_s2.resize(n);
    for (uint64_t i = 0; i < n; i++) h.inv_s2[i] = 1.0f / (emb_scale[i] * emb_scale[i]);
    h.in_ptr.assign(n + 1, 0);
    for (uint64_t e = 0; e < E; e++) h.in_ptr[col[e] + 1]++;
    for (uint64_t i = 0; i < n; i++) h.in_ptr[i + 1] += h.in_ptr[i];
    h.in_rec.resize(E);
    h.cum.resize(E);
    for (uint64_t i = 0; i < n; i++) {
        float acc = 0.0f;
        for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) { acc += p[e]; h.cum[e] = acc < 1.0f ? acc : 1.0f; }
        h.cum[row_ptr[i + 1] - 1] = 1.0f;
    }
    std::vector<uint64_t> fill(h.in_ptr.begin(), h.in_ptr.end() - 1);
    for (uint64_t i = 0; i < n; i++)
        for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) {   // ascending edge id inside each destination
            uint4 r; r.x = (uint32_t)i; r.w = as_uint(h.inv_s2[i]);
            r.y = as_uint(e == row_ptr[i] ? 0.0f : h.cum[e - 1]); r.z = as_uint(h.cum[e]);
            h.in_rec[fill[col[e]]++] = r;
        }
}

template <int DP, bool HUB>
static uint64_t run_epoch(const EpochArgs &a)
{
    uint64_t tot = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : tot)
    for (int64_t i = (int64_t)a.lo; i < (int64_t)a.hi; i++)
        tot += epoch_node_v2<DP, HUB>(a, (uint32_t)i);
    return tot;
}

template <bool HUB>
static uint64_t run_epoch_dp(int DP, const EpochArgs &a)
{
    switch (DP) {
    case 2: return run_epoch<2, HUB>(a);
    case 4: return run_epoch<4, HUB>(a);
    case 8: return run_epoch<8, HUB>(a);
    case 16: return run_epoch<16, HUB>(a);
    default: return run_epoch<32, HUB>(a);
    }
}

// y: n x d in/out.  neg_alias: nullable n x {prob bits, alias}.  Returns positive samples applied.
extern "C" int64_t hostsim_optimize(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col, const float *p,
                                    const float *emb_scale, float *y, double b, double grad_step0, uint32_t nbs,
                                    uint32_t nb_batch, uint32_t M, uint64_t seed, const uint32_t *neg_alias,
                                    uint32_t first_batch, uint32_t n_batches)
{
    const int DP = d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32;
    HostCtx h;
    build(h, n, row_ptr, col, p, emb_scale);
    std::vector<float> Y[2];
    Y[0].assign(n * DP, 0.0f); Y[1].assign(n * DP, 0.0f);
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) Y[0][i * DP + c] = y[i * d + c];
    int cur = 0;
    const uint64_t E = row_ptr[n];
    int64_t total = 0;
    for (uint32_t iter = first_batch; iter < first_batch + n_batches && iter <= nb_batch; iter++) {
        const double gs = grad_step0 * (1.0 - (double)iter / (double)nb_batch);
        for (uint32_t m = 0; m < M; m++) {
            EpochArgs a;
            memset(&a, 0, sizeof a);
            a.y_snap = Y[cur].data(); a.y_next = Y[cur ^ 1].data();
            a.row_ptr = row_ptr; a.col = col; a.p = p; a.inv_s2 = h.inv_s2.data();
            a.in_ptr = h.in_ptr.data(); a.in_rec = h.in_rec.data(); a.in_base = 0;
            a.neg_alias = (const uint2 *)neg_alias;
            a.cum = h.cum.data(); a.k2 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
            a.n = (uint32_t)n; a.lo = 0; a.hi = (uint32_t)n;
            a.epoch = (iter - 1) * M + m; a.ukey = epoch_ukey(a.epoch, a.k2); a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
            a.kappa = (float)((double)nbs * ((double)E / (double)n) / (double)M);
            a.K.gamma = (float)gs; a.K.b = (float)b; a.K.two_b = (float)(2.0 * b); a.K.b_is_one = b == 1.0;
            total += (int64_t)(neg_alias ? run_epoch_dp<true>(DP, a) : run_epoch_dp<false>(DP, a));
            cur ^= 1;
        }
    }
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) y[i * d + c] = Y[cur][i * DP + c];
    return total;
}


// ---- the cell-resident form (annembed_b200/csrc/cell_epoch.cuh) replayed on the host: identity numbering, cells = a
// fixed grid of `cell_nodes` nodes, S sub-steps per launch.  Inside a launch the nodes of a cell see the current
// positions of their own cell; every other row (partners in other cells, negatives) is the layout at the launch start.
template <int DP, bool HUB>
static uint64_t run_cells(const EpochArgs &a0, uint32_t cell_nodes, uint32_t substeps, uint32_t k2)
{
    const uint32_t n = a0.n;
    const int64_t ncell = (n + cell_nodes - 1) / cell_nodes;
    uint64_t tot = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot)
    for (int64_t c = 0; c < ncell; c++) {
        const uint32_t c0 = (uint32_t)c * cell_nodes, csize = std::min<uint32_t>(cell_nodes, n - c0);
        std::vector<float> cur((size_t)csize * DP), nxt((size_t)csize * DP);
        memcpy(cur.data(), a0.y_snap + (size_t)c0 * DP, cur.size() * sizeof(float));
        EpochArgs a = a0;
        for (uint32_t sub = 0; sub < substeps; sub++) {
            a.epoch = a0.epoch + sub;
            a.ukey = epoch_ukey(a.epoch, k2);
            const CellRows rows{a0.y_snap, cur.data(), c0, csize};
            for (uint32_t i = 0; i < csize; i++) {
                float y[DP];
                tot += epoch_node_rows<DP, HUB, false>(a, c0 + i, rows, y);
                for (int cc = 0; cc < DP; cc++) nxt[(size_t)i * DP + cc] = y[cc];
            }
            cur.swap(nxt);
        }
        memcpy(a0.y_next + (size_t)c0 * DP, cur.data(), cur.size() * sizeof(float));
    }
    return tot;
}

extern "C" int64_t hostsim_optimize_cells(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col, const float *p,
                                          const float *emb_scale, float *y, double b, double grad_step0, uint32_t nbs,
                                          uint32_t nb_batch, uint32_t M, uint64_t seed, const uint32_t *neg_alias,
                                          uint32_t first_batch, uint32_t n_batches, uint32_t cell_nodes, uint32_t S)
{
    const int DP = d <= 2 ? 2 : 4;
    if (d > 4) return -1;
    HostCtx h;
    build(h, n, row_ptr, col, p, emb_scale);
    std::vector<float> Y[2];
    Y[0].assign(n * DP, 0.0f); Y[1].assign(n * DP, 0.0f);
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) Y[0][i * DP + c] = y[i * d + c];
    int cur = 0;
    const uint64_t E = row_ptr[n];
    int64_t total = 0;
    for (uint32_t iter = first_batch; iter < first_batch + n_batches && iter <= nb_batch; iter++) {
        const double gs = grad_step0 * (1.0 - (double)iter / (double)nb_batch);
        for (uint32_t m = 0; m < M; m += S) {
            EpochArgs a;
            memset(&a, 0, sizeof a);
            a.y_snap = Y[cur].data(); a.y_next = Y[cur ^ 1].data();
            a.row_ptr = row_ptr; a.col = col; a.p = p; a.inv_s2 = h.inv_s2.data();
            a.in_ptr = h.in_ptr.data(); a.in_rec = h.in_rec.data(); a.in_base = 0;
            a.neg_alias = (const uint2 *)neg_alias;
            a.cum = h.cum.data(); a.k2 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
            a.n = (uint32_t)n; a.lo = 0; a.hi = (uint32_t)n;
            a.epoch = (iter - 1) * M + m; a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
            a.kappa = (float)((double)nbs * ((double)E / (double)n) / (double)M);
            a.K.gamma = (float)gs; a.K.b = (float)b; a.K.two_b = (float)(2.0 * b); a.K.b_is_one = b == 1.0;
            const uint32_t sub = std::min(S, M - m);
            if (DP == 2) total += (int64_t)(neg_alias ? run_cells<2, true>(a, cell_nodes, sub, a.k2) : run_cells<2, false>(a, cell_nodes, sub, a.k2));
            else total += (int64_t)(neg_alias ? run_cells<4, true>(a, cell_nodes, sub, a.k2) : run_cells<4, false>(a, cell_nodes, sub, a.k2));
            cur ^= 1;
        }
    }
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) y[i * d + c] = Y[cur][i * DP + c];
    return total;
}

// the draws of one mini-epoch (same contract as annembed_cuda_debug_draws; v2 sampler)
extern "C" void hostsim_draws(uint64_t n, const uint64_t *row_ptr, const uint32_t *col, const float *p, uint32_t nbs,
                              uint32_t M, uint64_t seed, uint32_t epoch, const uint32_t *neg_alias, uint32_t *counts,
                              uint32_t *negs_out)
{
    const uint64_t E = row_ptr[n];
    std::vector<float> cum(E);
    for (uint64_t i = 0; i < n; i++) {
        float acc = 0.0f;
        for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) { acc += p[e]; cum[e] = acc < 1.0f ? acc : 1.0f; }
        cum[row_ptr[i + 1] - 1] = 1.0f;
    }
    EpochArgs a;
    memset(&a, 0, sizeof a);
    a.row_ptr = row_ptr; a.col = col; a.p = p; a.neg_alias = (const uint2 *)neg_alias; a.cum = cum.data();
    a.n = (uint32_t)n; a.epoch = epoch; a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    a.k2 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
    a.ukey = epoch_ukey(epoch, a.k2);
    a.kappa = (float)((double)nbs * ((double)E / (double)n) / (double)M);
    for (uint64_t node = 0; node < n; node++) {
        const uint64_t r0 = row_ptr[node], r1 = row_ptr[node + 1];
        const float u = node_uniform((uint32_t)node, a.ukey);
        int c_lo = 0;
        for (uint64_t m = r0; m < r1; m++) {
            const int c_hi = cum_ceil(a.kappa, cum[m], u);
            const int c = c_hi - c_lo;
            counts[m] = (uint32_t)c;
            if (negs_out) {
                uint32_t negs[5] = {ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE};
                if (c > 0) {
                    const uint32_t s = (uint32_t)c_lo;
                    const uint32_t nk = neg_alias ? neg_stream_key<true>(a, (uint32_t)node) : neg_stream_key<false>(a, (uint32_t)node);
                    const Philox4 A = philox4x32_10(nk, s, epoch, 1u, a.k0, a.k1);
                    const GlobalRowRejector rej{col, r0, r1, (uint32_t)node, col[m]};
                    if (neg_alias) draw_negatives_v2<true>(a, a.epoch, (uint32_t)node, s, A, rej, negs);
                    else draw_negatives_v2<false>(a, a.epoch, (uint32_t)node, s, A, rej, negs);
                }
                for (int q = 0; q < 5; q++) negs_out[5 * m + q] = negs[q];
            }
            c_lo = c_hi;
        }
    }
}

// one fixed-list step with the product's fp32 arithmetic (host build of K3's body)
extern "C" void hostsim_step_fixed(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col, const float *p,
                                   const float *emb_scale, float *y, double b, double grad_step, uint64_t n_samples,
                                   const uint64_t *edge_idx, const uint32_t *negs)
{
    const int DP = d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32;
    std::vector<float> Y(n * DP, 0.0f);
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) Y[i * DP + c] = y[i * d + c];
    SgdConst K; K.gamma = (float)grad_step; K.b = (float)b; K.two_b = (float)(2.0 * b); K.b_is_one = b == 1.0;
    for (uint64_t s = 0; s < n_samples; s++) {
        const uint64_t e = edge_idx[s];
        uint64_t lo = 0, hi = n;
        while (hi - lo > 1) { uint64_t mid = (lo + hi) / 2; if (row_ptr[mid] <= e) lo = mid; else hi = mid; }
        const float inv_s2 = 1.0f / (emb_scale[lo] * emb_scale[lo]);
        switch (DP) {
        case 2: fixed_sample<2>(Y.data(), (uint32_t)lo, col[e], p[e], inv_s2, K, negs + 5 * s); break;
        case 4: fixed_sample<4>(Y.data(), (uint32_t)lo, col[e], p[e], inv_s2, K, negs + 5 * s); break;
        case 8: fixed_sample<8>(Y.data(), (uint32_t)lo, col[e], p[e], inv_s2, K, negs + 5 * s); break;
        case 16: fixed_sample<16>(Y.data(), (uint32_t)lo, col[e], p[e], inv_s2, K, negs + 5 * s); break;
        default: fixed_sample<32>(Y.data(), (uint32_t)lo, col[e], p[e], inv_s2, K, negs + 5 * s); break;
        }
    }
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) y[i * d + c] = Y[i * DP + c];
}
