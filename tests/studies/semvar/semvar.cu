// semvar.cu -- DESIGN-STUDY TOOLING ONLY (CPU).  Variants of the bulk-synchronous mini-epoch semantics built from the
// product's __host__ __device__ primitives (annembed_b200/csrc/sgd_core.cuh), to measure which deviation from the
// reference's sequential loop (embedder.rs:1167-1315) moves the layout statistics.  Never loaded by the product.
//   opt bit 0 (PAIR)   : an in-edge j->i whose reverse edge i->j fired in this mini-epoch continues the pair simulation of
//                        phase A (partner copy) instead of restarting from j's snapshot position
//   opt bit 1 (POISSON): independent Poisson(kappa p_e) firing counts per edge instead of systematic sampling per node
//   classes S          : only nodes of class (tile + epoch) % S fire (with kappa * S); all nodes receive in-edge moves
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include <omp.h>
#include "../../../annembed_b200/csrc/sgd_core.cuh"
using namespace annembed;

struct Var { uint32_t opt; uint32_t S; };

static inline uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x21F0AAADu; x ^= x >> 15; x *= 0x735A2D97u; x ^= x >> 15; return x; }

static inline int poisson_count(float lambda, uint32_t h)
{
    // inverse CDF on a 24-bit uniform
    const float u = u01_24(h);
    float pk = expf(-lambda), cdf = pk; int k = 0;
    while (u >= cdf && k < 60) { k++; pk *= lambda / (float)k; cdf += pk; }
    return k;
}

static inline bool node_active(uint32_t node, uint32_t epoch, uint32_t S) { return S <= 1 || ((node >> 5) + epoch) % S == 0; }

// firing count of edge (src, m-th edge with cumulative interval [Pl,Ph)) in this mini-epoch
static inline int edge_count(const EpochArgs &a, const Var &v, uint32_t src, uint32_t eid, float Pl, float Ph)
{
    if (!node_active(src, a.epoch, v.S)) return 0;
    const float kap = a.kappa * (float)(v.S ? v.S : 1);
    if (v.opt & 2u) return poisson_count(kap * (Ph - Pl), mix32(eid * 0x9E3779B1u + a.ukey));
    const float us = node_uniform(src, a.ukey);
    return cum_ceil(kap, Ph, us) - cum_ceil(kap, Pl, us);
}

template <int DP>
static unsigned epoch_node_var(const EpochArgs &a, const Var &v, const uint32_t *in_eid, uint32_t node, int phase, float *y_mid)
{
    float y[DP], g[DP];
    load_row<DP>(a.y_snap, node, y);
    if (phase == 2) load_row<DP>(y_mid, node, y);
    unsigned s = 0;
    const uint64_t r0 = a.row_ptr[node], r1 = a.row_ptr[node + 1];
    float part[16][DP]; bool fired[16];
    for (int t = 0; t < 16; t++) fired[t] = false;
    if (phase != 2) {
    const float inv_s2 = a.inv_s2[node];
    float P_lo = 0.0f;
    for (uint64_t m = r0; m < r1; m++) {
        const float P_hi = a.cum[m];
        const int c = edge_count(a, v, node, (uint32_t)m, P_lo, P_hi);
        const float pe = F_SUB(P_hi, P_lo);
        P_lo = P_hi;
        if (m - r0 < 16) fired[m - r0] = false;
        if (c <= 0) continue;
        const uint32_t j = a.col[m];
        float yj[DP];
        load_row<DP>(a.y_snap, j, yj);
        const GlobalRowRejector rej{a.col, r0, r1, node, j};
        for (int f = 0; f < c; f++, s++) {
            const uint32_t nk = node & ~3u;
            const Philox4 A = philox4x32_10(nk, s, a.epoch, 1u, a.k0, a.k1);
            uint32_t negs[ANNEMBED_NB_NEG];
            draw_negatives_v2<false>(a, a.epoch, node, s, A, rej, negs);
            apply_firing<DP, true>(a, node, y, yj, g, pe, inv_s2, negs);
        }
        if (m - r0 < 16) { fired[m - r0] = true; for (int cc = 0; cc < DP; cc++) part[m - r0][cc] = yj[cc]; }
    }
    }
    if (phase == 1) { store_row<DP>(y_mid, node, y); return s; }
    float yref[DP];
    for (int cc = 0; cc < DP; cc++) yref[cc] = y[cc];
    if (v.opt & 4u) load_row<DP>(a.y_snap, node, yref);
    const uint64_t q0 = a.in_ptr[node - a.lo], q1 = a.in_ptr[node - a.lo + 1];
    for (uint64_t q = q0; q < q1; q++) {
        const uint4 rec = a.in_rec[q - a.in_base];
        const float Pl = as_float(rec.y), Ph = as_float(rec.z);
        const int c = edge_count(a, v, rec.x, in_eid[q], Pl, Ph);
        if (c <= 0) continue;
        float ys[DP];
        load_row<DP>(phase == 2 ? y_mid : a.y_snap, rec.x, ys);
        if ((v.opt & 1u) && phase != 2) {
            for (uint64_t m = r0; m < r1 && m - r0 < 16; m++)
                if (a.col[m] == rec.x && fired[m - r0]) { for (int cc = 0; cc < DP; cc++) ys[cc] = part[m - r0][cc]; break; }
        }
        const float coef = attract_coeff<true>(sqdist<DP>(yref, ys), F_SUB(Ph, Pl), as_float(rec.w), a.K);
        apply_in_edge<DP>(y, ys, in_edge_factor(coef, c));
    }
    store_row<DP>(a.y_next, node, y);
    return s;
}

// y: n x d in/out; Ms[nb_batch]: mini-epochs of each batch
extern "C" int64_t semvar_optimize(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col, const float *p,
                                   const float *emb_scale, float *y, double grad_step0, uint32_t nbs, uint32_t nb_batch,
                                   const uint32_t *Ms, uint64_t seed, uint32_t opt, uint32_t S)
{
    const int DP = d <= 2 ? 2 : 16;
    const uint64_t E = row_ptr[n];
    std::vector<float> inv_s2(n), cum(E);
    std::vector<uint64_t> in_ptr(n + 1, 0);
    std::vector<uint4> in_rec(E);
    std::vector<uint32_t> in_eid(E);
    for (uint64_t i = 0; i < n; i++) inv_s2[i] = 1.0f / (emb_scale[i] * emb_scale[i]);
    for (uint64_t e = 0; e < E; e++) in_ptr[col[e] + 1]++;
    for (uint64_t i = 0; i < n; i++) in_ptr[i + 1] += in_ptr[i];
    for (uint64_t i = 0; i < n; i++) {
        float acc = 0.0f;
        for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) { acc += p[e]; cum[e] = acc < 1.0f ? acc : 1.0f; }
        cum[row_ptr[i + 1] - 1] = 1.0f;
    }
    std::vector<uint64_t> fill(in_ptr.begin(), in_ptr.end() - 1);
    for (uint64_t i = 0; i < n; i++)
        for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
            uint4 r; r.x = (uint32_t)i; r.w = as_uint(inv_s2[i]);
            r.y = as_uint(e == row_ptr[i] ? 0.0f : cum[e - 1]); r.z = as_uint(cum[e]);
            in_eid[fill[col[e]]] = (uint32_t)e;
            in_rec[fill[col[e]]++] = r;
        }
    std::vector<float> Y[2];
    Y[0].assign(n * DP, 0.0f); Y[1].assign(n * DP, 0.0f);
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) Y[0][i * DP + c] = y[i * d + c];
    std::vector<float> Ymid(n * DP, 0.0f);
    int cur = 0; int64_t total = 0; uint32_t epoch = 0;
    Var v{opt, S ? S : 1};
    for (uint32_t iter = 1; iter <= nb_batch; iter++) {
        const double gs = grad_step0 * (1.0 - (double)iter / (double)nb_batch);
        const uint32_t M = Ms[iter - 1];
        for (uint32_t m = 0; m < M; m++, epoch++) {
            EpochArgs a; memset(&a, 0, sizeof a);
            a.y_snap = Y[cur].data(); a.y_next = Y[cur ^ 1].data();
            a.row_ptr = row_ptr; a.col = col; a.p = p; a.inv_s2 = inv_s2.data();
            a.in_ptr = in_ptr.data(); a.in_rec = in_rec.data(); a.in_base = 0; a.cum = cum.data();
            a.k2 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
            a.n = (uint32_t)n; a.lo = 0; a.hi = (uint32_t)n;
            a.epoch = epoch; a.ukey = epoch_ukey(epoch, a.k2); a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
            a.kappa = (float)((double)nbs * ((double)E / (double)n) / (double)M);
            a.K.gamma = (float)gs; a.K.b = 1.0f; a.K.two_b = 2.0f; a.K.b_is_one = 1;
            int64_t tot = 0;
            for (int phase = (opt & 8u) ? 1 : 0; phase <= ((opt & 8u) ? 2 : 0); phase++) {
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : tot)
            for (int64_t i = 0; i < (int64_t)n; i++)
                tot += DP == 2 ? epoch_node_var<2>(a, v, in_eid.data(), (uint32_t)i, phase, Ymid.data()) : epoch_node_var<16>(a, v, in_eid.data(), (uint32_t)i, phase, Ymid.data());
            }
            total += tot;
            cur ^= 1;
        }
    }
    for (uint64_t i = 0; i < n; i++) for (uint32_t c = 0; c < d; c++) y[i * d + c] = Y[cur][i * DP + c];
    return total;
}
