// Per-sample arithmetic of the cross-entropy optimizer and the per-node mini-epoch body.
//
// Follows /root/reference/src/embedder.rs:1167-1302 (ce_optim_edge_shannon): one positive edge is an
// attraction that moves both ends, followed by 5 accepted negatives that repel the origin node only.
// The reference computes coefficients in f64 from f32 coordinates; here everything is fp32
// (tolerance 1e-4 for one step, tests/test_step_fixed.py).
//
// The functions are __host__ __device__ so that tests/hostsim can compile the SAME source for the host
// and replay a mini-epoch on the CPU when no GPU is around (test tooling; the product never does that).
#pragma once
#include <math.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "philox.cuh"

namespace annembed {

#define ANNEMBED_NB_NEG 5          // embedder.rs:1241 asked_nb_neg
#define ANNEMBED_MAX_REDRAW 24     // bounded replacement for the reference's unbounded rejection loop (:1244)
#define ANNEMBED_NO_NODE 0xFFFFFFFFu

struct SgdConst {
    float gamma;     // grad_step of this batch (embedder.rs:875)
    float b;         // Cauchy exponent
    float two_b;
    int   b_is_one;
};

__host__ __device__ __forceinline__ float fast_div(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
// reciprocal: one MUFU.RCP on the device (rcp.approx.ftz, <= 1 ulp), exact division on the host build
__host__ __device__ __forceinline__ float fast_rcp(float x)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
// Explicitly rounded fp32 operations: nvcc may not contract or reorder them, so every kernel that inlines the
// functions below (tiled, generic, fixed-step) produces bit-identical results for the same inputs.
#ifdef __CUDA_ARCH__
#define F_ADD(a, b) __fadd_rn((a), (b))
#define F_SUB(a, b) __fsub_rn((a), (b))
#define F_MUL(a, b) __fmul_rn((a), (b))
#define F_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define F_ADD(a, b) ((a) + (b))
#define F_SUB(a, b) ((a) - (b))
#define F_MUL(a, b) ((a) * (b))
#define F_FMA(a, b, c) fmaf((a), (b), (c))
#endif

__host__ __device__ __forceinline__ float as_float(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint32_t as_uint(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

// ---- rows of the layout: DP floats per node (asked_dim padded with zeros), vector loads ----------
template <int DP>
__host__ __device__ __forceinline__ void load_row(const float *__restrict__ Y, uint32_t idx, float (&v)[DP])
{
    const float *r = Y + (size_t)idx * DP;
    if constexpr (DP == 1) {
        v[0] = r[0];
    } else if constexpr (DP == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(r);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < DP; c += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(r + c);
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        }
    }
}
template <int DP>
__host__ __device__ __forceinline__ void store_row(float *__restrict__ Y, uint32_t idx, const float (&v)[DP])
{
    float *r = Y + (size_t)idx * DP;
    if constexpr (DP == 1) {
        r[0] = v[0];
    } else if constexpr (DP == 2) {
        *reinterpret_cast<float2 *>(r) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int c = 0; c < DP; c += 4)
            *reinterpret_cast<float4 *>(r + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    }
}

template <int DP>
__host__ __device__ __forceinline__ float sqdist(const float (&a)[DP], const float (&b)[DP])
{
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) { const float t = F_SUB(a[c], b[c]); s = F_FMA(t, t, s); }   // embedder.rs:1206-1211 order
    return s;
}

// common coefficient, embedder.rs:1216-1222 / :1276-1282.  B1 = (b == 1) known at compile time (the default model)
template <bool B1>
__host__ __device__ __forceinline__ float cauchy_coeff(float u, float inv_s2, const SgdConst &K)
{
    if (B1 || K.b_is_one) return fast_div(F_MUL(K.two_b, inv_s2), F_ADD(1.0f, u));
    const float pw = powf(u, K.b);
    return F_MUL(F_MUL(F_MUL(K.two_b, fast_div(1.0f, F_ADD(1.0f, pw))), powf(u, F_SUB(K.b, 1.0f))), inv_s2);
}

// attraction coefficient of a positive edge of proba p at squared distance D (embedder.rs:1212-1229).
// Negative = the ends move towards each other, at most 49 % of the gap each.  Only meaningful when D > 0 (:1223).
template <bool B1>
__host__ __device__ __forceinline__ float attract_coeff_raw(float D, float p, float inv_s2, const SgdConst &K)
{
    const float u = F_MUL(D, inv_s2);
    if (B1 || K.b_is_one) {
        // b == 1: gamma * 2/(s^2 (1+u)) * ((1-p)/m - p), m = max(u^2, 1/PROBA_MIN), over one common denominator:
        // a single reciprocal per interaction
        const float m = fmaxf(F_MUL(u, u), 1.0e4f);                               // alfa = 1/PROBA_MIN :1225-1226
        const float num = F_FMA(-p, m, F_SUB(1.0f, p));
        const float c0 = F_MUL(F_MUL(K.gamma, K.two_b), inv_s2);
        return fmaxf(F_MUL(F_MUL(c0, num), fast_rcp(F_MUL(F_ADD(1.0f, u), m))), -0.49f);   // :1228-1229
    }
    const float rep = fast_div(1.0f, fmaxf(F_MUL(u, u), 1.0e4f));
    const float w = F_FMA(F_SUB(1.0f, p), rep, -p);                               // -p + (1-p) * rep
    return fmaxf(F_MUL(F_MUL(K.gamma, cauchy_coeff<B1>(u, inv_s2, K)), w), -0.49f);
}
template <bool B1>
__host__ __device__ __forceinline__ float attract_coeff(float D, float p, float inv_s2, const SgdConst &K)
{
    const float a = attract_coeff_raw<B1>(D, p, inv_s2, K);
    return (F_MUL(D, inv_s2) > 0.0f) ? a : 0.0f;                                  // coincident points: no move (:1223)
}

// Positive edge (i -> j, proba p): embedder.rs:1202-1238.  yi/yj are local copies, g the sample's gradient.
// Branch-free: the coefficient is always evaluated and discarded by a select when the points coincide.
template <int DP, bool B1 = false>
__host__ __device__ __forceinline__ void attract(float (&yi)[DP], float (&yj)[DP], float (&g)[DP], float p,
                                                 float inv_s2, const SgdConst &K)
{
    const float D = sqdist<DP>(yi, yj);
    const float a = attract_coeff_raw<B1>(D, p, inv_s2, K);
    const bool ok = F_MUL(D, inv_s2) > 0.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) {
        const float gn = F_MUL(F_SUB(yj[c], yi[c]), a);                           // :1230
        g[c] = ok ? gn : g[c];
        yi[c] = F_SUB(yi[c], g[c]);                                               // :1237-1238
        yj[c] = F_ADD(yj[c], g[c]);
    }
}

// Negative node k: embedder.rs:1263-1297.  Only yi moves; g keeps its previous value when the two points
// coincide (the reference does the same).  `use` = false skips the negative entirely (no acceptable node found).
template <int DP, bool B1 = false>
__host__ __device__ __forceinline__ void repulse(float (&yi)[DP], const float (&yk)[DP], float (&g)[DP],
                                                 float inv_s2, const SgdConst &K, bool use = true)
{
    const float dk = sqdist<DP>(yi, yk);
    const float u = F_MUL(dk, inv_s2);
    float a;
    if (B1 || K.b_is_one) {                                                       // one reciprocal, see attract_coeff_raw
        const float m = fmaxf(F_MUL(u, u), 0.0625f);                              // alfa = 1/16 :1286-1288
        const float c0 = F_MUL(F_MUL(K.gamma, K.two_b), inv_s2);
        a = fminf(F_MUL(c0, fast_rcp(F_MUL(F_ADD(1.0f, u), m))), 2.0f);
    } else {
        const float rep = fast_div(1.0f, fmaxf(F_MUL(u, u), 0.0625f));
        a = fminf(F_MUL(F_MUL(K.gamma, cauchy_coeff<B1>(u, inv_s2, K)), rep), 2.0f);
    }
    const bool ok = dk > 0.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) {
        const float gn = F_MUL(F_SUB(yk[c], yi[c]), a);
        g[c] = (ok && use) ? gn : g[c];
        yi[c] = use ? F_SUB(yi[c], g[c]) : yi[c];
    }
}

// ---- destination side of a positive edge (src -> node) inside a mini-epoch -------------------------------------
// The destination's owner evaluates the attraction coefficient once, at its position after its own firings
// (y_ref), for the c firings the source drew for this edge.  With the coefficient a frozen, the pair's gap shrinks
// by (1 + 2a) per firing (both ends move, embedder.rs:1237-1238), so the destination moves by
//   A * (y - y_src),  A = a * (1 + (1+2a) + ... + (1+2a)^(c-1)).
// Applying "y += A (y - y_src)" entry after entry keeps every step a contraction towards y_src (|A| < 1/2) and
// leaves only a 2-flop dependency between consecutive in-edges of a node.
__host__ __device__ __forceinline__ float in_edge_factor(float a, int c)
{
    float A = a, f = 1.0f;
    const float r = F_FMA(2.0f, a, 1.0f);
    for (int i = 1; i < c; i++) { f = F_MUL(f, r); A = F_FMA(a, f, A); }
    return A;
}
template <int DP>
__host__ __device__ __forceinline__ void apply_in_edge(float (&y)[DP], const float (&ys)[DP], float A)
{
#pragma unroll
    for (int c = 0; c < DP; c++) y[c] = F_FMA(F_SUB(y[c], ys[c]), A, y[c]);
}

// One complete reference sample applied in place on a layout (serial semantics, K3).
template <int DP>
__host__ __device__ __forceinline__ void fixed_sample(float *Y, uint32_t i, uint32_t j, float p, float inv_s2,
                                                      const SgdConst &K, const uint32_t *negs)
{
    float yi[DP], yj[DP], g[DP], yk[DP];
    load_row<DP>(Y, i, yi);
    load_row<DP>(Y, j, yj);
#pragma unroll
    for (int c = 0; c < DP; c++) g[c] = 0.0f;
    attract<DP>(yi, yj, g, p, inv_s2, K);
    store_row<DP>(Y, j, yj);                                                      // :1239
    for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
        load_row<DP>(Y, negs[q], yk);
        repulse<DP>(yi, yk, g, inv_s2, K);
    }
    store_row<DP>(Y, i, yi);                                                      // :1301
}

// ---- arguments of one mini-epoch ------------------------------------------------------------------------------
// Every index below lives in the INTERNAL node numbering of the optimizer context (locality relabelling of the graph,
// DESIGN.md 4): the layout buffers, rows, transposed index and alias table are all permuted consistently, the
// permutation is undone at the C ABI.
struct EpochArgs {
    const float *__restrict__ y_snap;   // layout at the start of the mini-epoch (replicated, n x DP)
    float *__restrict__ y_next;         // layout after it; only rows [lo,hi) are written here
    const uint64_t *__restrict__ row_ptr;
    const uint32_t *__restrict__ col;
    const float *__restrict__ p;
    const float *__restrict__ inv_s2;   // 1 / embedded_scale^2 per node
    const uint64_t *__restrict__ in_ptr; // transposed index of the owned nodes: in_ptr[node-lo] .. in_ptr[node-lo+1]
    const uint4 *__restrict__ in_rec;   // {src node, bits(P_lo), bits(P_hi), bits(inv_s2[src])}, entry q at in_rec[q - in_base]
    uint64_t in_base;
    const uint2 *__restrict__ neg_alias; // {bits(prob), alias} per node, hubness sampler (embedder.rs:909-931)
    const uint4 *__restrict__ sec_alias; // sector-level alias table (2 x uint4 per sector of 4 nodes), null: node-level table only
    const uint2 *__restrict__ line_t1;   // line-level alias tables of the event kernels (annembed_cuda.cu build_sector_alias), or null
    const uint32_t *__restrict__ line_t2;
    uint32_t neg_group_shift;            // uniform sampler: 2^shift consecutive nodes share their negative streams (0 -> 2, see neg_stream_key)
    const float *__restrict__ cum;       // inclusive cumulative probability along each row (last entry exactly 1)
    // tiled kernels: rows padded to KP entries {col, bits(cum)} (pads: {NO_NODE, 1.0f}), 16-byte aligned per node
    const uint2 *__restrict__ rowpack;
    const uint32_t *__restrict__ erank;  // [n][KP]: position q of the out-edge in the transposed index
    unsigned char *__restrict__ fired;   // [E]: firing counts pushed by k_epoch_out, consumed (and cleared) by k_epoch_in_flags;
                                         // null: the in-edge kernel replays the sources' decisions instead (multi-rank)
    uint32_t k2;                         // Philox2x32 key of the per-mini-epoch key below
    uint32_t ukey;                       // epoch_ukey(epoch, k2): key of the per-node uniforms of this mini-epoch
    uint32_t n_peers;                    // fused exchange: replicas of y_next on the other ranks (peer memory over NVLink),
    float *peer_next[7];                 // or ONE multicast mapping of all replicas (NVSwitch replicates the store)
    uint32_t n, lo, hi;
    uint32_t epoch, k0, k1;
    float kappa;                        // expected firings of edge e in this mini-epoch = kappa * p_e
    SgdConst K;
};

template <bool HUB>
__host__ __device__ __forceinline__ uint32_t map_negative(const EpochArgs &a, uint32_t w_idx, uint32_t w_acc)
{
    uint32_t k = below_auto(w_idx, w_acc, a.n);                                   // uniform, embedder.rs:1121 (w_acc: its top 24 bits are the accept word)
    if constexpr (HUB) {
        const uint2 t = a.neg_alias[k];
        if (!(u01_24(w_acc) < as_float(t.x))) k = t.y;                            // alias method, :919,929
    }
    return k;
}

// =====================================================================================================
// the sampler: systematic (low-variance) sampling per node.
// Node i owns one uniform u_i(epoch) (node_uniform below).  Its sample points are s + u_i,
// s = 0,1,..; edge m of the row covers [kappa*P_{m-1}, kappa*P_m) where P is the cumulative edge probability
// of the row (row sums are 1), so  count_m = ceil(kappa*P_m - u) - ceil(kappa*P_{m-1} - u),
// E[count_m] = kappa * p_m  (the reference's expectation, embedder.rs:858,987,1182) and every node fires
// ceil(kappa - u) times per mini-epoch: the reference samples every node at the same rate (P(e) = p_e / N).
// The destination's owner replays the decision from (src, epoch, P_lo, P_hi): no communication, no atomics.
// =====================================================================================================
// ukey = Philox2x32-10(epoch; seed) is drawn once per mini-epoch on the host; per node the uniform is a 2-round
// multiply-xorshift finaliser of (node * golden + ukey) -- 9 integer instructions instead of a 10-round Philox per
// node AND per in-edge (the in-edge sweep replays its source's uniform).  Top 24 bits -> [0,1).
__host__ __device__ __forceinline__ uint32_t epoch_ukey(uint32_t epoch, uint32_t k2)
{
    return philox2x32_10(epoch, 0x75A1C0DEu, k2).x;
}
__host__ __device__ __forceinline__ float node_uniform(uint32_t node, uint32_t ukey)
{
    uint32_t x = node * 0x9E3779B1u + ukey;
    x ^= x >> 16; x *= 0x21F0AAADu;
    x ^= x >> 15; x *= 0x735A2D97u;
    x ^= x >> 15;
    return u01_24(x);
}
__host__ __device__ __forceinline__ int cum_ceil(float kappa, float P, float u)
{
#ifdef __CUDA_ARCH__
    return __float2int_ru(fmaf(kappa, P, -u));                // one F2I.CEIL
#else
    return (int)ceilf(fmaf(kappa, P, -u));
#endif
}

// the 5 accepted negatives of firing s of `node` (v2 streams: counter (node, sub, epoch, tag))
//   tag 1: sub = s          words x,y,z,w -> negatives 0..3; negative 4 from fifth_word(block): ONE Philox block per firing
//   tag 3: sub = s          accept words of the hubness alias sampler (fifth accept word likewise)
//   tag 0x80000000|q<<8|t   redraw t of negative q after a rejection (embedder.rs:1246-1252)
__host__ __device__ __forceinline__ uint32_t philox_word(const Philox4 &B, uint32_t i)
{
    return (i & 2u) ? ((i & 1u) ? B.w : B.z) : ((i & 1u) ? B.y : B.x);
}

// Fifth 32-bit word of a firing: a mix of the block's four words (rotations + the multiply-xorshift finaliser of
// node_uniform).  The negatives use the top ~24 bits of each word; the fifth word also depends on the 32 low bits that
// no other negative sees, so given the first four negatives it is still uniform for every practical purpose
// (chi-square of the accepted negatives: tests/test_host.py).
__host__ __device__ __forceinline__ uint32_t fifth_word(const Philox4 &A)
{
    uint32_t x = A.x ^ ((A.y << 8) | (A.y >> 24)) ^ ((A.z << 16) | (A.z >> 16)) ^ ((A.w << 24) | (A.w >> 8));
    x ^= x >> 16; x *= 0x21F0AAADu;
    x ^= x >> 15; x *= 0x735A2D97u;
    x ^= x >> 15;
    return x + 0x9E3779B9u;
}

struct GlobalRowRejector {          // nodeparam.rs:83-85 linear scan of the origin's row in global memory
    const uint32_t *__restrict__ col;
    uint64_t r0, r1;
    uint32_t node, j;
    __host__ __device__ __forceinline__ bool operator()(uint32_t k) const
    {
        if (k == node || k == j) return true;
        for (uint64_t m = r0; m < r1; m++)
            if (col[m] == k) return true;
        return false;
    }
};

// Counter word 0 of the negative streams.  Uniform sampler (default): the 4 nodes of an aligned group (ids 4g..4g+3,
// i.e. 4 adjacent lanes of a warp tile) share the stream, so that for every negative slot they draw the SAME random
// 32-byte sector of the layout and each takes a different row of it (rotated by a random offset): every sample still
// gets 5 independent, uniformly distributed negatives -- exactly the reference's per-sample law (embedder.rs:1121) --
// while the 4 lanes' gathers coalesce into one sector request.  Only samples of different nodes of a group become
// correlated, which no statistic of the optimizer depends on.  The hubness (alias) sampler does the same one level up:
// the group draws one random sector of the ALIAS TABLE, each lane reads a different entry of it and then accepts it or
// follows its alias (embedder.rs:909-931): per sample, 5 independent draws from the hubness law.
// The asynchronous event kernels widen the group of the uniform sampler from the rows of a 32-byte sector to the rows of a
// 128-byte LINE (16 nodes in dimension 2, 8 in dimension 3-4: EpochArgs::neg_group_shift): per negative slot a warp then
// asks the memory system for 2 (4) lines of 4 (2) sectors instead of 8 scattered sectors -- the same bytes in a quarter
// (half) of the requests, and DRAM bursts of 128 bytes.  Every sample still draws 5 independent uniform negatives.
__host__ __device__ __forceinline__ uint32_t neg_group_shift(const EpochArgs &a) { return a.neg_group_shift ? a.neg_group_shift : 2u; }
template <bool HUB>
__host__ __device__ __forceinline__ uint32_t neg_stream_key(const EpochArgs &a, uint32_t node)
{
    if constexpr (HUB) return a.line_t1 ? node & ~((1u << neg_group_shift(a)) - 1u) : node & ~3u;
    else return node & ~((1u << neg_group_shift(a)) - 1u);
}

// redraws of the rejected negatives (embedder.rs:1246-1252: the reference loops until accepted; bounded here)
template <bool HUB, class Rej>
__host__ __device__ __forceinline__ void redraw_negatives(const EpochArgs &a, uint32_t epoch, uint32_t node, uint32_t s, const Rej &rejected,
                                                          const bool (&rej)[ANNEMBED_NB_NEG], uint32_t (&negs)[ANNEMBED_NB_NEG])
{
#pragma unroll
    for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
        uint32_t k = negs[q];
        bool r = rej[q];
        for (uint32_t t = 0; r && t < ANNEMBED_MAX_REDRAW; t++) {
            const Philox4 R = philox4x32_10(node, s, epoch, 0x80000000u | ((uint32_t)q << 8) | t, a.k0, a.k1);
            k = map_negative<HUB>(a, R.x, R.y);
            r = rejected(k);
        }
        negs[q] = r ? ANNEMBED_NO_NODE : k;
    }
}

// first draw of the 5 negatives from the index words of block A (and the accept words of block C, hubness sampler), then
// the rare redraws.  `rot` distinguishes the 4 members of the group that shares A / C: each takes a different row of the
// drawn sector (rotated by two random bits of the word).
template <bool HUB, class Rej>
__host__ __device__ __forceinline__ void draw_negatives_core(const EpochArgs &a, uint32_t epoch, uint32_t node, uint32_t s, uint32_t rot,
                                                             const Philox4 &A, const Philox4 &C, const Rej &rejected,
                                                             uint32_t (&negs)[ANNEMBED_NB_NEG])
{
    uint32_t wi[ANNEMBED_NB_NEG] = {A.x, A.y, A.z, A.w, fifth_word(A)};
    uint32_t wa[ANNEMBED_NB_NEG] = {0u, 0u, 0u, 0u, 0u};
    if constexpr (HUB) { wa[0] = C.x; wa[1] = C.y; wa[2] = C.z; wa[3] = C.w; wa[4] = fifth_word(C); }
    const uint32_t nsec = (a.n + 3u) >> 2;
    const uint32_t gsh = neg_group_shift(a), gmask = (1u << gsh) - 1u, ngrp = (a.n + gmask) >> gsh;
    (void)gsh; (void)gmask; (void)ngrp; (void)nsec;
    // branch-free; the redraws (probability ~ (deg+2)/n per negative) are a rare path
    bool any_rej = false;
    bool rej[ANNEMBED_NB_NEG];
#pragma unroll
    for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
        uint32_t k;
        bool out_of_range;
        if (HUB && a.line_t1 != nullptr) {
            // line-level alias method (event kernels): the group = the nodes of one 128-byte line of the layout shares the
            // line draw and its accept / alias decision; lane `rot` then looks at ITS column of the final line's inner
            // alias table (columns rotated by a shared random offset: distinct within the group, uniform per lane)
            const uint32_t l0 = below_auto(wi[q], wa[q], ngrp);                 // (the accept test reads the top 24 bits of wa)
            uint32_t h = wi[q] ^ ((wa[q] << 16) | (wa[q] >> 16));
            h ^= h >> 16; h *= 0x21F0AAADu; h ^= h >> 15; h *= 0x735A2D97u; h ^= h >> 15;
            const uint32_t c = (rot + (h >> 28)) & gmask;                       // top bits: the rotation; low 24 bits: the accept
            // one round of reads: the line's {prob, alias line} and the lane's column of BOTH candidate lines' inner tables
            const uint2 t1 = a.line_t1[l0];
            const uint32_t *e2 = a.line_t2 + ((size_t)l0 << (gsh + 1)) + c;
            const uint32_t c_own = e2[0], c_alias = e2[gmask + 1u];
            const bool own = u01_24(wa[q]) < as_float(t1.x);
            const uint32_t line = own ? l0 : t1.y, t2 = own ? c_own : c_alias;
            k = (line << gsh) + ((h & 0xFFFFFFu) < (t2 >> 4) ? c : (t2 & 15u));
            out_of_range = k >= a.n;
        } else if (HUB && a.sec_alias != nullptr) {
            // sector-level alias method: the group shares the sector draw, its accept / alias decision and one uniform;
            // lane `rot` picks its row inside the final sector from the rotation frac(u + rot / 4) of that uniform
            // (the entry carries the thresholds of its alias sector too: one 32-byte gather, no branch)
            const uint32_t s0 = below_auto(wi[q], wa[q], nsec);                // (the accept test reads the top 24 bits of wa)
            const uint4 e0 = a.sec_alias[2 * (size_t)s0], e1 = a.sec_alias[2 * (size_t)s0 + 1];
            const bool own = u01_24(wa[q]) < as_float(e0.x);
            const uint32_t sct = own ? s0 : e0.y;
            const float t0 = as_float(own ? e0.z : e1.y), t1 = as_float(own ? e0.w : e1.z), t2 = as_float(own ? e1.x : e1.w);
            uint32_t h = wi[q] ^ ((wa[q] << 16) | (wa[q] >> 16));
            h ^= h >> 16; h *= 0x21F0AAADu; h ^= h >> 15; h *= 0x735A2D97u; h ^= h >> 15;
            float ul = u01_24(h) + 0.25f * (float)(rot & 3u);
            ul = ul >= 1.0f ? ul - 1.0f : ul;
            k = (sct << 2) + (ul >= t0 ? 1u : 0u) + (ul >= t1 ? 1u : 0u) + (ul >= t2 ? 1u : 0u);
            out_of_range = k >= a.n;
        } else {
            // random sector / line, rotated row; the 8 extra bits of the wide draw come from the middle of the next word
            if constexpr (HUB) k = (below_auto(wi[q], wa[q], nsec) << 2) | ((rot + wi[q]) & 3u);
            else k = (below_auto(wi[q], wi[(q + 1) % ANNEMBED_NB_NEG] >> 4, ngrp) << gsh) | ((rot + wi[q]) & gmask);
            out_of_range = k >= a.n;
            if constexpr (HUB) {
                if (!out_of_range) {                                           // node-level alias method on the shared sector's entries
                    const uint2 t = a.neg_alias[k];
                    if (!(u01_24(wa[q]) < as_float(t.x))) k = t.y;
                }
            }
        }
        rej[q] = out_of_range || rejected(k);
        any_rej |= rej[q];
        negs[q] = k;
    }
    if (any_rej) redraw_negatives<HUB>(a, epoch, node, s, rejected, rej, negs);
}

// bulk-synchronous kernels and the multi-firing sweep: the group is the 4 nodes of an aligned id quadruple (neg_stream_key)
template <bool HUB, class Rej>
__host__ __device__ __forceinline__ void draw_negatives_v2(const EpochArgs &a, uint32_t epoch, uint32_t node, uint32_t s, const Philox4 &A,
                                                           const Rej &rejected, uint32_t (&negs)[ANNEMBED_NB_NEG])
{
    Philox4 C;
    C.x = C.y = C.z = C.w = 0u;
    if constexpr (HUB) C = philox4x32_10(neg_stream_key<HUB>(a, node), s, epoch, 3u, a.k0, a.k1);
    draw_negatives_core<HUB>(a, epoch, node, s, node, A, C, rejected, negs);
}

// one firing of `node` on edge (node -> j): attraction against the local copy of y_j, then 5 repulsions
template <int DP, bool B1 = false>
__host__ __device__ __forceinline__ void apply_firing(const EpochArgs &a, uint32_t node, float (&y)[DP], float (&yj)[DP],
                                                      float (&g)[DP], float pe, float inv_s2,
                                                      const uint32_t (&negs)[ANNEMBED_NB_NEG])
{
#pragma unroll
    for (int cc = 0; cc < DP; cc++) g[cc] = 0.0f;
    if constexpr (DP <= 4) {
        float yk[ANNEMBED_NB_NEG][DP];                  // issue the five gathers before the dependent arithmetic
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++)
            load_row<DP>(a.y_snap, negs[q] == ANNEMBED_NO_NODE ? node : negs[q], yk[q]);
        attract<DP, B1>(y, yj, g, pe, inv_s2, a.K);
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++)
            if (negs[q] != ANNEMBED_NO_NODE) repulse<DP>(y, yk[q], g, inv_s2, a.K);
    } else {
        attract<DP, B1>(y, yj, g, pe, inv_s2, a.K);
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
            if (negs[q] == ANNEMBED_NO_NODE) continue;
            float yk[DP];
            load_row<DP>(a.y_snap, negs[q], yk);
            repulse<DP, B1>(y, yk, g, inv_s2, a.K);
        }
    }
}

// rows of the layout as a mini-epoch sees them: the global snapshot, or (cell-resident kernel, cell_epoch.cuh) the current
// positions of the node's own cell for the nodes of that cell and the snapshot for all the others
struct SnapshotRows {
    const float *__restrict__ y_snap;
    template <int DP>
    __host__ __device__ __forceinline__ void load(uint32_t idx, float (&v)[DP]) const { load_row<DP>(y_snap, idx, v); }
};
struct CellRows {
    const float *__restrict__ y_snap;
    const float *__restrict__ ycur;     // rows c0 .. c0 + csize of the current sub-step
    uint32_t c0, csize;
    template <int DP>
    __host__ __device__ __forceinline__ void load(uint32_t idx, float (&v)[DP]) const
    {
        const uint32_t loc = idx - c0;
        if (loc < csize) load_row<DP>(ycur, loc, v);
        else load_row<DP>(y_snap, idx, v);
    }
};

// in_rec: {src node, bits(P_lo), bits(P_hi), bits(inv_s2[src])}; p_e is taken as P_hi - P_lo on both sides.
// a.in_ptr is indexed by (node - a.lo).  The result is returned in `y` (not stored).
template <int DP, bool HUB, bool B1, class Rows>
__host__ __device__ __forceinline__ unsigned int epoch_node_rows(const EpochArgs &a, uint32_t node, const Rows &rows, float (&y)[DP])
{
    float g[DP];
    rows.template load<DP>(node, y);
    const float inv_s2 = a.inv_s2[node];
    const uint64_t r0 = a.row_ptr[node], r1 = a.row_ptr[node + 1];
    const float u = node_uniform(node, a.ukey);
    unsigned int s = 0;
    // phase A: out-edges in row order; firing index s runs over the node's sample points
    float P_lo = 0.0f;
    int c_lo = 0;                                     // ceil(-u) == 0 for u in [0,1)
    for (uint64_t m = r0; m < r1; m++) {
        const float P_hi = a.cum[m];
        const int c_hi = cum_ceil(a.kappa, P_hi, u);
        const int c = c_hi - c_lo;
        const float pe = F_SUB(P_hi, P_lo);
        P_lo = P_hi; c_lo = c_hi;
        if (c <= 0) continue;
        const uint32_t j = a.col[m];
        float yj[DP];
        rows.template load<DP>(j, yj);
        const GlobalRowRejector rej{a.col, r0, r1, node, j};
        for (int f = 0; f < c; f++, s++) {
            const uint32_t nk = neg_stream_key<HUB>(a, node);
            const Philox4 A = philox4x32_10(nk, s, a.epoch, 1u, a.k0, a.k1);
            uint32_t negs[ANNEMBED_NB_NEG];
            draw_negatives_v2<HUB>(a, a.epoch, node, s, A, rej, negs);
            apply_firing<DP, B1>(a, node, y, yj, g, pe, inv_s2, negs);
        }
    }
    // phase B: in-edges in transposed-index order, coefficients evaluated at the position after phase A
    float yref[DP];
#pragma unroll
    for (int cc = 0; cc < DP; cc++) yref[cc] = y[cc];
    const uint64_t q0 = a.in_ptr[node - a.lo], q1 = a.in_ptr[node - a.lo + 1];
    for (uint64_t q = q0; q < q1; q++) {
        const uint4 rec = a.in_rec[q - a.in_base];
        const float us = node_uniform(rec.x, a.ukey);
        const float Pl = as_float(rec.y), Ph = as_float(rec.z);
        const int c = cum_ceil(a.kappa, Ph, us) - cum_ceil(a.kappa, Pl, us);
        if (c <= 0) continue;
        float ys[DP];
        rows.template load<DP>(rec.x, ys);
        const float coef = attract_coeff<B1>(sqdist<DP>(yref, ys), F_SUB(Ph, Pl), as_float(rec.w), a.K);
        apply_in_edge<DP>(y, ys, in_edge_factor(coef, c));
    }
    return s;
}

template <int DP, bool HUB, bool B1 = false>
__host__ __device__ __forceinline__ unsigned int epoch_node_v2(const EpochArgs &a, uint32_t node)
{
    float y[DP];
    const SnapshotRows rows{a.y_snap};
    const unsigned int s = epoch_node_rows<DP, HUB, B1>(a, node, rows, y);
    store_row<DP>(a.y_next, node, y);
    return s;
}

} // namespace annembed
