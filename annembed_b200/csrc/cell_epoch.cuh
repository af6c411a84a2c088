// cell_epoch.cuh -- K4, cell-resident form: ONE kernel runs several consecutive mini-epochs ("sub-steps") of a CELL of
// up to CellCfg<DP>::CELL nodes with the cell's rows of the layout held in shared memory.
//
// Why.  A mini-epoch of the tiled pair k_epoch_out / k_epoch_in is latency-bound, not bandwidth-bound: every warp walks a
// chain of dependent global gathers (row -> partner row / firing-count slot -> source row) for a few hundred
// instructions of arithmetic (profiles/r02_*), and the per-mini-epoch sweep of rows, in-edge slots and layout is paid
// again at every mini-epoch.  The internal numbering (build_relabelling) packs graph-local nodes into cells, so most
// positive edges have both ends in one cell.  A CTA that owns a cell keeps its positions on chip:
//   * partner rows (y_j of an out-edge, the source row of an in-edge) of the same cell are shared-memory reads of the
//     CURRENT sub-step's positions -- no global gather, and no staleness beyond one sub-step, exactly as in the
//     per-mini-epoch kernels;
//   * the firing counts pushed by the sources are a byte map in shared memory;
//   * partner rows in OTHER cells and all negatives are read from the global snapshot written by the previous launch;
//     in-edges whose source lies in another cell are replayed from the in-edge record (as the multi-rank kernel does).
// Semantics per sub-step are those of one mini-epoch of k_epoch_out + k_epoch_in (same draws, same order of
// application): with one sub-step per launch the two paths produce identical bits (tests/test_gpu_parity.py).
// Deterministic, no atomics on the layout; the cells do not depend on the number of ranks.
#pragma once
#include "sgd_core.cuh"

namespace annembed {

// dynamic shared memory of k_cell_epochs; the device functions below address it by byte offset so that every access is
// a shared-space instruction (LDS/STS) even though the two position buffers swap roles every sub-step
extern __shared__ __align__(16) unsigned char cell_smem[];
__device__ __forceinline__ float *cell_f(uint32_t off) { return reinterpret_cast<float *>(cell_smem + off); }

template <int DP>
struct CellCfg {
    static constexpr int CELL = DP <= 2 ? 4096 : 2048;            // nodes per cell: 2 x 32 KB of positions in shared memory
    static constexpr int QCAP = 64;                                // ring of fired in-edges per warp
    static constexpr int RING_WORDS = 2 * QCAP + 32 * (1 + DP);    // q, count, running composite per owner lane
    static constexpr int Y_BYTES = 2 * CELL * DP * 4;
    static constexpr int REL_BYTES = (CELL + 32) * 4;              // first in-edge slot of every node, relative to the cell's
};
#define ANNEMBED_CELL_MAX_SUBSTEPS 128
// Rows of at most 8 neighbours: every warp owns a staging buffer for the static data of ONE tile ({neighbour, cumulative
// probability} rows, in-edge slots of the out-edges, 1/s^2) that a TMA bulk copy (cp.async.bulk, completion on an
// mbarrier) fills for the warp's NEXT tile while it works on the current one: the dependent global loads at the head
// of every tile, which bound the kernel at 24 resident warps, disappear.
template <int KP>
struct CellStage {
    static constexpr bool ON = KP <= 8;
    static constexpr int ROW_B = 32 * KP * 8, ER_B = 32 * KP * 4, S2_B = 128;
    static constexpr int BYTES = ON ? ROW_B + ER_B + S2_B : 0;
};
__device__ __forceinline__ uint32_t cell_saddr(uint32_t off) { return (uint32_t)__cvta_generic_to_shared(cell_smem) + off; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "CELL_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra CELL_DONE;\n"
                 "bra CELL_WAIT;\n"
                 "CELL_DONE:\n"
                 "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct CellArgs {
    EpochArgs e;                         // e.epoch = global index of the first sub-step; e.in_ptr = in_ptr_all (index = node id)
    const uint32_t *__restrict__ cell_start;   // [n_cells + 1] first node of every cell (multiples of 32)
    uint32_t cell_lo;                    // first cell of this rank
    uint32_t substeps;                   // sub-steps (mini-epochs) per launch
    const uint64_t *__restrict__ ext_ptr;      // [n_cells + 1] range in ext_slot of the cell's in-edges whose source is outside it
    const uint32_t *__restrict__ ext_slot;     // their positions in the transposed index, minus in_base
};

// partner / source rows: the current positions of the cell (shared memory) or the global snapshot
template <int DP>
__device__ __forceinline__ void load_row_cell(uint32_t ycur, uint32_t c0, uint32_t csize,
                                              const float *__restrict__ y_snap, uint32_t idx, float (&v)[DP])
{
    const uint32_t loc = idx - c0;
    if (loc < csize) load_row<DP>(cell_f(ycur), loc, v);
    else load_row<DP>(y_snap, idx, v);
}

// ---- phase A of one tile (lane = node): the node's own firings --------------------------------------------------
template <int DP, bool HUB, int KP>
__device__ __forceinline__ unsigned int cell_phase_a(const EpochArgs &a, uint32_t epoch, uint32_t ukey, uint32_t ycur,
                                                     uint32_t ymid, uint32_t fmap, uint32_t c0, uint32_t csize,
                                                     uint64_t Q0, uint32_t n_in, uint32_t n0, int lane,
                                                     uint32_t stage, uint32_t bar, uint32_t &parity, const void *next_rows,
                                                     const void *next_erank, const void *next_s2)
{
    using ST = CellStage<KP>;
    const uint32_t node = n0 + (uint32_t)lane;
    const bool valid = node - c0 < csize;
    float y[DP], g[DP];
    uint32_t rc[KP];
    float cm[KP];
    uint32_t chb[(KP + 3) / 4];
    float inv_s2 = 1.0f;
    int T = 0;
    {
        if constexpr (ST::ON) {                                    // the tile's static data was staged by the TMA copy
            mbar_wait(cell_saddr(bar), parity);
            parity ^= 1u;
            const uint4 *rp = reinterpret_cast<const uint4 *>(cell_smem + stage) + (size_t)lane * (KP / 2);
#pragma unroll
            for (int h = 0; h < KP / 2; h++) {
                const uint4 t = rp[h];
                rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
                rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
            }
        } else {
            const uint4 *rp = reinterpret_cast<const uint4 *>(a.rowpack + (size_t)(valid ? node : n0) * KP);
#pragma unroll
            for (int h = 0; h < KP / 2; h++) {
                const uint4 t = __ldg(rp + h);
                rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
                rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
            }
        }
#pragma unroll
        for (int w = 0; w < (KP + 3) / 4; w++) chb[w] = 0x7f7f7f7fu;
#pragma unroll
        for (int c = 0; c < DP; c++) y[c] = 0.0f;
        if (valid) {
            load_row<DP>(cell_f(ycur), node - c0, y);
            if constexpr (ST::ON) inv_s2 = cell_f(stage + ST::ROW_B + ST::ER_B)[lane];
            else inv_s2 = __ldg(a.inv_s2 + node);
            const float u = node_uniform(node, ukey);
            int prev = 0;
#pragma unroll
            for (int m = 0; m < KP; m++) {
                const int ch = cum_ceil(a.kappa, cm[m], u);
                chb[m >> 2] = (chb[m >> 2] & ~(0xffu << (8 * (m & 3)))) | ((uint32_t)ch << (8 * (m & 3)));
                if (ch > prev) {                                   // destination in this cell: push the count to its in-edge slot
                    uint32_t er;
                    if constexpr (ST::ON) er = reinterpret_cast<const uint32_t *>(cell_smem + stage + ST::ROW_B)[lane * KP + m];
                    else er = __ldg(a.erank + (size_t)node * KP + m);
                    const uint64_t loc = (uint64_t)er - Q0;
                    if (loc < (uint64_t)n_in) cell_smem[fmap + (uint32_t)loc] = (unsigned char)(ch - prev);
                }
                prev = ch;
            }
            T = prev;
        } else {
#pragma unroll
            for (int m = 0; m < KP; m++) rc[m] = ANNEMBED_NO_NODE;
        }
        if constexpr (ST::ON) {                                    // the staging buffer is free: fetch the warp's next tile
            __syncwarp();
            if (lane == 0 && next_rows != nullptr) {
                fence_proxy_async();
                mbar_expect_tx(cell_saddr(bar), ST::BYTES);
                tma_load_1d(cell_saddr(stage), next_rows, ST::ROW_B, cell_saddr(bar));
                tma_load_1d(cell_saddr(stage + ST::ROW_B), next_erank, ST::ER_B, cell_saddr(bar));
                tma_load_1d(cell_saddr(stage + ST::ROW_B + ST::ER_B), next_s2, ST::S2_B, cell_saddr(bar));
            }
        }
    }
    const uint32_t nkey = neg_stream_key<HUB>(a, node);
    uint32_t id_lo = node, id_hi = node;
#pragma unroll
    for (int m = 0; m < KP; m++) {
        const uint32_t v = rc[m] == ANNEMBED_NO_NODE ? node : rc[m];
        id_lo = min(id_lo, v); id_hi = max(id_hi, v);
    }
    const uint32_t id_span = id_hi - id_lo;
    auto rejected = [&](uint32_t kk) -> bool {
        bool r = false;
        if (kk - id_lo <= id_span) {
            r = (kk == node);
#pragma unroll
            for (int mm = 0; mm < KP; mm++) r |= (kk == rc[mm]);
        }
        return r;
    };
    auto edge = [&](int m, uint32_t &j, float &P_lo, float &P_hi) {
        j = rc[0]; P_hi = cm[0]; P_lo = 0.0f;
#pragma unroll
        for (int mm = 1; mm < KP; mm++) {
            const bool t = m >= mm;
            j = t ? rc[mm] : j; P_hi = t ? cm[mm] : P_hi; P_lo = t ? cm[mm - 1] : P_lo;
        }
    };
    // Same Philox block sharing inside aligned groups of 4 lanes as k_epoch_out (tiles start at multiples of 32).
    struct Pre {
        int m;
        float pe;
        unsigned use;
        float yj[DP], yk[ANNEMBED_NB_NEG][DP];
    };
    const int Tmax = __reduce_max_sync(0xffffffffu, T);
    const uint32_t r4 = (uint32_t)lane & 3u;
    Philox4 blk;
    blk.x = blk.y = blk.z = blk.w = 0u;
    Philox4 An;
    An.x = An.y = An.z = An.w = 0u;
    auto fetch = [&](int s) {        // executed by the whole warp, s warp-uniform
        if ((s & 3) == 0)             // lane r of an aligned group computes the block of firing s + r (shared stream)
            blk = philox4x32_10(nkey, (uint32_t)s + r4, epoch, 1u, a.k0, a.k1);
        const int src = (lane & ~3) + (s & 3);
        An.x = __shfl_sync(0xffffffffu, blk.x, src); An.y = __shfl_sync(0xffffffffu, blk.y, src);
        An.z = __shfl_sync(0xffffffffu, blk.z, src); An.w = __shfl_sync(0xffffffffu, blk.w, src);
    };
    auto prepare = [&](int s, Pre &P) {
        const int m = edge_of_firing<KP>(chb, s);
        P.m = m;
        uint32_t j; float P_lo, P_hi;
        edge(m, j, P_lo, P_hi);
        P.pe = F_SUB(P_hi, P_lo);
        load_row_cell<DP>(ycur, c0, csize, a.y_snap, j, P.yj);
        uint32_t negs[ANNEMBED_NB_NEG];
        draw_negatives_v2<HUB>(a, epoch, node, (uint32_t)s, An, rejected, negs);
        P.use = 0;
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
            const bool ok = negs[q] != ANNEMBED_NO_NODE;
            P.use |= ok ? (1u << q) : 0u;
            load_row<DP>(a.y_snap, ok ? negs[q] : node, P.yk[q]);
        }
    };
    int m_last = -1;
    float yl[DP];
#pragma unroll
    for (int c = 0; c < DP; c++) yl[c] = 0.0f;
    auto apply = [&](Pre &P) {
        if (P.m == m_last) {
#pragma unroll
            for (int c = 0; c < DP; c++) P.yj[c] = yl[c];
        }
#pragma unroll
        for (int c = 0; c < DP; c++) g[c] = 0.0f;
        attract<DP, true>(y, P.yj, g, P.pe, inv_s2, a.K);
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) repulse<DP, true>(y, P.yk[q], g, inv_s2, a.K, (P.use >> q) & 1u);
        m_last = P.m;
#pragma unroll
        for (int c = 0; c < DP; c++) yl[c] = P.yj[c];
    };
    Pre PA, PB;
    PA.m = PB.m = -1; PA.pe = PB.pe = 0.0f; PA.use = PB.use = 0u;
    if (Tmax > 0) { fetch(0); if (T > 0) prepare(0, PA); }
    for (int s = 0; s < Tmax; s += 2) {
        if (s + 1 < Tmax) { fetch(s + 1); if (s + 1 < T) prepare(s + 1, PB); }
        if (s < T) apply(PA);
        if (s + 2 < Tmax) { fetch(s + 2); if (s + 2 < T) prepare(s + 2, PA); }
        if (s + 1 < T) apply(PB);
    }
    if (valid) store_row<DP>(cell_f(ymid), node - c0, y);
    return (unsigned int)T;
}

// ---- phase A, event form (kappa <= 1: a node fires at most once per sub-step) -------------------------------------------
// At the fine levels of the schedule fewer than one node in two fires in a sub-step, so a warp that walks its nodes lane
// by lane runs the 300 instructions of a firing for a half-empty warp, and pays the row handling for nodes that do
// nothing.  Here the warp first DECIDES for a batch of 32 or 64 nodes (one hash and one compare per node: the row is
// not needed for that), compacts the firing nodes into a list in shared memory (ballot + popc, node order), and then
// runs the firings with lane = list entry.  Same draws, same arithmetic, same result as cell_phase_a.
template <int DP, bool HUB, int KP>
__device__ __forceinline__ unsigned int cell_phase_a_events(const EpochArgs &a, uint32_t epoch, uint32_t ukey, uint32_t ycur, uint32_t ymid,
                                                            uint32_t fmap, uint32_t c0, uint32_t csize, uint64_t Q0, uint32_t n_in,
                                                            uint32_t n0, uint32_t nb, uint32_t list, int lane)
{
    unsigned short *l_id = reinterpret_cast<unsigned short *>(cell_smem + list);          // [128] node - n0
    float *l_u = reinterpret_cast<float *>(cell_smem + list + 256);                        // [128] the node's uniform
    uint32_t cnt = 0;
    for (uint32_t g = 0; g < nb; g += 32u) {
        const uint32_t i = g + (uint32_t)lane;
        const bool valid = i < nb;
        float u = 2.0f;
        bool fire = false;
        if (valid) {
            u = node_uniform(n0 + i, ukey);
            fire = cum_ceil(a.kappa, 1.0f, u) > 0;               // the row's last cumulative probability is exactly 1
            if (!fire) {                                          // nothing of its own: the position carries over to phase B
                float y[DP];
                load_row<DP>(cell_f(ycur), n0 - c0 + i, y);
                store_row<DP>(cell_f(ymid), n0 - c0 + i, y);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, fire);
        if (fire) {
            const uint32_t e = cnt + (uint32_t)__popc(m & ((1u << lane) - 1u));
            l_id[e] = (unsigned short)i; l_u[e] = u;
        }
        cnt += (uint32_t)__popc(m);
    }
    __syncwarp();
    for (uint32_t e0 = 0; e0 < cnt; e0 += 32u) {
        if (e0 + (uint32_t)lane >= cnt) continue;                 // no warp-wide operation below
        const uint32_t node = n0 + l_id[e0 + lane];
        const float u = l_u[e0 + lane];
        uint32_t rc[KP];
        float cm[KP];
        const uint4 *rp = reinterpret_cast<const uint4 *>(a.rowpack + (size_t)node * KP);
#pragma unroll
        for (int h = 0; h < KP / 2; h++) {
            const uint4 t = __ldg(rp + h);
            rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
            rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
        }
        const float inv_s2 = __ldg(a.inv_s2 + node);
        float y[DP], g[DP], yj[DP];
        load_row<DP>(cell_f(ycur), node - c0, y);
        // the edge the single sample point lands on: the first whose count ceil(kappa P_m - u) is 1
        int m = 0;
        uint32_t j = rc[0];
        float P_lo = 0.0f, P_hi = cm[0];
        {
            bool found = cum_ceil(a.kappa, cm[0], u) > 0;
#pragma unroll
            for (int mm = 1; mm < KP; mm++) {
                const bool here = !found && cum_ceil(a.kappa, cm[mm], u) > 0;
                m = here ? mm : m; j = here ? rc[mm] : j; P_hi = here ? cm[mm] : P_hi; P_lo = here ? cm[mm - 1] : P_lo;
                found |= here;
            }
        }
        {   // destination in this cell: push the count to its in-edge slot
            const uint64_t loc = (uint64_t)__ldg(a.erank + (size_t)node * KP + m) - Q0;
            if (loc < (uint64_t)n_in) atomicOr(reinterpret_cast<uint32_t *>(cell_smem + fmap) + ((uint32_t)loc >> 5), 1u << ((uint32_t)loc & 31u));
        }
        load_row_cell<DP>(ycur, c0, csize, a.y_snap, j, yj);
        uint32_t id_lo = node, id_hi = node;
#pragma unroll
        for (int mm = 0; mm < KP; mm++) {
            const uint32_t v = rc[mm] == ANNEMBED_NO_NODE ? node : rc[mm];
            id_lo = min(id_lo, v); id_hi = max(id_hi, v);
        }
        const uint32_t id_span = id_hi - id_lo;
        auto rejected = [&](uint32_t kk) -> bool {
            bool r = false;
            if (kk - id_lo <= id_span) {
                r = (kk == node);
#pragma unroll
                for (int mm = 0; mm < KP; mm++) r |= (kk == rc[mm]);
            }
            return r;
        };
        const Philox4 An = philox4x32_10(neg_stream_key<HUB>(a, node), 0u, epoch, 1u, a.k0, a.k1);
        uint32_t negs[ANNEMBED_NB_NEG];
        draw_negatives_v2<HUB>(a, epoch, node, 0u, An, rejected, negs);
        float yk[ANNEMBED_NB_NEG][DP];
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) load_row<DP>(a.y_snap, negs[q] != ANNEMBED_NO_NODE ? negs[q] : node, yk[q]);
#pragma unroll
        for (int c = 0; c < DP; c++) g[c] = 0.0f;
        attract<DP, true>(y, yj, g, F_SUB(P_hi, P_lo), inv_s2, a.K);
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) repulse<DP, true>(y, yk[q], g, inv_s2, a.K, negs[q] != ANNEMBED_NO_NODE);
        store_row<DP>(cell_f(ymid), node - c0, y);
    }
    __syncwarp();
    return cnt;
}

// ---- phase B of one tile (lane = in-edge slot): the moves received through the in-edges -------------------------------
// Same machinery as InWarp (annembed_cuda.cu): fired slots are compacted into a ring, a full warp of them is turned
// into affine maps y -> (1+A) y - A y_src evaluated at the owner's position after phase A, composed per owner by a
// segmented scan in slot order.
template <int DP>
struct CellInWarp {
    static constexpr int QCAP = CellCfg<DP>::QCAP;
    const EpochArgs &a;
    uint32_t ycur;                 // byte offsets in cell_smem
    uint32_t c0, csize;
    int lane;
    uint32_t ring;                 // q[QCAP] | count[QCAP] | alpha[32] | beta[DP][32]
    float y[DP];
    uint32_t rel_lo;
    const uint4 *rec0;             // record of the tile's first in-edge
    uint32_t q_head = 0;
    int q_n = 0;

    __device__ __forceinline__ CellInWarp(const EpochArgs &a_, uint32_t ycur_, uint32_t c0_, uint32_t csize_, uint32_t ring_, int lane_)
        : a(a_), ycur(ycur_), c0(c0_), csize(csize_), lane(lane_), ring(ring_) {}
    __device__ __forceinline__ uint32_t &q_q(uint32_t e) const { return reinterpret_cast<uint32_t *>(cell_smem + ring)[e]; }
    __device__ __forceinline__ uint32_t &q_c(uint32_t e) const { return reinterpret_cast<uint32_t *>(cell_smem + ring)[QCAP + e]; }
    __device__ __forceinline__ float &t_alpha(uint32_t o) const { return cell_f(ring)[2 * QCAP + o]; }
    __device__ __forceinline__ float &t_beta(uint32_t cc, uint32_t o) const { return cell_f(ring)[2 * QCAP + 32 + cc * 32 + o]; }

    __device__ __forceinline__ void dense(uint32_t head, int cnt)
    {
        const bool act = lane < cnt;
        const uint32_t e = (head + (uint32_t)lane) & (QCAP - 1);
        uint32_t q = 0;
        int c = 0;
        if (act) { q = q_q(e); c = (int)q_c(e); }
        dense_regs(act, q, c);
    }

    // lane = fired in-edge (slot q of the tile, c firings), in slot order; inactive lanes at the end
    __device__ __forceinline__ void dense_regs(bool act, uint32_t q, int c)
    {
        uint32_t own = 32u + (uint32_t)lane;
        float alpha = 1.0f, beta[DP];
#pragma unroll
        for (int cc = 0; cc < DP; cc++) beta[cc] = 0.0f;
        float pe = 0.0f, is2 = 0.0f, ys[DP];
#pragma unroll
        for (int cc = 0; cc < DP; cc++) ys[cc] = 0.0f;
        if (act) {
            const uint4 r = __ldg(rec0 + q);
            pe = F_SUB(__uint_as_float(r.z), __uint_as_float(r.y)); is2 = __uint_as_float(r.w);
            load_row_cell<DP>(ycur, c0, csize, a.y_snap, r.x, ys);
        }
        {
            int o = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, rel_lo, o + step);
                o += (v <= q) ? step : 0;
            }
            if (act) own = (uint32_t)o;
        }
        float yr[DP];
#pragma unroll
        for (int cc = 0; cc < DP; cc++) yr[cc] = __shfl_sync(0xffffffffu, y[cc], (int)(own & 31u));
        if (act) {
            const float A = in_edge_factor(attract_coeff<true>(sqdist<DP>(yr, ys), pe, is2, a.K), c);
            alpha = F_ADD(1.0f, A);
#pragma unroll
            for (int cc = 0; cc < DP; cc++) beta[cc] = F_MUL(-A, ys[cc]);
        }
        const uint32_t prev_own = __shfl_up_sync(0xffffffffu, own, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || own != prev_own);
        const int seg_lo = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
        const int max_len = __reduce_max_sync(0xffffffffu, act ? lane - seg_lo + 1 : 0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (d >= max_len) break;
            const float ap = __shfl_up_sync(0xffffffffu, alpha, d);
            float bp[DP];
#pragma unroll
            for (int cc = 0; cc < DP; cc++) bp[cc] = __shfl_up_sync(0xffffffffu, beta[cc], d);
            if (lane - d >= seg_lo) {
#pragma unroll
                for (int cc = 0; cc < DP; cc++) beta[cc] = F_FMA(alpha, bp[cc], beta[cc]);
                alpha = F_MUL(alpha, ap);
            }
        }
        const bool tail = act && (lane == 31 || ((heads >> (lane + 1)) & 1u));
        if (tail) {
            const float at = t_alpha(own);
#pragma unroll
            for (int cc = 0; cc < DP; cc++) t_beta(cc, own) = F_FMA(alpha, t_beta(cc, own), beta[cc]);
            t_alpha(own) = F_MUL(alpha, at);
        }
        __syncwarp();
    }

    __device__ __forceinline__ void push(int c, uint32_t q)
    {
        const unsigned fired = __ballot_sync(0xffffffffu, c > 0);
        if (fired == 0u) return;
        if (c > 0) {
            const uint32_t e = (q_head + (uint32_t)q_n + (uint32_t)__popc(fired & ((1u << lane) - 1u))) & (QCAP - 1);
            q_q(e) = q; q_c(e) = (uint32_t)c;
        }
        q_n += __popc(fired);
        __syncwarp();
        if (q_n >= 32) {
            dense(q_head, 32);
            q_head = (q_head + 32u) & (QCAP - 1);
            q_n -= 32;
        }
    }
};

template <int DP>
__device__ __forceinline__ void cell_phase_b(const EpochArgs &a, uint32_t ycur, uint32_t ymid, uint32_t fmap, uint32_t ring, uint32_t rel,
                                             uint32_t c0, uint32_t csize, uint64_t Q0, uint32_t n0, int lane)
{
    constexpr int RC = 8;
    const uint32_t node = n0 + (uint32_t)lane;
    const bool valid = node - c0 < csize;
    const uint32_t nvalid = min(32u, c0 + csize - n0);
    CellInWarp<DP> W(a, ycur, c0, csize, ring, lane);
#pragma unroll
    for (int c = 0; c < DP; c++) W.y[c] = 0.0f;
    if (valid) load_row<DP>(cell_f(ymid), node - c0, W.y);
    const uint32_t *relp = reinterpret_cast<const uint32_t *>(cell_smem + rel) + (n0 - c0);   // slots relative to the cell's first
    const uint32_t my_q0 = relp[valid ? lane : 0];
    const uint32_t Q1t = relp[nvalid];
    const uint32_t Q0t = relp[0];
    W.rel_lo = valid ? my_q0 - Q0t : 0xffffffffu;
    const uint32_t n_in = Q1t - Q0t;
    if (n_in == 0) return;                                     // ymid already holds the result
    W.rec0 = a.in_rec + (Q0 + Q0t - a.in_base);
    unsigned char *fb = cell_smem + fmap + Q0t;
    W.t_alpha(lane) = 1.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) W.t_beta(c, lane) = 0.0f;
    __syncwarp();
    for (uint32_t base = 0; base < n_in; base += 32u * RC) {
        const uint32_t cnt = min(32u * RC, n_in - base);
        uint32_t fl[RC];
#pragma unroll
        for (int r = 0; r < RC; r++) {
            const uint32_t i = 32u * r + lane;
            fl[r] = i < cnt ? (uint32_t)fb[base + i] : 0u;
        }
#pragma unroll
        for (int r = 0; r < RC; r++)
            if (fl[r]) fb[base + 32u * r + lane] = 0;
#pragma unroll
        for (int r = 0; r < RC; r++) {
            if (32u * r >= cnt) break;
            W.push((int)fl[r], base + 32u * r + lane);
        }
    }
    if (W.q_n > 0) W.dense(W.q_head, W.q_n);
    if (valid) {
        const float at = W.t_alpha(lane);
#pragma unroll
        for (int c = 0; c < DP; c++) W.y[c] = F_FMA(at, W.y[c], W.t_beta(c, lane));
        store_row<DP>(cell_f(ymid), node - c0, W.y);
    }
    __syncwarp();
}

// ---- phase B, event form: with kappa <= 1 an edge fires at most once per sub-step, so the firing map is a BITMAP (one
// bit per in-edge slot of the cell, set with shared-memory atomicOr by the sources).  A tile's ~200 slots are 7 words:
// one load per lane, a warp scan of the popcounts, and lane k picks the k-th set bit with __fns -- the fired in-edges
// come out in slot order, 32 at a time, straight into the dense stage (no byte sweep, no ring).
template <int DP>
__device__ __forceinline__ void cell_phase_b_bits(const EpochArgs &a, uint32_t ycur, uint32_t ymid, uint32_t fmap, uint32_t ring, uint32_t rel,
                                                  uint32_t c0, uint32_t csize, uint64_t Q0, uint32_t n0, int lane)
{
    const uint32_t node = n0 + (uint32_t)lane;
    const bool valid = node - c0 < csize;
    const uint32_t nvalid = min(32u, c0 + csize - n0);
    CellInWarp<DP> W(a, ycur, c0, csize, ring, lane);
#pragma unroll
    for (int c = 0; c < DP; c++) W.y[c] = 0.0f;
    if (valid) load_row<DP>(cell_f(ymid), node - c0, W.y);
    const uint32_t *relp = reinterpret_cast<const uint32_t *>(cell_smem + rel) + (n0 - c0);
    const uint32_t my_q0 = relp[valid ? lane : 0];
    const uint32_t Q1t = relp[nvalid];
    const uint32_t Q0t = relp[0];
    W.rel_lo = valid ? my_q0 - Q0t : 0xffffffffu;
    const uint32_t n_in = Q1t - Q0t;
    if (n_in == 0) return;
    W.rec0 = a.in_rec + (Q0 + Q0t - a.in_base);
    uint32_t *bm = reinterpret_cast<uint32_t *>(cell_smem + fmap);
    W.t_alpha(lane) = 1.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) W.t_beta(c, lane) = 0.0f;
    __syncwarp();
    const uint32_t w_first = Q0t >> 5, w_last = (Q1t - 1u) >> 5;          // words that hold the tile's bits
    bool any = false;
    for (uint32_t wb = w_first; wb <= w_last; wb += 32u) {                  // 32 words = 1024 slots per round (hubs: several)
        const uint32_t w = wb + (uint32_t)lane;
        uint32_t bits = 0;
        if (w <= w_last) {
            bits = bm[w];
            if (w == w_first) bits &= 0xffffffffu << (Q0t & 31u);           // bits below / above belong to the neighbouring tiles
            if (w == w_last) bits &= 0xffffffffu >> (31u - ((Q1t - 1u) & 31u));
            if (bits) atomicAnd(bm + w, ~bits);                             // consumed
        }
        // inclusive scan of the popcounts over the lanes
        const uint32_t pc = (uint32_t)__popc(bits);
        uint32_t incl = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t k0 = 0; k0 < tot; k0 += 32u) {                        // entries k0 .. k0+31 of this round, in slot order
            const uint32_t k = k0 + (uint32_t)lane;
            const bool act = k < tot;
            // the word that holds entry k: the first lane whose inclusive count exceeds k (binary search by shuffles)
            int o = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, incl, o + step - 1);
                o += (v <= k) ? step : 0;
            }
            o = min(o, 31);
            const uint32_t wbits = __shfl_sync(0xffffffffu, bits, o);
            const uint32_t before = __shfl_sync(0xffffffffu, incl - pc, o);
            uint32_t q = 0;
            if (act) q = ((wb + (uint32_t)o) << 5) + __fns(wbits, 0u, (int)(k - before) + 1) - Q0t;
            W.dense_regs(act, q, act ? 1 : 0);
            any = true;
        }
    }
    if (valid && any) {
        const float at = W.t_alpha(lane);
#pragma unroll
        for (int c = 0; c < DP; c++) W.y[c] = F_FMA(at, W.y[c], W.t_beta(c, lane));
        store_row<DP>(cell_f(ymid), node - c0, W.y);
    }
    __syncwarp();
}

#ifndef ANNEMBED_CELL_THREADS
#define ANNEMBED_CELL_THREADS 768
#endif
#ifndef ANNEMBED_CELL_MINB
#define ANNEMBED_CELL_MINB 1
#endif

// dynamic shared memory: positions (2 buffers) | in-edge offsets | per-warp rings | per-warp staging buffers | per-warp
// mbarriers | firing-count byte map (map_bytes, a multiple of 16)
template <int DP, int KP>
__host__ __device__ constexpr uint32_t cell_smem_fixed_bytes(int nwarps)
{
    return (uint32_t)(CellCfg<DP>::Y_BYTES + CellCfg<DP>::REL_BYTES + nwarps * (CellCfg<DP>::RING_WORDS * 4 + CellStage<KP>::BYTES + 8));
}
#ifndef ANNEMBED_CELL_THREADS_EVENTS
#define ANNEMBED_CELL_THREADS_EVENTS 512
#endif
// EVENTS: every sub-step of the launch has kappa <= 1 and runs the event form of phase A (no staging buffers)
#ifndef ANNEMBED_CELL_MINB_EVENTS
#define ANNEMBED_CELL_MINB_EVENTS 2
#endif
template <int DP, bool HUB, int KP, bool EVENTS>
__global__ void __launch_bounds__(EVENTS ? ANNEMBED_CELL_THREADS_EVENTS : ANNEMBED_CELL_THREADS, EVENTS ? ANNEMBED_CELL_MINB_EVENTS : ANNEMBED_CELL_MINB)
k_cell_epochs(CellArgs A, unsigned long long *sample_counter)
{
    using CFG = CellCfg<DP>;
    using ST = CellStage<EVENTS ? 32 : KP>;                       // KP = 32: staging off
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rel = CFG::Y_BYTES;
    const uint32_t ring = rel + CFG::REL_BYTES + (uint32_t)warp * CFG::RING_WORDS * 4u;
    const uint32_t stage = rel + CFG::REL_BYTES + (uint32_t)nwarps * CFG::RING_WORDS * 4u + (uint32_t)warp * ST::BYTES;
    const uint32_t bar = rel + CFG::REL_BYTES + (uint32_t)nwarps * (CFG::RING_WORDS * 4u + ST::BYTES) + (uint32_t)warp * 8u;
    const uint32_t fmap = cell_smem_fixed_bytes<DP, EVENTS ? 32 : KP>(nwarps);

    const EpochArgs &a = A.e;
    const uint32_t cell = A.cell_lo + blockIdx.x;
    const uint32_t c0 = __ldg(A.cell_start + cell), csize = __ldg(A.cell_start + cell + 1) - c0;
    const uint64_t Q0 = __ldg(a.in_ptr + c0);
    const uint32_t n_in = (uint32_t)(__ldg(a.in_ptr + c0 + csize) - Q0);
    const uint32_t ntiles = (csize + 31u) >> 5;
    uint32_t parity = 0;
    if constexpr (ST::ON) {
        if (lane == 0) mbar_init(cell_saddr(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // the cell's rows of the snapshot (16-byte copies), its in-edge offsets, an empty byte map
        const float4 *src = reinterpret_cast<const float4 *>(a.y_snap + (size_t)c0 * DP);
        float4 *dst = reinterpret_cast<float4 *>(cell_smem);
        const uint32_t nv = csize * DP / 4;
        for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) dst[i] = __ldg(src + i);
        for (uint32_t i = nv * 4 + threadIdx.x; i < csize * DP; i += blockDim.x) cell_f(0)[i] = a.y_snap[(size_t)c0 * DP + i];
        uint32_t *relp = reinterpret_cast<uint32_t *>(cell_smem + rel);
        for (uint32_t i = threadIdx.x; i <= csize; i += blockDim.x) relp[i] = (uint32_t)(__ldg(a.in_ptr + c0 + i) - Q0);
        uint4 *fz = reinterpret_cast<uint4 *>(cell_smem + fmap);
        const uint32_t map_bytes = EVENTS ? ((n_in + 31u) >> 5) * 4u : n_in;     // bitmap / byte map
        for (uint32_t i = threadIdx.x; i < (map_bytes + 15u) / 16u; i += blockDim.x) fz[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if constexpr (ST::ON) {                                       // stage the warp's first tile
        if (lane == 0 && (uint32_t)warp < ntiles) {
            const uint32_t n0 = c0 + 32u * (uint32_t)warp;
            mbar_expect_tx(cell_saddr(bar), ST::BYTES);
            tma_load_1d(cell_saddr(stage), a.rowpack + (size_t)n0 * KP, ST::ROW_B, cell_saddr(bar));
            tma_load_1d(cell_saddr(stage + ST::ROW_B), a.erank + (size_t)n0 * KP, ST::ER_B, cell_saddr(bar));
            tma_load_1d(cell_saddr(stage + ST::ROW_B + ST::ER_B), a.inv_s2 + n0, ST::S2_B, cell_saddr(bar));
        }
    }
    uint32_t ycur = 0, ymid = CFG::CELL * DP * 4;
    const uint64_t x0 = __ldg(A.ext_ptr + cell), x1 = __ldg(A.ext_ptr + cell + 1);
    unsigned int applied = 0;
    for (uint32_t sub = 0; sub < A.substeps; sub++) {
        const uint32_t epoch = a.epoch + sub;
        const uint32_t ukey = epoch_ukey(epoch, a.k2);
        // in-edges whose source lies outside the cell: replay the source's decision into the byte map
        for (uint64_t x = x0 + threadIdx.x; x < x1; x += blockDim.x) {
            const uint32_t qq = __ldg(A.ext_slot + x);
            const uint4 rec = __ldg(a.in_rec + qq);
            const float us = node_uniform(rec.x, ukey);
            const int c = cum_ceil(a.kappa, __uint_as_float(rec.z), us) - cum_ceil(a.kappa, __uint_as_float(rec.y), us);
            if (c > 0) {
                const uint32_t loc = (uint32_t)((uint64_t)qq + a.in_base - Q0);
                if constexpr (EVENTS) atomicOr(reinterpret_cast<uint32_t *>(cell_smem + fmap) + (loc >> 5), 1u << (loc & 31u));
                else cell_smem[fmap + loc] = (unsigned char)c;
            }
        }
        if constexpr (EVENTS) {
            // batches of 32 nodes, or of 64 when fewer than half of them fire
            const uint32_t bn = a.kappa <= 0.5f ? 64u : 32u;
            for (uint32_t b0 = (uint32_t)warp * bn; b0 < csize; b0 += (uint32_t)nwarps * bn) {
                const unsigned int fired = cell_phase_a_events<DP, HUB, KP>(a, epoch, ukey, ycur, ymid, fmap, c0, csize, Q0, n_in, c0 + b0,
                                                                            min(bn, csize - b0), ring, lane);
                applied += lane == 0 ? fired : 0u;                 // warp-uniform count, summed over the lanes at the end
            }
        } else {
            for (uint32_t t = warp; t < ntiles; t += nwarps) {
                // the tile this warp works on after this one: the next of this sub-step, or its first of the next sub-step
                const uint32_t tn = t + nwarps < ntiles ? t + nwarps : (uint32_t)warp;
                const bool more = t + nwarps < ntiles || sub + 1 < A.substeps;
                const uint32_t nn = c0 + 32u * tn;
                applied += cell_phase_a<DP, HUB, KP>(a, epoch, ukey, ycur, ymid, fmap, c0, csize, Q0, n_in, c0 + 32u * t, lane, stage, bar, parity,
                                                     more ? (const void *)(a.rowpack + (size_t)nn * KP) : nullptr, a.erank + (size_t)nn * KP,
                                                     a.inv_s2 + nn);
            }
        }
        __syncthreads();
        for (uint32_t t = warp; t < ntiles; t += nwarps) {
            if constexpr (EVENTS) cell_phase_b_bits<DP>(a, ycur, ymid, fmap, ring, rel, c0, csize, Q0, c0 + 32u * t, lane);
            else cell_phase_b<DP>(a, ycur, ymid, fmap, ring, rel, c0, csize, Q0, c0 + 32u * t, lane);
        }
        __syncthreads();
        const uint32_t tmp = ycur; ycur = ymid; ymid = tmp;
    }
    {   // publish the cell: own replica, then the other ranks' replicas (peer memory over NVLink)
        const float4 *src = reinterpret_cast<const float4 *>(cell_smem + ycur);
        const float *srcf = cell_f(ycur);
        const uint32_t nv = csize * DP / 4;
        float4 *dst = reinterpret_cast<float4 *>(a.y_next + (size_t)c0 * DP);
        for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) dst[i] = src[i];
        for (uint32_t i = nv * 4 + threadIdx.x; i < csize * DP; i += blockDim.x) a.y_next[(size_t)c0 * DP + i] = srcf[i];
        for (uint32_t pr = 0; pr < a.n_peers; pr++) {
            float4 *pd = reinterpret_cast<float4 *>(a.peer_next[pr] + (size_t)c0 * DP);
            for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) pd[i] = src[i];
            for (uint32_t i = nv * 4 + threadIdx.x; i < csize * DP; i += blockDim.x) a.peer_next[pr][(size_t)c0 * DP + i] = srcf[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if (lane == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}

} // namespace annembed
