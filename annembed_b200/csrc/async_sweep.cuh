// K4, asynchronous form (the default): k_sweep_events / k_sweep_async / k_sweep_async_generic.
//
// The reference loop (gradient_iteration_threaded, embedder.rs:1311-1315 -> ce_optim_edge_shannon :1167-1302) is
// asynchronous itself: every rayon thread reads the two ends and the negatives at whatever position they have,
// computes the sample on local copies and publishes y_j (:1239) and y_i (:1301).  This kernel is that loop with the
// GPU's thread count: ONE layout buffer, every sample reads current positions (L2-coherent loads, ld.global.cg) and
// publishes its moves at once with float atomics (red.global.add: +g on y_j after the attraction, the node's own
// accumulated move at the end of its firings).  Atomic adds instead of the reference's read-modify-write under a row
// lock: no concurrent move is ever lost (the reference loses the moves that land between its read and its write).
// What a sample can miss is bounded by the samples in flight, not by a mini-epoch: there is no snapshot, no in-edge
// replay, no second kernel.
//
// Sampling is the systematic per-node sampler of sgd_core.cuh: the edge of a firing node is the one its uniform u_i(sweep)
// lands on in the cumulative row probability -- the reference's edge law (embedder.rs:858,987).  Negatives: the
// shared-sector draws of draw_negatives_v2 (uniform or hubness alias).
//
// Three things decide whether the layout statistics match the reference's (each measured against the oracle fixtures,
// DESIGN.md 4): the internal order of the nodes is RANDOM (nodes in flight together, or sharing negatives, must not be
// graph neighbours: build_relabelling); the nodes in flight are capped at a fraction of n (async_blocks); a node's
// samples must arrive irregularly, like the reference's independent draws (thinned sub-sweeps, k_sweep_events below).
//
// The result depends on the interleaving of the warps: runs are NOT bit-reproducible (neither are the reference's:
// unseeded thread-local RNG, embedder.rs:1182).  ANNEMBED_FLAG_BULK_SYNCHRONOUS selects the deterministic
// snapshot kernels instead.
#pragma once

#ifndef ANNEMBED_ASYNC_WINDOW_DIV
#define ANNEMBED_ASYNC_WINDOW_DIV 32        // layouts of dimension <= 4: at most n / 32 nodes in flight
#endif
#ifndef ANNEMBED_ASYNC_WINDOW_DIV_WIDE
#define ANNEMBED_ASYNC_WINDOW_DIV_WIDE 256  // wider layouts
#endif

// Order in which a sweep visits the 32-node tiles.  The internal numbering is local (graph neighbours sit in the same
// 4096-node cell, in random order inside it), which is what the gathers want; but tiles visited at the same time must
// NOT be neighbours, or a cell's nodes would all move against the same stale picture of each other (measured: the
// layout statistics then depend on the firings per sweep exactly like a bulk-synchronous mini-epoch).  Visit i goes to
// tile (i * mul) mod tiles, mul ~ tiles / golden ratio and coprime with tiles: the visits in flight at any moment are
// spread evenly over the whole numbering, a few tiles per cell.  Warp w visits i = w, w + W, ...: one modular add each.
struct TileOrder {
    uint32_t tiles, mul, step;       // step = (W * mul) mod tiles
    __device__ __forceinline__ uint32_t first(uint32_t w) const { return (uint32_t)(((uint64_t)w * mul) % tiles); }
    __device__ __forceinline__ uint32_t next(uint32_t t) const { const uint32_t v = t + step; return v >= tiles ? v - tiles : v; }
};

// The padded rows of the asynchronous form are stored tile-interleaved (k_rowpack, interleave = 1): 16-byte chunk h of the
// row of node i sits at chunk ((i >> 5) * KP/2 + h) * 32 + (i & 31), so that the h-th load of a warp that visits a tile
// reads 512 contiguous bytes (4 lines, 16 sectors).  With node-major rows (48 bytes apart at KP = 6) every one of the KP/2
// loads touched all 12 lines of the tile: 3 x the line requests and 2 x the L2 sectors for the same bytes (ncu: 6.9 instead
// of 5.4 sectors per sample), 8 x the line requests at KP = 16.  async_row_ptr -> chunk 0; chunk h is at rp[32 * h].
template <int KP>
__device__ __forceinline__ const uint4 *async_row_ptr(const uint2 *rowpack, uint32_t node)
{
    return reinterpret_cast<const uint4 *>(rowpack) + ((size_t)(node >> 5) * (KP / 2) * 32 + (node & 31u));
}

template <int DP>
__device__ __forceinline__ void load_row_cg(const float *Y, uint32_t idx, float (&v)[DP])
{
    const float *r = Y + (size_t)idx * DP;
    if constexpr (DP == 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2 *>(r));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < DP; c += 4) {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(r + c));
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        }
    }
}

// Several ranks (one GPU each): the layout is replicated, rank r fires the nodes [lo[r], lo[r+1]) of the internal order
// and is the owner of their rows.  A sample reads every position from the local replica; its own node's move goes to the
// local replica (the node is owned), the move of the other end y_j goes to the replica of j's OWNER -- a reduction on
// peer memory over NVLink when j lives on another rank (the ranks own graph-local parts, so that is the minority of the
// edges).  The owners' rows are copied to all replicas every few launches (annembed_cuda.cu, exchange_rows_nccl).
struct PeerMap {
    float *y[8];                         // y[r]: rank r's replica (y[rank] = the local one)
    uint32_t lo[9];                      // shard boundaries
    uint32_t nranks;
    // Select chain over the kernel parameters: NO dynamic index into y[] -- indexing the parameter array by a computed rank
    // makes the compiler copy it to local memory and read it back with an LDL before every reduction, and that load was
    // where 42 % of the stall samples of k_sweep_events sat (ncu source page, profiles/r02_k4_async_events_v4_ncu_full.json:
    // the IMAD.WIDE after `LDL.64`), on one rank too, where the result is never used.
    __device__ __forceinline__ float *owner_replica(uint32_t j, float *local) const
    {
        float *p = local;
        if (nranks > 1) {
            p = y[0];
#pragma unroll
            for (uint32_t q = 1; q < 8; q++) p = (q < nranks && j >= lo[q]) ? y[q] : p;
        }
        return p;
    }
};

// Y[idx] += v, one vector reduction per 8 / 16 bytes (fire and forget, performed by the L2)
template <int DP>
__device__ __forceinline__ void red_add_row(float *Y, uint32_t idx, const float (&v)[DP])
{
    float *r = Y + (size_t)idx * DP;
    if constexpr (DP == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(r), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
#pragma unroll
        for (int c = 0; c < DP; c += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(r + c), "f"(v[c]), "f"(v[c + 1]), "f"(v[c + 2]), "f"(v[c + 3]) : "memory");
    }
}

template <int DP, bool HUB, int KP>
__global__ void __launch_bounds__(EpochTile<DP, KP>::WARPS * 32, EpochTile<DP, KP>::MINB)
k_sweep_async(EpochArgs a, float *Y, PeerMap pm, TileOrder ord, unsigned long long *sample_counter)
{
    using TL = EpochTile<DP, KP>;
    static_assert(KP % 2 == 0, "rows are padded to an even number of entries (16-byte loads)");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned int applied = 0;
    const uint32_t w0 = blockIdx.x * TL::WARPS + wib, wtot = gridDim.x * TL::WARPS;
    uint32_t tile = ord.first(w0);
    for (uint32_t idx = w0; idx < ord.tiles; idx += wtot, tile = ord.next(tile)) {
        const uint64_t n0 = (uint64_t)a.lo + (uint64_t)tile * 32;
        const int nvalid = (int)min((uint64_t)32, (uint64_t)a.hi - n0);
        const uint32_t node = (uint32_t)n0 + lane;
        const bool valid = lane < nvalid;
        float y[DP], ystart[DP], g[DP];
        uint32_t rc[KP];
        float cm[KP];
        uint32_t chb[(KP + 3) / 4];
        float inv_s2 = 1.0f;
        int T = 0;
        {
            const uint4 *rp = async_row_ptr<KP>(a.rowpack, valid ? node : (uint32_t)n0);
#pragma unroll
            for (int h = 0; h < KP / 2; h++) {
                const uint4 t = __ldcs(rp + 32 * h);
                rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
                rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
            }
#pragma unroll
            for (int w = 0; w < (KP + 3) / 4; w++) chb[w] = 0x7f7f7f7fu;
#pragma unroll
            for (int c = 0; c < DP; c++) y[c] = 0.0f;
            if (valid) {
                load_row_cg<DP>(Y, node, y);
                inv_s2 = __ldcs(a.inv_s2 + node);
                const float u = node_uniform(node, a.ukey);
                int prev = 0;
#pragma unroll
                for (int m = 0; m < KP; m++) {
                    const int ch = cum_ceil(a.kappa, cm[m], u);    // pads have cum == 1: ch == T, they never fire
                    chb[m >> 2] = (chb[m >> 2] & ~(0xffu << (8 * (m & 3)))) | ((uint32_t)ch << (8 * (m & 3)));
                    prev = ch;
                }
                T = prev;
#ifdef ANNEMBED_ASYNC_POISSON   /* design study only: Poisson(kappa) firings per visit, every firing's edge drawn independently */
                {
                    const float v = node_uniform(node, a.ukey ^ 0x5bd1e995u);
                    float pk = __expf(-a.kappa), cdf = pk;
                    int k = 0;
                    while (v >= cdf && k < 30) { k++; pk *= a.kappa / (float)k; cdf += pk; }
                    T = k;
                }
#endif
            } else {
#pragma unroll
                for (int m = 0; m < KP; m++) rc[m] = ANNEMBED_NO_NODE;
            }
#pragma unroll
            for (int c = 0; c < DP; c++) ystart[c] = y[c];
        }
        const uint32_t nkey = neg_stream_key<HUB>(a, node);
        uint32_t id_lo = node, id_hi = node;
#pragma unroll
        for (int m = 0; m < KP; m++) {
            const uint32_t v = rc[m] == ANNEMBED_NO_NODE ? node : rc[m];
            id_lo = min(id_lo, v); id_hi = max(id_hi, v);
        }
        const uint32_t id_span = id_hi - id_lo;
        auto rejector = [&](uint32_t j) {
            return [&, j](uint32_t kk) -> bool {
                (void)j;                                           // j is one of rc[]
                bool r = false;
                if (kk - id_lo <= id_span) {
                    r = (kk == node);
#pragma unroll
                    for (int mm = 0; mm < KP; mm++) r |= (kk == rc[mm]);
                }
                return r;
            };
        };
        auto pick = [&](int s) -> int {
#ifdef ANNEMBED_ASYNC_POISSON
            const float us = node_uniform(node, a.ukey + 0x9E3779B9u * (uint32_t)(s + 1));
            int m = 0;
#pragma unroll
            for (int mm = 0; mm < KP; mm++) m += (cm[mm] <= us && rc[mm] != ANNEMBED_NO_NODE) ? 1 : 0;
            return m;
#else
            return edge_of_firing<KP>(chb, s);
#endif
        };
        auto edge = [&](int m, uint32_t &j, float &P_lo, float &P_hi) {
            j = rc[0]; P_hi = cm[0]; P_lo = 0.0f;
#pragma unroll
            for (int mm = 1; mm < KP; mm++) {
                const bool t = m >= mm;
                j = t ? rc[mm] : j; P_hi = t ? cm[mm] : P_hi; P_lo = t ? cm[mm - 1] : P_lo;
            }
        };
        if constexpr (DP <= 4) {
            // two register sets: the 6 row gathers of firings s+1 and s+2 are in flight during the arithmetic of firing s
            struct Pre {
                int m;
                uint32_t j;
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
                    blk = philox4x32_10(nkey, (uint32_t)s + r4, a.epoch, 1u, a.k0, a.k1);
                const int src = (lane & ~3) + (s & 3);
                An.x = __shfl_sync(0xffffffffu, blk.x, src); An.y = __shfl_sync(0xffffffffu, blk.y, src);
                An.z = __shfl_sync(0xffffffffu, blk.z, src); An.w = __shfl_sync(0xffffffffu, blk.w, src);
            };
            auto prepare = [&](int s, Pre &P) {
                const int m = pick(s);
                P.m = m;
                float P_lo, P_hi;
                edge(m, P.j, P_lo, P_hi);
                P.pe = F_SUB(P_hi, P_lo);
                load_row_cg<DP>(Y, P.j, P.yj);
                uint32_t negs[ANNEMBED_NB_NEG];
                draw_negatives_v2<HUB>(a, a.epoch, node, (uint32_t)s, An, rejector(P.j), negs);
                P.use = 0;
#pragma unroll
                for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
                    const bool ok = negs[q] != ANNEMBED_NO_NODE;
                    P.use |= ok ? (1u << q) : 0u;
                    load_row_cg<DP>(Y, ok ? negs[q] : node, P.yk[q]);
                }
            };
            int m_last = -1;                 // edge whose partner copy yl is live: consecutive firings of one edge continue
            float yl[DP];                    // on the local copy (the prefetched row predates this thread's own reduction)
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
                red_add_row<DP>(pm.owner_replica(P.j, Y), P.j, g);                                      // publish y_j += g (embedder.rs:1239)
#pragma unroll
                for (int q = 0; q < ANNEMBED_NB_NEG; q++) repulse<DP, true>(y, P.yk[q], g, inv_s2, a.K, (P.use >> q) & 1u);
                m_last = P.m;
#pragma unroll
                for (int c = 0; c < DP; c++) yl[c] = P.yj[c];
            };
            Pre PA, PB;
            PA.m = PB.m = -1; PA.j = PB.j = 0; PA.pe = PB.pe = 0.0f; PA.use = PB.use = 0u;
            if (Tmax > 0) { fetch(0); if (T > 0) prepare(0, PA); }
            if (Tmax > 1) { fetch(1); if (T > 1) prepare(1, PB); }
            for (int s = 0; s < Tmax; s += 2) {
                if (s < T) apply(PA);
                if (s + 2 < Tmax) { fetch(s + 2); if (s + 2 < T) prepare(s + 2, PA); }
                if (s + 1 < T) apply(PB);
                if (s + 3 < Tmax) { fetch(s + 3); if (s + 3 < T) prepare(s + 3, PB); }
            }
        } else {
            int m_prev = -1;
            uint32_t j = 0;
            float pe = 0.0f;
            float yj[DP];
            for (int s = 0; s < T; s++) {
                const int m = pick(s);
                if (m != m_prev) {
                    float P_lo, P_hi;
                    edge(m, j, P_lo, P_hi);
                    pe = F_SUB(P_hi, P_lo);
                    load_row_cg<DP>(Y, j, yj);
                    m_prev = m;
                }
                const Philox4 A = philox4x32_10(nkey, (uint32_t)s, a.epoch, 1u, a.k0, a.k1);
                uint32_t negs[ANNEMBED_NB_NEG];
                draw_negatives_v2<HUB>(a, a.epoch, node, (uint32_t)s, A, rejector(j), negs);
#pragma unroll
                for (int c = 0; c < DP; c++) g[c] = 0.0f;
                attract<DP, true>(y, yj, g, pe, inv_s2, a.K);
                red_add_row<DP>(pm.owner_replica(j, Y), j, g);
                for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
                    if (negs[q] == ANNEMBED_NO_NODE) continue;
                    float yk[DP];
                    load_row_cg<DP>(Y, negs[q], yk);
                    repulse<DP, true>(y, yk, g, inv_s2, a.K);
                }
            }
        }
        if (T > 0) {                                                             // publish the node's own move (:1301)
#pragma unroll
            for (int c = 0; c < DP; c++) g[c] = F_SUB(y[c], ystart[c]);
            red_add_row<DP>(Y, node, g);
        }
        applied += (unsigned int)T;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if (lane == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}

// thread per node, rows in global memory: rows longer than 16 neighbours, b != 1, many firings per node and sweep
template <int DP, bool HUB>
__global__ void __launch_bounds__(128) k_sweep_async_generic(EpochArgs a, float *Y, PeerMap pm, TileOrder ord, unsigned long long *sample_counter)
{
    unsigned int applied = 0;
    const uint32_t w0 = blockIdx.x * 4 + (threadIdx.x >> 5), wtot = gridDim.x * 4;
    uint32_t tile = ord.first(w0);
    for (uint32_t idx = w0; idx < ord.tiles; idx += wtot, tile = ord.next(tile)) {
        const uint64_t nn = (uint64_t)a.lo + (uint64_t)tile * 32 + (threadIdx.x & 31);
        if (nn >= a.hi) continue;
        const uint32_t node = (uint32_t)nn;
        float y[DP], ystart[DP], g[DP];
        load_row_cg<DP>(Y, node, y);
#pragma unroll
        for (int c = 0; c < DP; c++) ystart[c] = y[c];
        const float inv_s2 = a.inv_s2[node];
        const uint64_t r0 = a.row_ptr[node], r1 = a.row_ptr[node + 1];
        const float u = node_uniform(node, a.ukey);
        unsigned int s = 0;
        float P_lo = 0.0f;
        int c_lo = 0;
        for (uint64_t m = r0; m < r1; m++) {
            const float P_hi = a.cum[m];
            const int c_hi = cum_ceil(a.kappa, P_hi, u);
            const int cnt = c_hi - c_lo;
            const float pe = F_SUB(P_hi, P_lo);
            P_lo = P_hi; c_lo = c_hi;
            if (cnt <= 0) continue;
            const uint32_t j = a.col[m];
            float yj[DP];
            load_row_cg<DP>(Y, j, yj);
            const GlobalRowRejector rej{a.col, r0, r1, node, j};
            for (int f = 0; f < cnt; f++, s++) {
                const Philox4 A = philox4x32_10(neg_stream_key<HUB>(a, node), s, a.epoch, 1u, a.k0, a.k1);
                uint32_t negs[ANNEMBED_NB_NEG];
                draw_negatives_v2<HUB>(a, a.epoch, node, s, A, rej, negs);
#pragma unroll
                for (int c = 0; c < DP; c++) g[c] = 0.0f;
                attract<DP, false>(y, yj, g, pe, inv_s2, a.K);
                red_add_row<DP>(pm.owner_replica(j, Y), j, g);
                for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
                    if (negs[q] == ANNEMBED_NO_NODE) continue;
                    float yk[DP];
                    load_row_cg<DP>(Y, negs[q], yk);
                    repulse<DP, false>(y, yk, g, inv_s2, a.K);
                }
            }
        }
        if (s) {
#pragma unroll
            for (int c = 0; c < DP; c++) g[c] = F_SUB(y[c], ystart[c]);
            red_add_row<DP>(Y, node, g);
        }
        applied += s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if ((threadIdx.x & 31) == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}

// ---- kappa <= 1: at most one firing per node and sweep (the default schedule) ---------------------------------------
// Thinned sweeps.  In the reference the samples of a node arrive like a Poisson process (independent edge draws,
// embedder.rs:1182); a sweep in which EVERY node fires exactly once is too regular, and the regularity is visible in the
// layout statistics (measured, DESIGN.md 4: the deviation from the serial reference is proportional to the firing
// probability per sweep).  The default schedule therefore cuts a sweep of kappa = 1 into `subs` sub-sweeps of firing
// probability kappa' = kappa / subs: the gaps between the samples of a node are geometric, with (1 - kappa') of the
// exponential law's variance.  The thinning is done per TILE: in sub-sweep s the 32 nodes of tile t fire iff
// hash(t, s) < kappa' (tile-mates are unrelated nodes: the internal order is random), each on the edge its own uniform
// u_i(s) lands on (cumulative row probability).  A tile that does not fire costs a dozen warp-uniform instructions and no
// memory access; a tile that fires is read with full-width coalesced loads.  The sub-sweeps of one launch follow each
// other inside the persistent warps without any grid-wide synchronisation (nothing in the asynchronous form needs one).
// The work on the firing tiles is pipelined three deep across the warp's visits:
//   load    visit v+2: the tile's rows (KP/2 16-byte streaming loads per node), positions and scales;
//   gather  visit v+1: rows have arrived -> edge, negatives (rejection against the row in registers), then the 6 row
//                      gathers (y_j + 5 negatives) are issued and the row registers are dead;
//   apply   visit v  : gathers have arrived -> attraction, reduction on y_j, 5 repulsions, reduction of the node's move.
// Two register sets of gathered rows alternate (PA / PB) for layouts of dimension <= 4, one set of tile rows.  The firing
// tiles are found 32 at a time (every lane tests one candidate of the warp's visiting sequence, one ballot).
// Measured and dropped (profiles/r02_ab_tma_vs_ldg.txt): staging the tile rows through shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier, 3 slots per warp) instead of the `load` stage's register set -- 7.6 % slower on the C3
// workload: the rows are consumed once, by the thread that would have loaded them, and the slot hand-over costs more
// issue slots than the 15 registers it frees are worth at 5 resident blocks.
#ifndef ANNEMBED_EVENTS_MINB
#define ANNEMBED_EVENTS_MINB 5
#endif
template <int DP, int KP>
struct EventTile {
    static constexpr int WARPS = 4;
    static constexpr int MINB = DP <= 2 ? (KP <= 8 ? ANNEMBED_EVENTS_MINB : 4) : (DP <= 4 ? 3 : (DP <= 16 ? 2 : 1));
    static constexpr int VISITS = DP <= 4 ? 3 : 2;       // visits a warp has in flight (in-flight window of the launch)
};

__device__ __forceinline__ bool tile_fires(uint32_t tile, uint32_t ukey, float kappa)
{
    return node_uniform(tile, ukey ^ 0x68E31DA4u) < kappa;
}

template <int DP, bool HUB, int KP>
__global__ void __launch_bounds__(EventTile<DP, KP>::WARPS * 32, EventTile<DP, KP>::MINB)
k_sweep_events(EpochArgs a, float *Y, PeerMap pm, TileOrder ord, uint32_t subs, unsigned long long *sample_counter)
{
    static_assert(KP % 2 == 0, "rows are padded to an even number of entries (16-byte loads)");
    using TL = EventTile<DP, KP>;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t w0 = blockIdx.x * TL::WARPS + wib, wtot = gridDim.x * TL::WARPS;
    unsigned int applied = 0;

    struct Rows {                        // one visit, stage `load`
        uint32_t node;                   // ANNEMBED_NO_NODE: lane beyond the end of the numbering
        uint32_t rc[KP];
        float cm[KP];
        float y[DP];
        float inv_s2;
        uint32_t epoch, ukey;
    };
    struct Pre {                         // one visit, stage `gather`
        uint32_t node, j;                // node == NO_NODE: idle lane
        float pe, inv_s2;
        unsigned use;
        float y[DP], yj[DP], yk[ANNEMBED_NB_NEG][DP];
    };
    // ---- the warp's visits: sub-sweep `sub`, positions idx = w0, w0 + wtot, ... of the visiting order, firing tiles only.
    // Window of 32 candidates: lane l holds position base_idx + l * wtot, i.e. tile base_tile + l * ord.step (mod tiles).
    uint32_t sub = 0;
    uint32_t base_idx = w0;              // < tiles + 32 wtot < 2^32 (tiles <= 2^27)
    uint32_t base_tile = ord.first(w0);
    uint32_t epoch = a.epoch, ukey = a.ukey;
    const uint32_t lane_off = (uint32_t)(((uint64_t)lane * ord.step) % ord.tiles);
    const uint32_t win_step = (uint32_t)((32ull * ord.step) % ord.tiles);
    uint32_t my_tile = 0, pend = 0;
    auto scan_window = [&]() {
        uint32_t t = base_tile + lane_off;
        t = t >= ord.tiles ? t - ord.tiles : t;
        my_tile = t;
        const bool fire = base_idx + (uint32_t)lane * wtot < ord.tiles && tile_fires(t, ukey, a.kappa);
        pend = __ballot_sync(0xffffffffu, fire);
    };
    if (subs > 0) scan_window();
    auto next_tile = [&](uint32_t &tile) -> bool {
        for (;;) {
            if (pend) {
                const int b = __ffs(pend) - 1;
                pend &= pend - 1u;
                tile = __shfl_sync(0xffffffffu, my_tile, b);
                return true;
            }
            if (sub >= subs) return false;
            base_idx += 32u * wtot;
            if (base_idx >= ord.tiles) {
                if (++sub >= subs) return false;
                base_idx = w0; base_tile = ord.first(w0);
                epoch = a.epoch + sub; ukey = epoch_ukey(epoch, a.k2);
            } else {
                const uint32_t t = base_tile + win_step;
                base_tile = t >= ord.tiles ? t - ord.tiles : t;
            }
            scan_window();
        }
    };
    auto next_visit = [&](Rows &R) -> bool {
        uint32_t tile = 0;
        if (!next_tile(tile)) return false;
        const uint64_t n0 = (uint64_t)a.lo + (uint64_t)tile * 32;
        const bool valid = n0 + lane < a.hi;
        const uint32_t node = (uint32_t)n0 + (valid ? lane : 0);
        const uint4 *rp = async_row_ptr<KP>(a.rowpack, node);
#pragma unroll
        for (int h = 0; h < KP / 2; h++) {
            const uint4 t = __ldcs(rp + 32 * h);
            R.rc[2 * h] = t.x; R.cm[2 * h] = __uint_as_float(t.y);
            R.rc[2 * h + 1] = t.z; R.cm[2 * h + 1] = __uint_as_float(t.w);
        }
        load_row_cg<DP>(Y, node, R.y);
        R.inv_s2 = __ldcs(a.inv_s2 + node);
        R.node = valid ? node : ANNEMBED_NO_NODE;
        R.epoch = epoch; R.ukey = ukey;
        return true;
    };
    auto gather = [&](const Rows &R, Pre &P) {
        const bool act = R.node != ANNEMBED_NO_NODE;
        const uint32_t node = act ? R.node : a.lo;
        const float u = node_uniform(node, R.ukey);
        // the edge the node's sample point u lands on: the first edge whose cumulative probability exceeds u
        int m = 0;
#pragma unroll
        for (int mm = 0; mm < KP; mm++) m += R.cm[mm] <= u ? 1 : 0;     // == (cum_ceil(1, cm, u) <= 0); pads have cum == 1 > u
        uint32_t j = R.rc[0];
        float P_hi = R.cm[0], P_lo = 0.0f;
#pragma unroll
        for (int mm = 1; mm < KP; mm++) {
            const bool t = m >= mm;
            j = t ? R.rc[mm] : j; P_hi = t ? R.cm[mm] : P_hi; P_lo = t ? R.cm[mm - 1] : P_lo;
        }
        const bool fires = act && j != ANNEMBED_NO_NODE;
        P.node = fires ? node : ANNEMBED_NO_NODE;
        P.j = fires ? j : node;
        P.pe = F_SUB(P_hi, P_lo);
        P.inv_s2 = R.inv_s2;
#pragma unroll
        for (int c = 0; c < DP; c++) P.y[c] = R.y[c];
        load_row_cg<DP>(Y, P.j, P.yj);
        // rejection test against the row in registers (the internal order is random: no id range worth a pre-test)
        auto rejected = [&](uint32_t kk) -> bool {
            bool r = (kk == node);
#pragma unroll
            for (int mm = 0; mm < KP; mm++) r |= (kk == R.rc[mm]);
            return r;
        };
        const Philox4 A = philox4x32_10(neg_stream_key<HUB>(a, node), 0u, R.epoch, 1u, a.k0, a.k1);
        uint32_t negs[ANNEMBED_NB_NEG];
        draw_negatives_v2<HUB>(a, R.epoch, node, 0u, A, rejected, negs);
        P.use = 0;
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
            const bool ok = negs[q] != ANNEMBED_NO_NODE;
            P.use |= ok ? (1u << q) : 0u;
            load_row_cg<DP>(Y, ok ? negs[q] : node, P.yk[q]);
        }
    };
    auto apply = [&](Pre &P) {
        if (P.node == ANNEMBED_NO_NODE) return;
        float y[DP], g[DP];
#pragma unroll
        for (int c = 0; c < DP; c++) { y[c] = P.y[c]; g[c] = 0.0f; }
        attract<DP, true>(y, P.yj, g, P.pe, P.inv_s2, a.K);
        red_add_row<DP>(pm.owner_replica(P.j, Y), P.j, g);                   // publish y_j += g (embedder.rs:1239)
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) repulse<DP, true>(y, P.yk[q], g, P.inv_s2, a.K, (P.use >> q) & 1u);
#pragma unroll
        for (int c = 0; c < DP; c++) g[c] = F_SUB(y[c], P.y[c]);
        red_add_row<DP>(Y, P.node, g);                                       // publish the node's own move (:1301)
        applied++;
    };

    Rows R;
    bool more = next_visit(R);
    if constexpr (DP <= 4) {
        Pre PA, PB;
        PA.node = PB.node = ANNEMBED_NO_NODE;
        bool haveA = false, haveB = false;
        if (more) { gather(R, PA); haveA = true; more = next_visit(R); }
        while (haveA) {
            if (more) { gather(R, PB); haveB = true; more = next_visit(R); } else haveB = false;
            apply(PA);
            if (more) { gather(R, PA); haveA = true; more = next_visit(R); } else haveA = false;
            if (haveB) apply(PB);
        }
    } else {                             // wide rows: one set of gathered rows; the next visit's tile rows load during apply
        Pre PA;
        while (more) {
            gather(R, PA);
            more = next_visit(R);
            apply(PA);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if (lane == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}
