// K4, asynchronous form, layouts of dimension <= 4: k_sweep_events_cp -- k_sweep_events (async_sweep.cuh) with every global
// read of the pipeline turned into an asynchronous copy to shared memory (cp.async, LDGSTS).
//
// Why.  k_sweep_events pipelines three visits per warp (rows of visit v+2 loading, the 6 row gathers of visit v+1 in
// flight, visit v applied) in registers.  ptxas gives ALL the global loads of that loop ONE scoreboard (SB5: decoded from
// the SASS control words of every LDG, profiles/r02_sweep_events_scoreboards.txt), and a scoreboard wait releases only
// when every load counted on it has returned: the wait for the gathers of visit v also waits for the rows of visit v+2
// issued a moment ago (ncu: 52-67 % of the stall samples are long-scoreboard).  cp.async groups complete in order and
// cp.async.wait_group N leaves the N newest groups in flight -- the partial wait the register pipeline cannot express
// (ncu of this kernel: long-scoreboard 16 % of the stall samples).
// What it buys (A/B on one box, profiles/r02_ab_pipeline_and_negative_groups.txt): uniform negatives 199 vs 200 G edge
// updates/s -- nothing: with either pipeline the memory system's rate of scattered requests is the bound; hubness
// negatives 101 vs 82 G: the alias-table reads are a third dependent round of memory accesses inside the gather stage,
// which the register pipeline exposes in full.  Default for the hubness sampler, ANNEMBED_FLAG_CP_ASYNC_PIPELINE otherwise.
//
// Per thread and visit: group L = the node's padded row (KP/2 16-byte copies) + its scale; group G = 7 layout rows (own,
// partner, 5 negatives; 8 bytes each in dimension 2, 16 in dimension 3-4).  All copies are the .ca form, which goes
// through the L1 like an ordinary load: the accesses of a warp to one line are ONE request to the L2.  (The .cg / BYPASS
// form sends every thread's bytes on its own: measured, L2 throughput 81 % instead of 43 % and 25 % slower.  A row read
// through the L1 can be stale by the few microseconds a line survives in the ~90 KB the shared memory leaves it -- less than
// the samples in flight already allow; the layout statistics do not move, tests/test_gpu_fidelity.py.)
// Iteration t of a warp:  issue L(t+2); wait_group 2 -> L(t+1) has landed: edge, negatives, issue G(t+1);
// wait_group 2 -> G(t) has landed: attraction, 5 repulsions, the two reductions.  Every group has a whole iteration of
// the warp (and of the other resident warps) to arrive.  The slots are private to the thread (no barrier), laid out
// [slot][item][thread] so that all shared-memory accesses are conflict-free 8- or 16-byte accesses.
// Sampling, draws, arithmetic and publication are those of k_sweep_events (same functions of sgd_core.cuh).
#pragma once

namespace cpa {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// .ca copies go through the L1 like an ordinary load: the accesses of a warp to one line are ONE request to the L2
// (the .cg / BYPASS form sends every thread's 16 bytes on its own: measured, L2 throughput 81 % instead of 43 %)
__device__ __forceinline__ void cp16_ca(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp8_ca(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
} // namespace cpa

template <int DP, int KP>
struct EventCp {
    static_assert(DP == 2 || DP == 4, "k_sweep_events_cp: layouts of dimension <= 4");
    static constexpr int WARPS = 4, THREADS = WARPS * 32;
    static constexpr int ITEMS = 2 + ANNEMBED_NB_NEG;               // own row, partner's row, negatives
    static constexpr int H = KP / 2;                                // 16-byte units of a padded row
    static constexpr int RB = DP * 4;                               // bytes of a gathered row slot (8 or 16)
    static constexpr size_t SMEM = (size_t)2 * H * THREADS * 16 + (size_t)2 * ITEMS * THREADS * RB + 2 * THREADS * sizeof(float);
    static constexpr int FIT = (int)((227u * 1024u) / (SMEM + 1024u));
    static constexpr int MINB = FIT < 1 ? 1 : (FIT > (DP == 2 ? 5 : 4) ? (DP == 2 ? 5 : 4) : FIT);
    static constexpr int VISITS = 3;                                // visits a warp has in flight
};

// the warp's visiting sequence: sub-sweep after sub-sweep, positions w0, w0 + wtot, ... of the visiting order, firing
// tiles only, found 32 candidates at a time (one ballot) -- the scan of k_sweep_events
struct FiringTiles {
    uint32_t sub, subs, base_idx, base_tile, epoch, ukey, lane_off, win_step, my_tile, pend, w0, wtot;
    __device__ __forceinline__ void scan(const TileOrder &ord, float kappa, int lane)
    {
        uint32_t t = base_tile + lane_off;
        t = t >= ord.tiles ? t - ord.tiles : t;
        my_tile = t;
        const bool fire = base_idx + (uint32_t)lane * wtot < ord.tiles && tile_fires(t, ukey, kappa);
        pend = __ballot_sync(0xffffffffu, fire);
    }
    __device__ __forceinline__ void init(const EpochArgs &a, const TileOrder &ord, uint32_t subs_, uint32_t w0_, uint32_t wtot_, int lane)
    {
        sub = 0; subs = subs_; w0 = w0_; wtot = wtot_;
        base_idx = w0; base_tile = ord.first(w0);
        epoch = a.epoch; ukey = a.ukey;
        lane_off = (uint32_t)(((uint64_t)lane * ord.step) % ord.tiles);
        win_step = (uint32_t)((32ull * ord.step) % ord.tiles);
        my_tile = 0; pend = 0;
        if (subs > 0) scan(ord, a.kappa, lane);
    }
    __device__ __forceinline__ bool next(const EpochArgs &a, const TileOrder &ord, int lane, uint32_t &tile)
    {
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
            scan(ord, a.kappa, lane);
        }
    }
};

template <int DP, bool HUB, int KP>
__global__ void __launch_bounds__(EventCp<DP, KP>::THREADS, EventCp<DP, KP>::MINB)
k_sweep_events_cp(EpochArgs a, float *Y, PeerMap pm, TileOrder ord, uint32_t subs, unsigned long long *sample_counter)
{
    static_assert(KP % 2 == 0, "rows are padded to an even number of entries (16-byte copies)");
    using TC = EventCp<DP, KP>;
    constexpr int ITEMS = TC::ITEMS, H = TC::H, NT = TC::THREADS;
    using GRow = typename std::conditional<DP == 2, uint2, uint4>::type;     // one gathered row
    extern __shared__ uint4 sm4[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    uint4 *const sm_rows = sm4 + tid;                               // [slot][h]    -> sm_rows[(slot * H + h) * NT]
    GRow *const sm_g = reinterpret_cast<GRow *>(sm4 + 2 * H * NT) + tid;      // [slot][item] -> sm_g[(slot * ITEMS + item) * NT]
    float *const sm_inv = reinterpret_cast<float *>(reinterpret_cast<GRow *>(sm4 + 2 * H * NT) + 2 * ITEMS * NT) + tid;   // [slot] -> sm_inv[slot * NT]
    unsigned int applied = 0;

    FiringTiles ft;
    ft.init(a, ord, subs, blockIdx.x * TC::WARPS + wib, gridDim.x * TC::WARPS, lane);
    bool more = true;

    // per slot: stage `load` (Rn: node or NO_NODE for a lane beyond the end; Rv: warp-uniform, the slot holds a visit) and
    // stage `gather` (Gn: node or NO_NODE for an idle lane)
    uint32_t Rn[2], Re[2], Ru[2], Gn[2], Gj[2];
    float Gpe[2], Ginv[2];
    unsigned Guse[2];
    bool Rv[2] = {false, false}, Gv[2] = {false, false};
    Rn[0] = Rn[1] = Gn[0] = Gn[1] = ANNEMBED_NO_NODE;
    Re[0] = Re[1] = Ru[0] = Ru[1] = Gj[0] = Gj[1] = 0u;
    Gpe[0] = Gpe[1] = Ginv[0] = Ginv[1] = 0.0f;
    Guse[0] = Guse[1] = 0u;

    auto copy_row = [&](GRow *dst, uint32_t idx) {                  // row idx of the layout -> the thread's slot
        if constexpr (DP == 2) cpa::cp8_ca(dst, Y + (size_t)idx * DP);
        else cpa::cp16_ca(dst, Y + (size_t)idx * DP);
    };

    auto load = [&](auto SLOT) {                                    // group L of the next visit (an empty group when there is none)
        constexpr int S = decltype(SLOT)::value;
        uint32_t tile = 0;
        const bool ok = more && ft.next(a, ord, lane, tile);
        more = ok;
        Rv[S] = ok;
        if (ok) {
            const uint64_t n0 = (uint64_t)a.lo + (uint64_t)tile * 32;
            const bool valid = n0 + lane < a.hi;
            const uint32_t node = (uint32_t)n0 + (valid ? lane : 0);
            const uint4 *rp = async_row_ptr<KP>(a.rowpack, node);
#pragma unroll
            for (int h = 0; h < H; h++) cpa::cp16_ca(sm_rows + (S * H + h) * NT, rp + 32 * h);
            cpa::cp4(sm_inv + S * NT, a.inv_s2 + node);
            Rn[S] = valid ? node : ANNEMBED_NO_NODE;
            Re[S] = ft.epoch; Ru[S] = ft.ukey;
        }
        cpa::commit();
    };

    auto gather = [&](auto SLOT) {                                  // rows of slot S have landed -> edge, negatives, group G
        constexpr int S = decltype(SLOT)::value;
        Gv[S] = Rv[S];
        if (Rv[S]) {
            uint32_t rc[KP];
            float cm[KP];
#pragma unroll
            for (int h = 0; h < H; h++) {
                const uint4 t = sm_rows[(S * H + h) * NT];
                rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
                rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
            }
            const bool act = Rn[S] != ANNEMBED_NO_NODE;
            const uint32_t node = act ? Rn[S] : a.lo;
            const float u = node_uniform(node, Ru[S]);
            int m = 0;
#pragma unroll
            for (int mm = 0; mm < KP; mm++) m += cm[mm] <= u ? 1 : 0;   // first edge whose cumulative probability exceeds u
            uint32_t j = rc[0];
            float P_hi = cm[0], P_lo = 0.0f;
#pragma unroll
            for (int mm = 1; mm < KP; mm++) {
                const bool t = m >= mm;
                j = t ? rc[mm] : j; P_hi = t ? cm[mm] : P_hi; P_lo = t ? cm[mm - 1] : P_lo;
            }
            const bool fires = act && j != ANNEMBED_NO_NODE;
            j = fires ? j : node;
            auto rejected = [&](uint32_t kk) -> bool {
                bool r = (kk == node);
#pragma unroll
                for (int mm = 0; mm < KP; mm++) r |= (kk == rc[mm]);
                return r;
            };
            const Philox4 A = philox4x32_10(neg_stream_key<HUB>(a, node), 0u, Re[S], 1u, a.k0, a.k1);
            uint32_t negs[ANNEMBED_NB_NEG];
            draw_negatives_v2<HUB>(a, Re[S], node, 0u, A, rejected, negs);
            unsigned use = 0;
            GRow *const g0 = sm_g + (S * ITEMS) * NT;
            copy_row(g0, node);
            copy_row(g0 + NT, j);
#pragma unroll
            for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
                const bool okq = negs[q] != ANNEMBED_NO_NODE;
                use |= okq ? (1u << q) : 0u;
                copy_row(g0 + (2 + q) * NT, okq ? negs[q] : node);
            }
            Gn[S] = fires ? node : ANNEMBED_NO_NODE;
            Gj[S] = j;
            Gpe[S] = F_SUB(P_hi, P_lo);
            Ginv[S] = sm_inv[S * NT];
            Guse[S] = use;
        }
        cpa::commit();
    };

    auto apply = [&](auto SLOT) {                                   // gathered rows of slot S have landed
        constexpr int S = decltype(SLOT)::value;
        if (!Gv[S] || Gn[S] == ANNEMBED_NO_NODE) return;
        const GRow *const g0 = sm_g + (S * ITEMS) * NT;
        auto row = [&](int item, float (&v)[DP]) {
            const GRow t = g0[item * NT];
            v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y);
            if constexpr (DP == 4) { v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w); }
        };
        float y[DP], y0[DP], yj[DP], g[DP];
        row(0, y);
        row(1, yj);
#pragma unroll
        for (int c = 0; c < DP; c++) { y0[c] = y[c]; g[c] = 0.0f; }
        attract<DP, true>(y, yj, g, Gpe[S], Ginv[S], a.K);
        red_add_row<DP>(pm.owner_replica(Gj[S], Y), Gj[S], g);         // publish y_j += g (embedder.rs:1239)
#pragma unroll
        for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
            float yk[DP];
            row(2 + q, yk);
            repulse<DP, true>(y, yk, g, Ginv[S], a.K, (Guse[S] >> q) & 1u);
        }
#pragma unroll
        for (int c = 0; c < DP; c++) g[c] = F_SUB(y[c], y0[c]);
        red_add_row<DP>(Y, Gn[S], g);                                  // publish the node's own move (:1301)
        applied++;
    };

    using S0 = std::integral_constant<int, 0>;
    using S1 = std::integral_constant<int, 1>;
    // prologue: L(0), L(1); G(0)
    load(S0{});
    load(S1{});
    cpa::wait<1>();
    gather(S0{});
    // iteration t (slot s = t & 1): L(t+2) -> slot s; G(t+1) from slot s^1; apply(t) from slot s
    while (Gv[0]) {
        load(S0{});
        cpa::wait<2>();
        gather(S1{});
        cpa::wait<2>();
        apply(S0{});
        if (!Gv[1]) break;
        load(S1{});
        cpa::wait<2>();
        gather(S0{});
        cpa::wait<2>();
        apply(S1{});
    }
    cpa::wait<0>();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if (lane == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}
