// Host-side construction of the hubness sampler's grouped alias tables (plain C++ + CUDA vector types; used by
// annembed_cuda.cu build_sector_alias and, for the CPU tests, by tests/hostsim).  `weight(i)` = sampling weight of node i of
// the INTERNAL numbering, as a double.  The law the tables encode is checked exactly in tests/test_host.py
// (test_grouped_alias_tables_encode_the_node_law) and by chi-square on the device draws (tests/test_gpu_parity.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace annembed_host {

inline uint32_t fbits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

// The per-group passes (weights gathered through the numbering, thresholds, the Vose construction inside a line, the copy
// of the alias group's data) touch every node once with a random gather and are independent per group: they run on
// min(hardware threads, 16) host threads, contiguous ranges of groups.  Sums that feed a table entry are taken
// sequentially over the per-group results, so that the tables do not depend on the number of threads.  The Vose loops
// over the groups themselves stay sequential (0.7M lines / 2.75M sectors at 11M nodes).
template <class Body>
void parallel_ranges(uint64_t count, Body body /* (uint64_t begin, uint64_t end) */)
{
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1u : (nt > 16u ? 16u : nt);
    if (count < (1u << 16) || nt == 1) { body((uint64_t)0, count); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    const uint64_t per = (count + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const uint64_t b = std::min<uint64_t>(count, t * per), e = std::min<uint64_t>(count, b + per);
        if (b < e) th.emplace_back([=]() { body(b, e); });
    }
    for (auto &x : th) x.join();
}

// Node-level table (≙ WeightedAliasIndex::new in NodeSampler::new, embedder.rs:916-919): Vose's alias method over the nodes,
// entry of node i = {bits(prob), alias node}.  `tot` = the sum of the weights (sequential, taken by the caller while it
// validates them).  Used for the redraws of every kernel, the wide layouts and ANNEMBED_FLAG_NODE_ALIAS.
template <class Weight>
void build_node_alias_table(uint64_t n, Weight weight, double tot, std::vector<uint2> &tab)
{
    std::vector<double> q(n);
    tab.assign(n, make_uint2(0u, 0u));
    parallel_ranges(n, [&](uint64_t i0, uint64_t i1) { for (uint64_t i = i0; i < i1; i++) q[i] = weight(i) * (double)n / tot; });
    std::vector<uint32_t> small, large;
    small.reserve(n); large.reserve(n);
    for (uint64_t i = 0; i < n; i++) (q[i] < 1.0 ? small : large).push_back((uint32_t)i);
    auto put = [&](uint32_t i, float prob, uint32_t alias) { tab[i] = make_uint2(fbits(prob), alias); };
    while (!small.empty() && !large.empty()) {
        const uint32_t s = small.back(); small.pop_back();
        const uint32_t l = large.back(); large.pop_back();
        put(s, (float)q[s], l);
        q[l] = (q[l] + q[s]) - 1.0;
        (q[l] < 1.0 ? small : large).push_back(l);
    }
    for (uint32_t l : large) put(l, 1.0f, l);
    for (uint32_t s : small) put(s, 1.0f, s);
}

// Sector-level table (embedder.rs:909-931 NodeSampler restated one level up): Vose's alias method over the SECTORS of 4
// consecutive nodes (weight = the sum of the 4 node weights) + the 3 cumulative thresholds that pick a node inside a sector.
// Entry of sector s (32 bytes, so that a draw is ONE gather and branch-free):
// uint4 {bits(prob), alias sector, bits(thr0), bits(thr1)}, uint4 {bits(thr2), bits(thr0), bits(thr1), bits(thr2) of the ALIAS sector}.
template <class Weight>
void build_sector_alias_table(uint64_t n, Weight weight, std::vector<uint4> &tab)
{
    const uint64_t nsec = (n + 3) / 4;
    std::vector<double> q(nsec);
    tab.assign(2 * nsec, make_uint4(0u, 0u, 0u, 0u));
    parallel_ranges(nsec, [&](uint64_t s0, uint64_t s1) {
        for (uint64_t s = s0; s < s1; s++) {
            double w4[4], W = 0.0;
            for (int r = 0; r < 4; r++) { const uint64_t i = 4 * s + r; w4[r] = i < n ? weight(i) : 0.0; W += w4[r]; }
            q[s] = W;
            // cumulative thresholds; a sector of zero weight is never drawn (prob 0 -> its alias), missing rows weigh 0
            float t0 = 1.0f, t1 = 1.0f, t2 = 1.0f;
            if (W > 0.0) { t0 = (float)(w4[0] / W); t1 = (float)((w4[0] + w4[1]) / W); t2 = (float)((w4[0] + w4[1] + w4[2]) / W); }
            tab[2 * s] = make_uint4(0u, (uint32_t)s, fbits(t0), fbits(t1));
            tab[2 * s + 1] = make_uint4(fbits(t2), 0u, 0u, 0u);
        }
    });
    double tot = 0.0;
    for (uint64_t s = 0; s < nsec; s++) tot += q[s];
    std::vector<uint32_t> small, large;
    small.reserve(nsec); large.reserve(nsec);
    for (uint64_t s = 0; s < nsec; s++) { q[s] = q[s] * (double)nsec / tot; (q[s] < 1.0 ? small : large).push_back((uint32_t)s); }
    auto put = [&](uint32_t s, float prob, uint32_t alias) { tab[2 * (size_t)s].x = fbits(prob); tab[2 * (size_t)s].y = alias; };
    while (!small.empty() && !large.empty()) {
        const uint32_t sm = small.back(); small.pop_back();
        const uint32_t lg = large.back(); large.pop_back();
        put(sm, (float)q[sm], lg);
        q[lg] = (q[lg] + q[sm]) - 1.0;
        (q[lg] < 1.0 ? small : large).push_back(lg);
    }
    for (uint32_t lg : large) put(lg, 1.0f, lg);
    for (uint32_t sm : small) put(sm, 1.0f, sm);
    parallel_ranges(nsec, [&](uint64_t s0, uint64_t s1) {   // the alias sector's thresholds ride in the entry: no second table gather
        for (uint64_t s = s0; s < s1; s++) {                 // (reads .z/.w/.x of other entries, writes .y/.z/.w of the second half: disjoint words)
            const uint32_t al = tab[2 * s].y;
            tab[2 * s + 1].y = tab[2 * (size_t)al].z; tab[2 * s + 1].z = tab[2 * (size_t)al].w; tab[2 * s + 1].w = tab[2 * (size_t)al + 1].x;
        }
    });
}

// Line-level tables (event kernels, layouts of dimension <= 4): the G = 16 (dimension 2) or 8 (dimension 3-4) nodes whose
// rows fill a 128-byte line of the layout share the line draw.  Two alias methods, one inside the other: T1 over the LINES
// (weight = sum of the line's node weights): {bits(prob), alias line};  T2 inside every line over its G rows (conditional
// law w_i / W_line): per row (accept threshold in units of 2^-24) << 4 | alias row.  The T2 entry of a line carries the
// inner table of its ALIAS line behind its own ([line][0..G) own, [line][G..2G) alias line's), so that T1 and both candidate
// columns are read in one round.  P(node) = P(line) * P(row | line): exactly the node law.
template <class Weight>
void build_line_alias_tables(uint64_t n, uint32_t G, Weight weight, std::vector<uint2> &t1, std::vector<uint32_t> &t2)
{
    std::vector<uint32_t> small, large;
    const uint64_t nl = (n + G - 1) / G;
    t1.assign(nl, make_uint2(0u, 0u));
    t2.assign(nl * 2 * G, 0u);                     // [line][0..G): own inner table, [line][G..2G): the alias line's
    std::vector<double> ql(nl);
    parallel_ranges(nl, [&](uint64_t l0, uint64_t l1) {
        for (uint64_t l = l0; l < l1; l++) {
            double wr[16], W = 0.0;
            for (uint32_t r = 0; r < G; r++) { const uint64_t i = l * G + r; wr[r] = i < n ? weight(i) : 0.0; W += wr[r]; }
            ql[l] = W;
            // Vose inside the line (rows of weight 0 -- the padding of the last line -- get threshold 0: never accepted)
            uint32_t sm[16], lg[16], nsm = 0, nlg = 0;
            double qi[16];
            for (uint32_t r = 0; r < G; r++) {
                qi[r] = W > 0.0 ? wr[r] * (double)G / W : 1.0;
                if (qi[r] < 1.0) sm[nsm++] = r; else lg[nlg++] = r;
                t2[l * 2 * G + r] = (16777216u << 4) | r;
            }
            while (nsm && nlg) {
                const uint32_t a = sm[--nsm], b = lg[--nlg];
                t2[l * 2 * G + a] = ((uint32_t)std::min(16777216.0, std::floor(qi[a] * 16777216.0 + 0.5)) << 4) | b;
                qi[b] = (qi[b] + qi[a]) - 1.0;
                if (qi[b] < 1.0) sm[nsm++] = b; else lg[nlg++] = b;
            }
        }
    });
    double totl = 0.0;
    for (uint64_t l = 0; l < nl; l++) totl += ql[l];
    for (uint64_t l = 0; l < nl; l++) { ql[l] = ql[l] * (double)nl / totl; (ql[l] < 1.0 ? small : large).push_back((uint32_t)l); t1[l] = make_uint2(fbits(1.0f), (uint32_t)l); }
    while (!small.empty() && !large.empty()) {
        const uint32_t sm = small.back(); small.pop_back();
        const uint32_t lg = large.back(); large.pop_back();
        t1[sm] = make_uint2(fbits((float)ql[sm]), lg);
        ql[lg] = (ql[lg] + ql[sm]) - 1.0;
        (ql[lg] < 1.0 ? small : large).push_back(lg);
    }
    parallel_ranges(nl, [&](uint64_t l0, uint64_t l1) {     // reads the first halves, writes the second halves of the entries
        for (uint64_t l = l0; l < l1; l++) {
            const uint64_t al = t1[l].y;
            for (uint32_t r = 0; r < G; r++) t2[l * 2 * G + G + r] = t2[al * 2 * G + r];
        }
    });
}

} // namespace annembed_host
