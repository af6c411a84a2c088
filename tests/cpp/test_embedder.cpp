// C++ driver of the drop-in boundary (T9): the reference's only hot-path test restated against the C++ mirror
// (embedder.rs:1435-1467 mini_embed_full: random points, kNN graph with knbn = 10, asked_dim = 5, assert embed() is Ok),
// plus the ABI's error behaviour.  Exit code 0 = pass.  Usage: test_embedder [n]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>

#include "../../include/annembed_embedder.hpp"

using namespace annembed;

static KGraph random_knn_graph(size_t n, size_t dim, size_t k, uint64_t seed)
{
    // 500 random points in dim 20, L1 distance like the reference test (DistL1), exact kNN instead of hnsw_rs
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    std::vector<float> x(n * dim);
    for (auto &v : x) v = u(rng);
    KGraph g;
    g.row_ptr.push_back(0);
    std::vector<std::pair<float, uint32_t>> cand(n);
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) {
            float d = 0;
            for (size_t c = 0; c < dim; c++) d += std::fabs(x[i * dim + c] - x[j * dim + c]);
            cand[j] = {j == i ? 1e30f : d, (uint32_t)j};
        }
        std::partial_sort(cand.begin(), cand.begin() + k, cand.end());
        for (size_t m = 0; m < k; m++) { g.col.push_back(cand[m].second); g.dist.push_back(cand[m].first); }
        g.row_ptr.push_back(g.col.size());
    }
    g.max_nbng = k;
    g.data_id.resize(n);
    std::iota(g.data_id.begin(), g.data_id.end(), 0);
    std::shuffle(g.data_id.begin(), g.data_id.end(), rng);
    return g;
}

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(int argc, char **argv)
{
    const size_t n = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 500;
    KGraph g = random_knn_graph(n, 20, 10, 4664397);
    EmbedderParams params;                       // EmbedderParams::default()
    CHECK(params.asked_dim == 2 && params.nb_grad_batch == 20 && params.nb_sampling_by_edge == 10 && params.grad_step == 2.0);
    params.asked_dim = 5;                        // embedder.rs:1463
    params.dmap_init = false;                    // random initial layout (embedder.rs:348)
    Embedder embedder(g, params);
    int res = 0;
    try { res = embedder.embed(); } catch (const EmbedError &e) { std::fprintf(stderr, "embed failed: %s\n", e.what()); return 1; }
    CHECK(res == 1);                             // assert!(embed_res.is_ok())
    const auto &y = embedder.get_embedded();
    CHECK(y.size() == n * 5);
    for (float v : y) CHECK(std::isfinite(v));
    const auto r = embedder.get_embedded_reindexed();
    for (size_t i = 0; i < n; i++)
        for (size_t c = 0; c < 5; c++) CHECK(r[g.data_id[i] * 5 + c] == y[i * 5 + c]);
    CHECK(std::isfinite(embedder.get_final_cross_entropy()));
    CHECK(embedder.get_stats().positive_samples > 0 && embedder.get_stats().epoch_launches > 0);

    // hubness-weighted negatives (embedder.rs:810-837)
    params.hubness_weighting = true;
    Embedder hub(g, params);
    CHECK(hub.embed() == 1);
    uint64_t tot = 0; for (uint32_t c : hub.get_hubness()) tot += c;
    CHECK(tot == g.col.size());

    // dmap_init = true (the default): the diffusion-map layout is computed on the device (embedder.rs:308-345),
    // boxed to [-5, 5] by set_data_box(., 10)
    EmbedderParams p2;
    Embedder e2(g, p2);
    CHECK(e2.embed() == 1);
    float mx = 0.f;
    for (float v : e2.get_initial_embedding()) mx = std::max(mx, std::fabs(v));
    CHECK(e2.get_initial_embedding().size() == n * 2 && std::fabs(mx - 5.f) < 1e-3f);
    // error behaviour: an empty neighbourhood (kdumap.rs:75-85), wrong layout size
    KGraph bad = g;
    bad.row_ptr[4] = bad.row_ptr[3];             // node 3 loses its neighbours (and row 3/4 become inconsistent -> status)
    p2.dmap_init = false;
    Embedder e3(bad, p2);
    try { e3.embed(); return 1; } catch (const EmbedError &e) { CHECK(e.status == ANNEMBED_ERR_EMPTY_ROW || e.status == ANNEMBED_ERR_INVALID_ARG || e.status == ANNEMBED_ERR_UNSORTED_ROW); }
    Embedder e4(g, p2);
    e4.set_initial_embedding(std::vector<float>(7, 0.f));
    try { e4.embed(); return 1; } catch (const EmbedError &e) { CHECK(e.status == ANNEMBED_ERR_INVALID_ARG); }
    try { e4.get_embedded_reindexed(); return 1; } catch (const std::logic_error &) {}
    std::printf("test_embedder ok: n=%zu ce %.4e -> %.4e, %llu positive samples\n", n, embedder.get_initial_cross_entropy(),
                embedder.get_final_cross_entropy(), (unsigned long long)embedder.get_stats().positive_samples);
    return 0;
}
