// annembed_embedder.hpp -- C++ host-side mirror of the reference's `Embedder` for the hot path, header-only,
// sitting directly on the C ABI of annembed_cuda.h.
//
// The reference is compiled Rust (/root/reference/src/embedder.rs); there is no Rust toolchain in the build image,
// so the host side above the ABI is C++ (and, for the tests and the bench, the Python mirror in annembed_b200/).
// Same names, argument meaning and error behaviour as the Rust struct:
//   Embedder::new(&kgraph, EmbedderParams)        embedder.rs:107   -> Embedder(const KGraph&, EmbedderParams)
//   embed() -> Result<usize, usize>               embedder.rs:183   -> int embed(): 1 on success ("Ok(1)"), throws EmbedError ("Err(1)")
//   get_embedded() / get_embedded_reindexed()     embedder.rs:378,384
//   get_embedded_by_dataid / _by_nodeid           embedder.rs:409,421
//   get_initial_embedding[_reindexed]             embedder.rs:426,430
//   get_nb_nodes, get_asked_dimension, ...        embedder.rs:135-153,785
// dmap_init = true without an explicit initial layout computes the diffusion-map layout on the device
// (annembed_cuda_dmap_init ≙ embedder.rs:308-345).  Not mirrored here (the Python mirror has them): from_hkgraph,
// get_quality_estimate_from_edge_length.
#pragma once
#include <cstdint>
#include <cstdio>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "annembed_cuda.h"

namespace annembed {

// ≙ EmbedderParams, embedparams.rs:76-132 (defaults :107-132)
struct EmbedderParams {
    size_t asked_dim = 2;
    bool dmap_init = true;
    double beta = 1.0;
    double b = 1.0;
    double scale_rho = 1.0;
    double grad_step = 2.0;
    size_t nb_sampling_by_edge = 10;
    size_t nb_grad_batch = 20;
    size_t grad_factor = 4;
    size_t hierarchy_layer = 0;
    bool hubness_weighting = false;
    // device-side additions
    uint32_t mini_epochs_per_batch = 0;
    uint64_t seed = 0x5EED;
    uint32_t flags = 0;

    void set_dmap_init(bool v) { dmap_init = v; }                       // embedparams.rs:151-183
    void set_nb_gradient_batch(size_t nb) { nb_grad_batch = nb; }
    void set_dim(size_t dim) { asked_dim = dim; }
    void set_nb_edge_sampling(size_t nb) { nb_sampling_by_edge = nb; }
    size_t get_dimension() const { return asked_dim; }
    void set_hierarchy_layer(size_t layer) { hierarchy_layer = layer; }
    size_t get_hierarchy_layer() const { return hierarchy_layer; }

    annembed_cuda_params to_c() const
    {
        annembed_cuda_params p;
        annembed_cuda_default_params(&p);
        p.asked_dim = (uint32_t)asked_dim; p.dmap_init = dmap_init; p.beta = beta; p.b = b; p.scale_rho = scale_rho;
        p.grad_step = grad_step; p.nb_sampling_by_edge = (uint32_t)nb_sampling_by_edge;
        p.nb_grad_batch = (uint32_t)nb_grad_batch; p.grad_factor = (uint32_t)grad_factor;
        p.hierarchy_layer = (uint32_t)hierarchy_layer; p.hubness_weighting = hubness_weighting;
        p.mini_epochs_per_batch = mini_epochs_per_batch; p.seed = seed; p.flags = flags;
        return p;
    }
};

// ≙ KGraph<F>, fromhnsw/kgraph.rs:108-120, flattened: neighbours of node i are col/dist[row_ptr[i] .. row_ptr[i+1]),
// sorted by increasing distance (kgraph.rs:508-509); data_id[i] is the caller's DataId of node i (node_set).
struct KGraph {
    std::vector<uint64_t> row_ptr;
    std::vector<uint32_t> col;
    std::vector<float> dist;
    std::vector<uint64_t> data_id;
    size_t max_nbng = 0;

    size_t get_nb_nodes() const { return row_ptr.empty() ? 0 : row_ptr.size() - 1; }   // kgraph.rs:147-164
    size_t get_max_nbng() const { return max_nbng; }
    uint64_t get_data_id_from_idx(size_t idx) const { return data_id.empty() ? idx : data_id[idx]; }   // :335
};

// The ANNKGCSR interchange file (the same bytes as annembed_b200/kgraph.py write_csr / read_csr and rust/kgraph_csr.rs;
// the reference's KGraph has no serialisation).  Little endian (the only byte order this library runs on):
//   magic 8 bytes "ANNKGCSR"; u32 version = 1, u32 flags = 0; u64 n, u64 E, u64 max_nbng;
//   u64 row_ptr[n+1]; u32 col[E]; f32 dist[E]; u64 data_id[n]
inline void write_csr(const std::string &path, const KGraph &g)
{
    const uint64_t n = g.get_nb_nodes(), E = g.col.size();
    if (g.row_ptr.size() != n + 1 || g.dist.size() != E || g.row_ptr.back() != E || (!g.data_id.empty() && g.data_id.size() != n))
        throw std::invalid_argument("write_csr: inconsistent CSR arrays");
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("write_csr: cannot open " + path);
    const uint32_t head[2] = {1u, 0u};
    const uint64_t sizes[3] = {n, E, (uint64_t)g.max_nbng};
    bool ok = std::fwrite("ANNKGCSR", 1, 8, f) == 8 && std::fwrite(head, 4, 2, f) == 2 && std::fwrite(sizes, 8, 3, f) == 3 &&
              std::fwrite(g.row_ptr.data(), 8, n + 1, f) == n + 1 && std::fwrite(g.col.data(), 4, E, f) == E &&
              std::fwrite(g.dist.data(), 4, E, f) == E;
    if (ok) {
        if (g.data_id.empty()) { for (uint64_t i = 0; i < n && ok; i++) ok = std::fwrite(&i, 8, 1, f) == 1; }   // identity DataIds
        else ok = std::fwrite(g.data_id.data(), 8, n, f) == n;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("write_csr: short write to " + path);
}

inline KGraph read_csr(const std::string &path)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("read_csr: cannot open " + path);
    struct Closer { FILE *f; ~Closer() { std::fclose(f); } } closer{f};
    char magic[8];
    uint32_t head[2];
    uint64_t sizes[3];
    if (std::fread(magic, 1, 8, f) != 8 || std::string(magic, 8) != "ANNKGCSR") throw std::runtime_error("not an ANNKGCSR file");
    if (std::fread(head, 4, 2, f) != 2 || head[0] != 1u) throw std::runtime_error("unsupported ANNKGCSR version");
    if (std::fread(sizes, 8, 3, f) != 3) throw std::runtime_error("truncated ANNKGCSR file");
    const uint64_t n = sizes[0], E = sizes[1];
    if (n >= 0xFFFFFFFFull || E >= 0xFFFFFFFFull) throw std::runtime_error("ANNKGCSR: n and E must be below 2^32 - 1");   // the library's limits
    KGraph g;
    g.max_nbng = (size_t)sizes[2];
    g.row_ptr.resize(n + 1); g.col.resize(E); g.dist.resize(E); g.data_id.resize(n);
    if (std::fread(g.row_ptr.data(), 8, n + 1, f) != n + 1 || std::fread(g.col.data(), 4, E, f) != E ||
        std::fread(g.dist.data(), 4, E, f) != E || std::fread(g.data_id.data(), 8, n, f) != n)
        throw std::runtime_error("truncated ANNKGCSR file");
    if (g.row_ptr[0] != 0 || g.row_ptr[n] != E) throw std::runtime_error("ANNKGCSR: row_ptr[n] != E");
    return g;
}

// ≙ Err(1) of Embedder::embed (embedder.rs:183,366-369)
struct EmbedError : std::runtime_error {
    int status;
    EmbedError(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

class Embedder {
public:
    // ≙ Embedder::new (embedder.rs:107): borrows the graph, copies the parameters
    Embedder(const KGraph &kgraph, EmbedderParams parameters, int device = 0)
        : kgraph_(kgraph), parameters_(parameters), device_(device) {}

    // the initial layout (n x asked_dim, row-major, node-index order); optional: without it embed() computes the
    // diffusion-map layout on the device (dmap_init = true, >= 80 nodes) or draws the random one (embedder.rs:348)
    void set_initial_embedding(std::vector<float> y) { initial_embedding_ = std::move(y); }

    size_t get_asked_dimension() const { return parameters_.asked_dim; }          // embedder.rs:135-153
    double get_scale_rho() const { return parameters_.scale_rho; }
    double get_b() const { return parameters_.b; }
    double get_grad_step() const { return parameters_.grad_step; }
    size_t get_nb_grad_batch() const { return parameters_.nb_grad_batch; }
    size_t get_nb_nodes() const { return kgraph_.get_nb_nodes(); }                // :785
    const KGraph &get_kgraph() const { return kgraph_; }
    const std::vector<uint32_t> &get_hubness() const { return hubness_counts_; }  // :156
    double get_initial_cross_entropy() const { return ce_initial_; }              // logged at :846-852
    double get_final_cross_entropy() const { return ce_final_; }                  // :885-886

    // ≙ embed -> one_step_embed (embedder.rs:183,298-371): to_proba_edges + entropy_optimize on the device
    int embed()
    {
        const size_t n = get_nb_nodes(), d = parameters_.asked_dim;
        bool device_dmap = initial_embedding_.empty() && parameters_.dmap_init;         // embedder.rs:308-345 on the device
        if (device_dmap && n < 80) {
            // the device range finder is the reference's rank-20 one (graphlaplace.rs:113) and needs 4 x 20 nodes; the
            // reference switches to a full SVD below 500 nodes (graphlaplace.rs:100-108), which the device path does not
            // have: tiny graphs start from the random layout of the dmap_init = false branch (same rule as embedder.py)
            fprintf(stderr, "annembed: dmap_init needs >= 80 nodes on the device: using the random initial layout (embedder.rs:348)\n");
            device_dmap = false;
        }
        if (initial_embedding_.empty() && !device_dmap) {
            // ≙ get_random_init(1.) (embedder.rs:348,456-470): uniform in [-0.5, 0.5]^d
            std::mt19937_64 rng(parameters_.seed);
            std::uniform_real_distribution<float> u(-0.5f, 0.5f);
            initial_embedding_.resize(n * d);
            for (auto &v : initial_embedding_) v = u(rng);
        }
        if (!device_dmap && initial_embedding_.size() != n * d) throw EmbedError(ANNEMBED_ERR_INVALID_ARG, "initial embedding must be n x asked_dim");
        annembed_cuda_ctx *ctx = nullptr;
        const annembed_cuda_params cp = parameters_.to_c();
        int st = annembed_cuda_create(&ctx, &cp, device_);
        if (st) throw EmbedError(st, annembed_cuda_last_error(nullptr));
        auto check = [&](int s) {
            if (s) {
                std::string msg = annembed_cuda_last_error(ctx);
                annembed_cuda_destroy(ctx);
                throw EmbedError(s, msg);
            }
        };
        if (kgraph_.row_ptr.size() != n + 1 || kgraph_.col.size() != kgraph_.dist.size() || kgraph_.row_ptr.back() != kgraph_.col.size())
            check(ANNEMBED_ERR_INVALID_ARG);                                       // the C side reads row_ptr[n] entries of col / dist
        check(annembed_cuda_set_graph_csr(ctx, n, kgraph_.row_ptr.data(), kgraph_.col.data(), kgraph_.dist.data()));
        check(annembed_cuda_edge_weights(ctx, nullptr, nullptr));                  // embedder.rs:351
        if (parameters_.hubness_weighting) {                                       // embedder.rs:810-837
            hubness_counts_.resize(n);
            check(annembed_cuda_get_hubness_counts(ctx, hubness_counts_.data()));
            std::vector<float> w(n);
            for (size_t i = 0; i < n; i++) w[i] = std::min(std::max((float)hubness_counts_[i], 1.0f), (float)n);
            check(annembed_cuda_set_neg_weights(ctx, w.data()));
        }
        if (device_dmap) {
            initial_embedding_.resize(n * d);
            check(annembed_cuda_dmap_init(ctx, 12, 5.0f, initial_embedding_.data()));   // gnbn, diffusion time: embedder.rs:316-317
        } else {
            check(annembed_cuda_set_embedding(ctx, initial_embedding_.data()));
        }
        check(annembed_cuda_optimize(ctx, &ce_initial_, &ce_final_));              // embedder.rs:356
        embedding_.resize(n * d);
        check(annembed_cuda_get_embedding(ctx, embedding_.data()));
        annembed_cuda_get_stats(ctx, &stats_);
        annembed_cuda_destroy(ctx);
        return 1;
    }

    // node-index order (embedder.rs:378); empty before embed
    const std::vector<float> &get_embedded() const { return embedding_; }
    // ≙ embedder.rs:384-405: row i goes to row DataId(i); throws before embed (the reference panics)
    std::vector<float> get_embedded_reindexed() const
    {
        if (embedding_.empty()) throw std::logic_error("get_embedded_reindexed called before embed");
        return reindex(embedding_);
    }
    std::vector<float> get_initial_embedding_reindexed() const { return reindex(initial_embedding_); }   // :430
    const std::vector<float> &get_initial_embedding() const { return initial_embedding_; }               // :426
    const float *get_embedded_by_nodeid(size_t node) const { return embedding_.data() + node * parameters_.asked_dim; }   // :421
    const annembed_cuda_stats &get_stats() const { return stats_; }

private:
    std::vector<float> reindex(const std::vector<float> &a) const
    {
        const size_t n = get_nb_nodes(), d = parameters_.asked_dim;
        std::vector<float> out(n * d, 0.0f);
        for (size_t i = 0; i < n; i++) {
            const uint64_t id = kgraph_.get_data_id_from_idx(i);
            if (id >= n) throw std::out_of_range("DataIds must fill 0..n to reindex (embedder.rs:397-403)");
            for (size_t c = 0; c < d; c++) out[id * d + c] = a[i * d + c];
        }
        return out;
    }

    const KGraph &kgraph_;
    EmbedderParams parameters_;
    int device_;
    std::vector<float> initial_embedding_, embedding_;
    std::vector<uint32_t> hubness_counts_;
    double ce_initial_ = 0.0, ce_final_ = 0.0;
    annembed_cuda_stats stats_{};
};

} // namespace annembed
