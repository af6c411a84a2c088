// annembed_cuda.cu -- kernels and C ABI of the B200-native cross-entropy embedding optimizer.
// See include/annembed_cuda.h for the boundary and DESIGN.md for the data layout and rooflines.
// Reference citations are relative to /root/reference/src.
#include "../../include/annembed_cuda.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen'ed in comm_init, never linked

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "sgd_core.cuh"
#include "alias_tables.hpp"
#include "quality.cuh"
#include "dmap.cuh"

using namespace annembed;

// =====================================================================================================
// small host utilities
// =====================================================================================================
static thread_local std::string g_create_error;

// Device buffer.  Allocations come from the device's stream-ordered memory pool (cudaMallocAsync) whose release
// threshold is raised at context creation: destroying a context hands its gigabytes back to the pool instead of
// unmapping them (cudaFree of the 11M-node working set took 30-900 ms), and the next context reuses them.
// `plain` buffers use cudaMalloc: the layout replicas are exported through CUDA IPC (fused multi-GPU exchange).
// Every release happens after the owning context's stream has been synchronised, so freeing on the legacy stream is safe.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool plain = false;
    cudaError_t alloc(size_t count, bool use_plain = false)
    {
        release();
        if (count == 0) { return cudaSuccess; }
        plain = use_plain;
        cudaError_t e;
        if (plain) {
            e = cudaMalloc((void **)&p, count * sizeof(T));
        } else {
            e = cudaMallocAsync((void **)&p, count * sizeof(T), (cudaStream_t)0);
            if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);
        }
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    void release()
    {
        if (p) { if (plain) cudaFree(p); else cudaFreeAsync(p, (cudaStream_t)0); }
        p = nullptr; n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool load(std::string &err)
    {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define LOADSYM(name) name = (decltype(name))dlsym(handle, "nccl" #name); if (!name) { err = "missing symbol nccl" #name; return false; }
        LOADSYM(GetUniqueId) LOADSYM(CommInitRank) LOADSYM(CommDestroy) LOADSYM(AllGather) LOADSYM(AllReduce) LOADSYM(Broadcast)
        LOADSYM(GroupStart) LOADSYM(GroupEnd) LOADSYM(GetErrorString)
#undef LOADSYM
        return true;
    }
};
static NcclApi g_nccl;

struct annembed_cuda_ctx {
    annembed_cuda_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // second stream: staggered sub-ranges when the exchange is fused
    cudaStream_t launch_stream = nullptr;        // stream the epoch kernels are launched on (stream or stream2)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int n_sm = 148;
    int last_epoch_kernels = 1;                  // kernels per mini-epoch of the path in use (tiled: 2, generic: 1)
    int l2_persist_max = 0, l2_window_max = 0;   // bytes (device attributes)
    int sm_count = 148;
    std::vector<double> dmap_sigma;              // singular values of the last dmap_init (diagnostics)
    std::vector<float> dmap_omega;               // caller-provided test matrix of the range finder (tests), else Philox

    // graph (replicated on every rank)
    uint64_t n = 0, E = 0;
    uint32_t kmax = 0, kmin = 0;
    DevBuf<uint64_t> row_ptr;
    DevBuf<uint32_t> col;
    DevBuf<float> dist, rho;
    DevBuf<float> scale, proba;
    bool have_graph = false, have_weights = false, have_build = false, have_embedding = false, have_alias = false;

    // internal node numbering of the optimizer (locality relabelling, built once per graph by ensure_struct)
    DevBuf<uint32_t> new_of_old, old_of_new;
    DevBuf<uint64_t> row_ptr2;      // CSR of the relabelled graph
    DevBuf<uint32_t> col2;
    DevBuf<float> cum;              // inclusive cumulative row probability, relabelled edge order
    DevBuf<float> inv_s2n;          // 1 / embedded_scale^2, relabelled node order
    DevBuf<uint2> rowpack;          // [n][KP] {col, bits(cum)} rows of the tiled kernels
    DevBuf<uint32_t> erank;         // [n][KP] position of every out-edge in the transposed index (single rank)
    DevBuf<unsigned char> fired;    // [E] firing counts pushed by k_epoch_out (single rank)
    int KP = 0;                     // padded row length of the tiled kernels (0: generic kernel only)
    bool alias_dirty = false;
    // cells of the internal numbering (cell_epoch.cuh): consecutive node ranges of at most cell_nodes nodes, cut at the
    // boundaries of graph-local segments where possible; ranks own whole cells
    uint32_t cell_nodes = 4096;
    std::vector<uint32_t> cell_start_h;          // n_cells + 1
    DevBuf<uint32_t> cell_start;
    uint32_t n_cells = 0, cell_lo = 0, cell_hi = 0;
    std::vector<uint32_t> shard_lo;              // first node of every rank's shard (nranks + 1 entries)
    DevBuf<uint64_t> ext_ptr;                    // per cell: its in-edges whose source lies in another cell
    DevBuf<uint32_t> ext_slot;
    uint64_t ext_cnt = 0;                        // owned in-edges whose source lies in another cell
    uint64_t ext_cnt_global = 0;                 // the same count over the whole graph (identical on every rank)
    uint64_t cross_rank_edges = 0;               // edges whose ends are owned by different ranks (whole graph)
    uint32_t cell_map_bytes = 0;                 // largest number of in-edges of a cell (whole graph), rounded up to 16
    int smem_optin_max = 0;
    uint32_t last_substeps = 0;

    // optimizer context (≙ EntropyOptim, embedder.rs:936-951)
    DevBuf<float> emb_scale, inv_s2;
    DevBuf<uint64_t> in_ptr_all;   // n+1, transposed index of the whole graph
    DevBuf<uint4> in_rec;          // in-edge records of the owned slice
    DevBuf<uint32_t> in_src, in_eid; // structure of the transposed index (graph only)
    uint64_t in_cnt = 0;
    bool have_struct = false;
    bool struct_async = false;     // the structures were built for the asynchronous form (no transposed index)
    uint64_t in_base = 0;
    DevBuf<uint2> neg_alias_old, neg_alias;   // alias table in the caller's / in the internal numbering
    DevBuf<uint4> sec_alias;                  // sector-level alias table, internal numbering (2 x uint4 per sector of 4 nodes)
    DevBuf<uint2> line_t1;                    // line-level alias table (event kernels, dimension <= 4): {bits(prob), alias line} per line
    DevBuf<uint32_t> line_t2;                 // ... and the alias table INSIDE a line: (accept threshold of 2^24 << 4) | alias row, per row; per line its own and its alias line's
    std::vector<float> neg_w_host;            // the caller's sampling weights (kept to build the sector table once the numbering is known)
    DevBuf<float> yapi, y0;        // current and initial layout in the caller's node order
    DevBuf<float> y[2];            // double-buffered layout of the epoch loop, internal node order (exported to the peers)
    int cur = 0;
    int DP = 2;

    // shard
    int rank = 0, nranks = 1;
    uint32_t lo = 0, hi = 0, n_pad = 0;
    ncclComm_t comm = nullptr;
    float *peer_y[8][2] = {};      // fused exchange: the two layout buffers of every rank, opened through CUDA IPC
    bool have_peers = false;
    DevBuf<float> barrier_buf;

    // scratch
    DevBuf<double> partials;
    DevBuf<unsigned long long> counter;
    DevBuf<unsigned long long> errword;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;

    annembed_cuda_stats st{};
};

static void close_peers(annembed_cuda_ctx *ctx);
static int alloc_layout(annembed_cuda_ctx *ctx);

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return ANNEMBED_ERR_CUDA;                                                                \
        }                                                                                            \
    } while (0)

#define REQUIRE(cond, code, msg)                                                                     \
    do { if (!(cond)) { ctx->err = (msg); return (code); } } while (0)

static inline unsigned int nblocks(uint64_t n, unsigned int bs) { return (unsigned int)((n + bs - 1) / bs); }
static int pad_dim(uint32_t d) { return d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32; }

// =====================================================================================================
// kernels: graph validation, K0, K1, K1b, perplexity
// =====================================================================================================
enum { GERR_EMPTY = 1, GERR_COL = 2, GERR_SELF = 3, GERR_UNSORTED = 4, GERR_ROWPTR = 5 };

__global__ void k_validate_graph(uint64_t n, uint64_t E, const uint64_t *__restrict__ row_ptr,
                                 const uint32_t *__restrict__ col, const float *__restrict__ dist,
                                 unsigned long long *err, unsigned int *kmax, unsigned int *kmin)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    unsigned int code = 0;
    if (r1 < r0 || r1 > E) code = GERR_ROWPTR;
    else if (r1 == r0) code = GERR_EMPTY;
    else {
        float prev = dist[r0];
        if (!(prev >= 0.0f)) code = GERR_UNSORTED;          // NaN or negative distance
        for (uint64_t m = r0; m < r1 && !code; m++) {
            const uint32_t c = col[m];
            const float d = dist[m];
            if (c >= n) code = GERR_COL;
            else if (c == i) code = GERR_SELF;
            else if (!(d >= prev)) code = GERR_UNSORTED;
            prev = d;
        }
        atomicMax(kmax, (unsigned int)(r1 - r0));
        atomicMin(kmin, (unsigned int)(r1 - r0));
    }
    if (code) atomicMin(err, ((unsigned long long)i << 4) | code);
}

// K0: first-neighbour distance of every node, compacted (kdumap.rs:149-152 reads it k+1 times per node)
__global__ void k_first_dist(uint64_t n, const uint64_t *__restrict__ row_ptr, const float *__restrict__ dist,
                             float *__restrict__ rho)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rho[i] = dist[row_ptr[i]];
}

// K1: kdumap.rs:132-235.  One thread per node, sums in the reference's left-to-right order.
// exp/pow are evaluated in fp64 and rounded once, i.e. correctly rounded fp32 functions.
__global__ void __launch_bounds__(256)
k_edge_weights(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
               const float *__restrict__ dist, const float *__restrict__ rho, float scale_rho, float beta,
               float *__restrict__ scale_out, float *__restrict__ p_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    const uint64_t k = r1 - r0;
    const float rho_x = rho[i];
    float sum = 0.0f;
    for (uint64_t m = r0; m < r1; m++) sum = __fadd_rn(sum, rho[col[m]]);              // :149-152
    sum = __fadd_rn(sum, rho_x);                                                        // :154
    const float mean_rho = __fdiv_rn(sum, (float)(k + 1));                              // :155
    const float scale = __fmul_rn(scale_rho, mean_rho);                                 // :159
    scale_out[i] = scale;
    bool all_equal = true;
    float last_dist = 0.0f;
    for (uint64_t m = r1; m > r0; m--) {                                                // :163-170
        const float d = dist[m - 1];
        if (d > 0.0f) { last_dist = d; all_equal = false; break; }
    }
    if (!all_equal && last_dist > rho_x) {
        float wsum = 0.0f;
        for (uint64_t m = r0; m < r1; m++) {
            const float a = __fdiv_rn(fmaxf(__fsub_rn(dist[m], rho_x), 0.0f), scale);   // :172-174
            const float pw = (beta == 1.0f) ? a : (float)pow((double)a, (double)beta);
            float w = (float)exp(-(double)pw);
            w = fmaxf(w, 1.0e-4f);                                                      // PROBA_MIN, NaN -> floor
            p_out[m] = w;
            wsum = __fadd_rn(wsum, w);                                                  // :215
        }
        for (uint64_t m = r0; m < r1; m++) p_out[m] = __fdiv_rn(p_out[m], wsum);        // :216-218
    } else {
        const float u = __fdiv_rn(1.0f, (float)k);                                      // :224-230
        for (uint64_t m = r0; m < r1; m++) p_out[m] = u;
    }
}

// nodeparam.rs:88-91
__global__ void k_perplexity(uint64_t n, const uint64_t *__restrict__ row_ptr, const float *__restrict__ p,
                             float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h = 0.0f;
    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) h += -p[m] * logf(p[m]);
    out[i] = expf(h);
}

// K1b: embedder.rs:760-783 + tools/dichotomy.rs:4-65 (dead code in the reference; kept as an option)
__device__ __forceinline__ float umap_f(const float *__restrict__ d, uint64_t r0, uint64_t r1, float rho, float beta)
{
    float s = 0.0f;
    for (uint64_t m = r0; m < r1; m++) s = __fadd_rn(s, expf(-(d[m] - rho) * beta));
    return s;
}
__global__ void k_edge_weights_umap(uint64_t n, const uint64_t *__restrict__ row_ptr, const float *__restrict__ dist,
                                    float target, float *__restrict__ scale_out, float *__restrict__ w_out,
                                    uint8_t *__restrict__ status)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r0 = row_ptr[i], r1 = row_ptr[i + 1];
    const float rho = dist[r0];
    const float FMAX = 3.402823466e+38f;
    float lower = 0.0f, upper = FMAX, middle = 1.0f;
    const float fl = umap_f(dist, r0, r1, rho, lower), fu = umap_f(dist, r0, r1, rho, upper);
    uint8_t st = 0;
    if (fmaxf(fl, fu) < target || fminf(fu, fl) > target || fu > fl) st = 2;            // dichotomy.rs:21-33 panics
    if (!st) {
        int it = 0;
        while (fabsf(target - umap_f(dist, r0, r1, rho, middle)) > 1.0e-5f) {          // :41
            if (umap_f(dist, r0, r1, rho, middle) > target) lower = middle; else upper = middle;   // decreasing f
            middle = (lower + upper) * 0.5f;
            if (++it > 100) { st = 1; break; }                                          // :59-61
        }
    }
    scale_out[i] = 1.0f / middle;
    for (uint64_t m = r0; m < r1; m++) w_out[m] = expf(-(dist[m] - rho) * middle);
    if (status) status[i] = st;
}

// =====================================================================================================
// kernels: K2 and the transposed index
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) k_partial_sum_f64(uint64_t n, const T *__restrict__ x, double *__restrict__ partials)
{
    double s = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        s += (double)x[i];
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double tot = BR(tmp).Sum(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}
// fixed-order final sum -> deterministic
__global__ void __launch_bounds__(256) k_final_sum_f64(unsigned int m, const double *__restrict__ partials, double *__restrict__ out)
{
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < m; i += 256) s += partials[i];
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double tot = BR(tmp).Sum(s);
    if (threadIdx.x == 0) out[0] = tot;
}

// K2: embedder.rs:1356-1373.  mean from a deterministic fp64 reduction (the reference's sequential fp32 sum
// drifts by ~1e-4 relative at 1e7 nodes)
__global__ void k_embedded_scales(uint64_t n, const float *__restrict__ scale, const double *__restrict__ sum,
                                  float *__restrict__ emb_scale, float *__restrict__ inv_s2)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mean = (float)(sum[0] / (double)n);
    const float s = 0.2f * fmaxf(fminf(__fdiv_rn(scale[i], mean), 4.0f), 0.25f);
    emb_scale[i] = s;
    inv_s2[i] = __fdiv_rn(1.0f, __fmul_rn(s, s));
}

__global__ void k_iota(uint64_t n, uint32_t *x)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (uint32_t)i;
}

// in_ptr_all[v] = first position q with sorted_dst[q] >= v
__global__ void k_in_ptr(uint64_t E, uint64_t n, const uint32_t *__restrict__ sorted_dst, uint64_t *__restrict__ in_ptr)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > E) return;
    const uint64_t prev = (q == 0) ? 0 : (uint64_t)sorted_dst[q - 1] + 1;
    const uint64_t curr = (q == E) ? n + 1 : (uint64_t)sorted_dst[q] + 1;   // exclusive upper of v to fill
    for (uint64_t v = prev; v < curr; v++) in_ptr[v] = q;
}

// structure of the transposed index (depends on the graph only): source node, edge id and owner lane of every
// owned in-edge, in (destination, edge id) order
__global__ void k_in_struct(uint64_t q_lo, uint64_t q_hi, uint64_t n, const uint32_t *__restrict__ sorted_eid,
                            const uint32_t *__restrict__ sorted_dst, uint32_t lo, const uint64_t *__restrict__ row_ptr,
                            uint32_t *__restrict__ in_src, uint32_t *__restrict__ in_eid)
{
    const uint64_t q = q_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= q_hi) return;
    const uint32_t e = sorted_eid[q];
    uint64_t a = 0, b = n;                       // largest i with row_ptr[i] <= e
    while (b - a > 1) { const uint64_t mid = (a + b) >> 1; if (row_ptr[mid] <= e) a = mid; else b = mid; }
    in_src[q - q_lo] = (uint32_t)a;
    in_eid[q - q_lo] = e;
}
// payload (depends on the weights): {src, P_lo, P_hi, 1/s_src^2}
__global__ void k_in_rec(uint64_t cnt, const uint32_t *__restrict__ in_src, const uint32_t *__restrict__ in_eid,
                         const uint64_t *__restrict__ row_ptr, const float *__restrict__ cum,
                         const float *__restrict__ inv_s2, uint4 *__restrict__ rec)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cnt) return;
    const uint32_t src = in_src[q], e = in_eid[q];
    const float P_lo = (row_ptr[src] == e) ? 0.0f : cum[e - 1];
    rec[q] = make_uint4(src, __float_as_uint(P_lo), __float_as_uint(cum[e]), __float_as_uint(inv_s2[src]));
}

// ---- locality relabelling of the nodes (internal numbering of the optimizer context) -------------------------------
// Min-label propagation over the symmetrised graph: after r rounds L_r(i) is the smallest random label within r hops of
// i, so equal labels mark graph-local cells whose size grows with r.  Sorting the nodes by (L_8, L_4, L_2, L_1) nests
// the cells: neighbours in the graph get nearby ids, which turns the y_j / source-row gathers, the in-edge
// records and the firing-count pushes of a tile into accesses to a few nearby cache lines.  Deterministic (min is
// order-independent); depends on the graph only.
__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x21F0AAADu; x ^= x >> 15; x *= 0x735A2D97u; x ^= x >> 15;
    return x;
}
__global__ void k_lp_init(uint64_t n, uint32_t *__restrict__ L)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L[i] = mix32((uint32_t)i * 0x9E3779B1u + 0x7F4A7C15u);
}
// Ln (initialised to L) <- min over the closed neighbourhood, both edge directions
__global__ void k_lp_round(uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                           const uint32_t *__restrict__ L, uint32_t *__restrict__ Ln)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t mine = L[i];
    uint32_t m = mine;
    for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
        const uint32_t j = col[e], lj = L[j];
        m = min(m, lj);
        if (mine < lj) atomicMin(Ln + j, mine);
    }
    if (m < mine) atomicMin(Ln + i, m);
}
__global__ void k_gather_u32(uint64_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ src, uint32_t *__restrict__ dst)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_gather_f32(uint64_t n, const uint32_t *__restrict__ idx, const float *__restrict__ src, float *__restrict__ dst)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_invert_perm(uint64_t n, const uint32_t *__restrict__ order, uint32_t *__restrict__ inv)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[order[i]] = (uint32_t)i;
}
// degrees of the relabelled rows (entry n = 0 so that an exclusive scan over n+1 entries yields row_ptr2)
__global__ void k_relabel_degrees(uint64_t n, const uint32_t *__restrict__ old_of_new, const uint64_t *__restrict__ row_ptr,
                                  uint64_t *__restrict__ deg2)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { deg2[n] = 0; return; }
    const uint32_t o = old_of_new[i];
    deg2[i] = row_ptr[o + 1] - row_ptr[o];
}
__global__ void k_relabel_rows(uint64_t n, const uint32_t *__restrict__ old_of_new, const uint32_t *__restrict__ new_of_old,
                               const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                               const uint64_t *__restrict__ row_ptr2, uint32_t *__restrict__ col2)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = old_of_new[i];
    const uint64_t r0 = row_ptr[o], k = row_ptr[o + 1] - r0, w0 = row_ptr2[i];
    for (uint64_t m = 0; m < k; m++) col2[w0 + m] = new_of_old[col[r0 + m]];
}
// rows of the tiled kernels: KP entries {neighbour, bits(cumulative probability)} per node, pads {NO_NODE, 1.0f}
// Rows padded to KP {neighbour, cumulative probability} pairs.  interleave = 0: [node][KP] (bulk-synchronous kernels; the
// cell kernel copies whole tiles).  interleave = 1 (asynchronous form, async_sweep.cuh async_row_ptr): per tile of 32 nodes
// [pair h = m / 2][lane][m & 1], so that the h-th 16-byte load of a warp reads 512 contiguous bytes (no effect at KP = 6,
// where the L1 merged the three loads of a lane -- profiles/r02_ab_build_flags.txt; kept for the wide rows).
__global__ void k_rowpack(uint64_t n, int KP, int interleave, const uint64_t *__restrict__ row_ptr2, const uint32_t *__restrict__ col2,
                          const float *__restrict__ cum, uint2 *__restrict__ rowpack)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (uint64_t)KP) return;
    const uint64_t i = t / KP; const uint64_t m = t % KP;
    const uint64_t r0 = row_ptr2[i], k = row_ptr2[i + 1] - r0;
    const uint64_t w = interleave ? ((((i >> 5) * (uint64_t)(KP / 2) + (m >> 1)) * 32 + (i & 31)) * 2 + (m & 1)) : t;
    rowpack[w] = m < k ? make_uint2(col2[r0 + m], __float_as_uint(cum[r0 + m])) : make_uint2(ANNEMBED_NO_NODE, __float_as_uint(1.0f));
}
// erank[src * KP + m] = position q of out-edge m of src in the transposed index
__global__ void k_erank(uint64_t cnt, uint64_t q_lo, int KP, const uint32_t *__restrict__ in_src, const uint32_t *__restrict__ in_eid,
                        const uint64_t *__restrict__ row_ptr2, uint32_t *__restrict__ erank)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cnt) return;
    const uint32_t src = in_src[q];
    erank[(uint64_t)src * KP + (in_eid[q] - row_ptr2[src])] = (uint32_t)(q_lo + q);
}
// alias table of the hubness sampler carried into the internal numbering (slots and alias targets)
__global__ void k_relabel_alias(uint64_t n, const uint32_t *__restrict__ new_of_old, const uint2 *__restrict__ tab_old,
                                uint2 *__restrict__ tab_new)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 t = tab_old[i];
    tab_new[new_of_old[i]] = make_uint2(t.x, new_of_old[t.y]);
}
// layout rows between the caller's node order and the internal one
__global__ void k_rows_to_internal(uint64_t n, int DP, const uint32_t *__restrict__ old_of_new, const float *__restrict__ in,
                                   float *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (uint64_t)DP) return;
    const uint64_t i = t / DP; const int c = (int)(t % DP);
    out[t] = in[(uint64_t)old_of_new[i] * DP + c];
}
__global__ void k_rows_from_internal(uint64_t n, int DP, const uint32_t *__restrict__ old_of_new, const float *__restrict__ in,
                                     float *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (uint64_t)DP) return;
    const uint64_t i = t / DP; const int c = (int)(t % DP);
    out[(uint64_t)old_of_new[i] * DP + c] = in[t];
}

// ---- cells (cell_epoch.cuh) --------------------------------------------------------------------------------------
// first position of every run of equal keys (segments of the top-level locality label in the sorted order)
__global__ void k_seg_flags(uint64_t n, const uint32_t *__restrict__ key, unsigned char *__restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}
__device__ __forceinline__ uint32_t cell_of_node(uint32_t n_cells, const uint32_t *__restrict__ cell_start, uint32_t node)
{
    uint32_t a = 0, b = n_cells;                 // largest c with cell_start[c] <= node
    while (b - a > 1) { const uint32_t mid = (a + b) >> 1; if (cell_start[mid] <= node) a = mid; else b = mid; }
    return a;
}
// sort key of position `pos` of the locality order: (its cell, a hash of the node) -- the order INSIDE a cell is random.
// The cell kernel reads intra-cell partners from shared memory, so it needs no finer locality; and the 4 nodes of an
// aligned group, which share their negative streams (sgd_core.cuh neg_stream_key), must not be graph neighbours: groups
// of adjacent nodes repelled by the same far nodes at the same time move coherently, which measurably (3 % on the
// neighbourhood-conservation statistics of the Higgs-shape case) tightens neighbourhoods relative to the reference's
// independent draws.
__global__ void k_cell_sort_keys(uint64_t n, uint32_t n_cells, const uint32_t *__restrict__ cell_start,
                                 const uint32_t *__restrict__ order, unsigned long long *__restrict__ key)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of_node(n_cells, cell_start, (uint32_t)i);
    key[i] = ((unsigned long long)c << 32) | mix32(order[i] * 0x9E3779B1u + 0x3C6EF372u);
}
// in-edge q (destination dst, source src) is external when src lies outside dst's cell
__global__ void k_ext_flags(uint64_t cnt, uint64_t q_lo, const uint32_t *__restrict__ sorted_dst, const uint32_t *__restrict__ in_src,
                            uint32_t n_cells, const uint32_t *__restrict__ cell_start, unsigned char *__restrict__ flag)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cnt) return;
    const uint32_t c = cell_of_node(n_cells, cell_start, sorted_dst[q_lo + q]);
    const uint32_t src = in_src[q];
    flag[q] = (src < cell_start[c] || src >= cell_start[c + 1]) ? 1 : 0;
}
// ext_ptr[c] = first entry of ext_slot (ascending slots, relative to in_base) at or after the cell's first in-edge
__global__ void k_ext_ptr(uint32_t n_cells, const uint32_t *__restrict__ cell_start, const uint64_t *__restrict__ in_ptr_all,
                          uint64_t in_base, uint64_t cnt, uint64_t ext_cnt, const uint32_t *__restrict__ ext_slot,
                          uint64_t *__restrict__ ext_ptr)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    const uint64_t qa = in_ptr_all[cell_start[c]];
    const uint64_t rel = qa <= in_base ? 0 : (qa - in_base >= cnt ? cnt : qa - in_base);
    uint64_t a = 0, b = ext_cnt;                 // first x with ext_slot[x] >= rel
    while (a < b) { const uint64_t mid = (a + b) >> 1; if ((uint64_t)ext_slot[mid] < rel) a = mid + 1; else b = mid; }
    ext_ptr[c] = a;
}
// edges of the whole (relabelled) graph whose two ends lie in different cells
__global__ void __launch_bounds__(256)
k_count_cross_cell(uint64_t n, const uint64_t *__restrict__ row_ptr2, const uint32_t *__restrict__ col2, uint32_t n_cells,
                   const uint32_t *__restrict__ cell_start, unsigned long long *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int c = 0;
    if (i < n) {
        const uint32_t cc = cell_of_node(n_cells, cell_start, (uint32_t)i);
        const uint32_t a = cell_start[cc], b = cell_start[cc + 1];
        for (uint64_t e = row_ptr2[i]; e < row_ptr2[i + 1]; e++) { const uint32_t j = col2[e]; c += (j < a || j >= b) ? 1u : 0u; }
    }
    typedef cub::BlockReduce<unsigned int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const unsigned int tot = BR(tmp).Sum(c);
    if (threadIdx.x == 0 && tot) atomicAdd(out, (unsigned long long)tot);
}
__global__ void k_cell_in_max(uint32_t cell_lo, uint32_t cell_hi, const uint32_t *__restrict__ cell_start,
                              const uint64_t *__restrict__ in_ptr_all, unsigned int *__restrict__ out_max)
{
    const uint32_t c = cell_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cell_hi) return;
    atomicMax(out_max, (unsigned int)(in_ptr_all[cell_start[c + 1]] - in_ptr_all[cell_start[c]]));
}

__global__ void k_in_degree(uint64_t E, const uint32_t *__restrict__ col, uint32_t *__restrict__ cnt)
{
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < E) atomicAdd(cnt + col[m], 1u);
}
__global__ void k_degree_u32(uint64_t n, const uint64_t *__restrict__ ptr, uint32_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(ptr[i + 1] - ptr[i]);
}

// N1: second-step initial layout of the hierarchical embedding, embedder.rs:245-269.  One thread per node.
// Gaussian noise: Box-Muller on Philox4x32-10 words keyed (seed; node, coordinate block, 0, tag 0x51).
__global__ void k_project_init(uint64_t n, uint64_t n_small, int d, int DP, const float *__restrict__ first,
                               const uint32_t *__restrict__ proj_node, const float *__restrict__ proj_dist,
                               float median_dist, uint32_t k0, uint32_t k1, float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *o = out + i * DP;
    if (i < n_small) {                                                            // :249-253
        for (int j = 0; j < DP; j++) o[j] = j < d ? first[i * d + j] : 0.0f;
        return;
    }
    const float ratio = proj_dist[i] / median_dist;                               // :262
    const float correction = sqrtf(ratio / (float)d);                             // :263
    const float *src = first + (uint64_t)proj_node[i] * d;
    for (int j0 = 0; j0 < DP; j0 += 4) {
        const Philox4 w = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)(j0 >> 2), 0x51u, k0, k1);
        // two Box-Muller pairs -> four standard normals
        const float u1 = ((float)(w.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(w.y >> 8) * (1.0f / 16777216.0f);
        const float u3 = ((float)(w.z >> 8) + 1.0f) * (1.0f / 16777216.0f), u4 = (float)(w.w >> 8) * (1.0f / 16777216.0f);
        const float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u3));
        float z[4];
        sincospif(2.0f * u2, &z[1], &z[0]); z[0] *= r1; z[1] *= r1;
        sincospif(2.0f * u4, &z[3], &z[2]); z[2] *= r2; z[3] *= r2;
        for (int t = 0; t < 4 && j0 + t < DP; t++) {
            const int j = j0 + t;
            const float c = fminf(fmaxf(correction * z[t], -2.0f), 2.0f);         // clip(.., 2.) tools/clip.rs:5-18
            o[j] = j < d ? src[j] + c : 0.0f;                                     // :265-267
        }
    }
}

__global__ void k_pad_rows(uint64_t n, int d, int DP, const float *__restrict__ in, float *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (uint64_t)DP) return;
    const uint64_t i = t / DP; const int c = (int)(t % DP);
    out[t] = c < d ? in[i * d + c] : 0.0f;
}
__global__ void k_unpad_rows(uint64_t n, int d, int DP, const float *__restrict__ in, float *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (uint64_t)d) return;
    const uint64_t i = t / d; const int c = (int)(t % d);
    out[t] = in[i * DP + c];
}

// =====================================================================================================
// kernels: K3, K4, K5, draws
// =====================================================================================================
// K3: strict list order on one thread (test mode: correctness, not speed)
template <int DP>
__global__ void k_step_fixed(float *Y, uint64_t n, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                             const float *__restrict__ p, const float *__restrict__ inv_s2, SgdConst K, uint64_t n_samples,
                             const uint64_t *__restrict__ edge_idx, const uint32_t *__restrict__ negs)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint64_t s = 0; s < n_samples; s++) {
        const uint64_t e = edge_idx[s];
        uint64_t a = 0, b = n;
        while (b - a > 1) { const uint64_t mid = (a + b) >> 1; if (row_ptr[mid] <= e) a = mid; else b = mid; }
        fixed_sample<DP>(Y, (uint32_t)a, col[e], p[e], inv_s2[a], K, negs + 5 * s);
    }
}

// K4 (generic fallback): one thread per owned node, rows read from global memory.  Used when a row can be longer
// than 16 neighbours; also the on-device cross-check of the tiled kernel below (tests/test_gpu_parity.py).
template <int DP, bool HUB>
__global__ void __launch_bounds__(256) k_epoch_generic(EpochArgs a, unsigned long long *sample_counter)
{
    const uint32_t node = a.lo + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int applied = 0;
    if (node < a.hi) applied = epoch_node_v2<DP, HUB>(a, node);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if ((threadIdx.x & 31) == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}

// K4 (tiled) = two kernels per mini-epoch; one warp owns a tile of 32 consecutive nodes in both, warps are independent
// (no block barrier), nothing is atomic on the layout.
//  k_epoch_out  lane = node.  The node's row (KP padded {neighbour, cumulative probability} pairs, 16-byte vector loads,
//               no shared memory) lives in registers: rejection test, edge pick (select chain) and probabilities are
//               register-only.  Every node fires ceil(kappa - u) times (systematic sampling) so lanes stay converged;
//               the 6 gathers of firings s+1 and s+2 are in flight during the arithmetic of firing s.  Writes the
//               node's position after its own firings to y_next and (single rank) pushes the firing count of every
//               fired edge into the byte map `fired` at the edge's position in the transposed index.
//  k_epoch_in_flags (single rank)  lane = in-edge slot.  Reads the tile's slice of the byte map (coalesced), clears
//               it, and only for the fired slots (15-30 % at the fine levels) loads the in-edge record and gathers the
//               source row; compacted into a ring and applied by the dense stage below.
//  k_epoch_in   (multi rank: sources live on other ranks)  lane = in-edge.  The warp sweeps the tile's in-edge records
//               (coalesced 16-byte streaming loads) and REPLAYS the source's firing decision from (src, epoch, P_lo,
//               P_hi) -- no communication -- then proceeds like the flags kernel.
//  Dense stage: every fired in-edge is the affine map y -> (1+A) y - A y_src with A evaluated at the owner's position
//               after k_epoch_out; the maps of each owner are composed with a segmented warp scan and the owner lane
//               applies the composite: y_next <- alpha * y_next + beta.
template <int DP, int KP>
struct EpochTile {
#ifndef ANNEMBED_WARPS_OUT
#define ANNEMBED_WARPS_OUT 4
#endif
#ifndef ANNEMBED_MINB_OUT
#define ANNEMBED_MINB_OUT 6
#endif
    static constexpr int WARPS = DP <= 4 ? ANNEMBED_WARPS_OUT : (DP <= 16 ? 4 : 2);
    static constexpr int MINB = DP <= 2 ? (KP <= 8 ? ANNEMBED_MINB_OUT : 4) : (DP <= 4 ? 4 : (DP <= 8 ? 2 : 1));  // blocks/SM the register budget aims at
    static constexpr int MAX_FIRINGS = 126;                                      // per node and mini-epoch (byte counters)
};

// Number of edges of the row whose cumulative firing count is <= s, i.e. the edge firing s lands on.  The counts
// ceil(kappa P_m - u) of the row are kept as bytes (<= 126; 0x7f pads the last word) in registers: byte-wise
// (0x80 | s) - ch never borrows and leaves bit 7 set exactly when ch <= s.
template <int KP>
__device__ __forceinline__ int edge_of_firing(const uint32_t (&chb)[(KP + 3) / 4], int s)
{
    const uint32_t S = (uint32_t)s * 0x01010101u | 0x80808080u;
    int m = 0;
#pragma unroll
    for (int w = 0; w < (KP + 3) / 4; w++) m += __popc((S - chb[w]) & 0x80808080u);
    return m;
}

#include "cell_epoch.cuh"     // the cell-resident form of K4 (uses edge_of_firing)

template <int DP, bool HUB, int KP>
__global__ void __launch_bounds__(EpochTile<DP, KP>::WARPS * 32, EpochTile<DP, KP>::MINB)
k_epoch_out(EpochArgs a, unsigned long long *sample_counter)
{
    using TL = EpochTile<DP, KP>;
    static_assert(KP % 2 == 0, "rows are padded to an even number of entries (16-byte loads)");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * TL::WARPS + wib;
    const uint64_t n0 = (uint64_t)a.lo + tile * 32;
    if (n0 >= a.hi) return;                                    // whole warp leaves together

    const int nvalid = (int)min((uint64_t)32, (uint64_t)a.hi - n0);
    const uint32_t node = (uint32_t)n0 + lane;
    const bool valid = lane < nvalid;
    float y[DP], g[DP];
    uint32_t rc[KP];
    float cm[KP];
    uint32_t chb[(KP + 3) / 4];
    float inv_s2 = 1.0f;
    int T = 0;
    // ---------------- the node's row: KP/2 vector loads straight into registers
    {
        const uint4 *rp = reinterpret_cast<const uint4 *>(a.rowpack + (size_t)(valid ? node : (uint32_t)n0) * KP);
#pragma unroll
        for (int h = 0; h < KP / 2; h++) {
            const uint4 t = __ldg(rp + h);
            rc[2 * h] = t.x; cm[2 * h] = __uint_as_float(t.y);
            rc[2 * h + 1] = t.z; cm[2 * h + 1] = __uint_as_float(t.w);
        }
#pragma unroll
        for (int w = 0; w < (KP + 3) / 4; w++) chb[w] = 0x7f7f7f7fu;
        if (valid) {
            load_row<DP>(a.y_snap, node, y);
            inv_s2 = __ldcs(a.inv_s2 + node);
            const float u = node_uniform(node, a.ukey);
            int prev = 0;
#pragma unroll
            for (int m = 0; m < KP; m++) {
                const int ch = cum_ceil(a.kappa, cm[m], u);    // pads have cum == 1: ch == T, they never fire
                chb[m >> 2] = (chb[m >> 2] & ~(0xffu << (8 * (m & 3)))) | ((uint32_t)ch << (8 * (m & 3)));
                if (a.fired != nullptr && ch > prev)           // push the firing count to the destination's in-edge slot
                    a.fired[__ldg(a.erank + (size_t)node * KP + m)] = (unsigned char)(ch - prev);
                prev = ch;
            }
            T = prev;
        } else {
#pragma unroll
            for (int m = 0; m < KP; m++) rc[m] = ANNEMBED_NO_NODE;
        }
    }
    // ---------------- the node's own firings
    const uint32_t nkey = neg_stream_key<HUB>(a, node);
    // range of the ids the rejection test can match (node, its neighbours): with the locality relabelling it is narrow,
    // so that almost every negative is accepted by two comparisons
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
    // edge m of the row: neighbour and probability interval by select chains (the row is in registers)
    auto edge = [&](int m, uint32_t &j, float &P_lo, float &P_hi) {
        j = rc[0]; P_hi = cm[0]; P_lo = 0.0f;
#pragma unroll
        for (int mm = 1; mm < KP; mm++) {
            const bool t = m >= mm;
            j = t ? rc[mm] : j; P_hi = t ? cm[mm] : P_hi; P_lo = t ? cm[mm - 1] : P_lo;
        }
    };
    if constexpr (DP <= 4) {
        // Software pipelined over two register sets: the 6 row gathers (y_j + 5 negatives) of firings s+1 and s+2 are
        // in flight during the arithmetic of firing s.
        struct Pre {
            int m;                       // edge of the firing
            float pe;
            unsigned use;                // negatives that found an acceptable node
            float yj[DP], yk[ANNEMBED_NB_NEG][DP];
        };
        // The 4 lanes of an aligned group share their negative streams (neg_stream_key): a chunk of 4 consecutive
        // firings needs 5 Philox blocks per group (one per firing and the block of the fifth words).  Lane r computes the
        // block of firing 4c+r and they are exchanged by shuffles; the fifth-words block is computed by lane 3 when the
        // chunk has at most 3 firings (the common case: 1 block per lane and mini-epoch), else by every lane.
        // The loop trip count is made warp-uniform for the shuffles.
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
            const int m = edge_of_firing<KP>(chb, s);          // < row length because ch[last] == T > s
            P.m = m;
            uint32_t j; float P_lo, P_hi;
            edge(m, j, P_lo, P_hi);
            P.pe = F_SUB(P_hi, P_lo);
            load_row<DP>(a.y_snap, j, P.yj);
            uint32_t negs[ANNEMBED_NB_NEG];
            draw_negatives_v2<HUB>(a, a.epoch, node, (uint32_t)s, An, rejector(j), negs);
            P.use = 0;
#pragma unroll
            for (int q = 0; q < ANNEMBED_NB_NEG; q++) {
                const bool ok = negs[q] != ANNEMBED_NO_NODE;
                P.use |= ok ? (1u << q) : 0u;
                load_row<DP>(a.y_snap, ok ? negs[q] : node, P.yk[q]);
            }
        };
        int m_last = -1;                 // edge whose partner copy yl is live (pair simulation across firings)
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
        if (Tmax > 1) { fetch(1); if (T > 1) prepare(1, PB); }
        for (int s = 0; s < Tmax; s += 2) {
            if (s < T) apply(PA);
            if (s + 2 < Tmax) { fetch(s + 2); if (s + 2 < T) prepare(s + 2, PA); }
            if (s + 1 < T) apply(PB);
            if (s + 3 < Tmax) { fetch(s + 3); if (s + 3 < T) prepare(s + 3, PB); }
        }
    } else {
        // wide rows: the prefetch registers do not fit, plain sequential firings
        int m_prev = -1;
        uint32_t j = 0;
        float pe = 0.0f;
        float yj[DP];
        for (int s = 0; s < T; s++) {
            const int m = edge_of_firing<KP>(chb, s);
            if (m != m_prev) {
                float P_lo, P_hi;
                edge(m, j, P_lo, P_hi);
                pe = F_SUB(P_hi, P_lo);
                load_row<DP>(a.y_snap, j, yj);
                m_prev = m;
            }
            const Philox4 A = philox4x32_10(nkey, (uint32_t)s, a.epoch, 1u, a.k0, a.k1);
            uint32_t negs[ANNEMBED_NB_NEG];
            draw_negatives_v2<HUB>(a, a.epoch, node, (uint32_t)s, A, rejector(j), negs);
            apply_firing<DP, true>(a, node, y, yj, g, pe, inv_s2, negs);
        }
    }
    if (valid) store_row<DP>(a.y_next, node, y);
    unsigned int applied = (unsigned int)T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, o);
    if (lane == 0 && applied) atomicAdd(sample_counter + (blockIdx.x & 255), (unsigned long long)applied);
}

#include "async_sweep.cuh"    // the asynchronous form of K4 (default on one rank)
#include "sweep_events_cp.cuh" // ... its kappa <= 1 kernel for layouts of dimension <= 4, pipelined through cp.async groups

#ifndef ANNEMBED_WARPS_IN
#define ANNEMBED_WARPS_IN 4
#endif
#ifndef ANNEMBED_MINB_IN
#define ANNEMBED_MINB_IN 10
#endif
// shared memory of one in-edge warp: a ring of fired in-edges (SoA, 64 slots) and the running composite per owner
template <int DP>
struct InTile {
    static constexpr int QCAP = 64;                              // >= 31 left over + 32 new entries
    static constexpr int WORDS = QCAP * (4 + DP) + 32 * (1 + DP);
    static constexpr int PER_WARP = WORDS * 4;
    static constexpr int SMEM = ANNEMBED_WARPS_IN * PER_WARP;
    static constexpr int MINB = DP <= 2 ? ANNEMBED_MINB_IN : (DP <= 4 ? 8 : (DP <= 8 ? 6 : (DP <= 16 ? 4 : 2)));
};

// Per-warp state of the in-edge kernels (both variants): the tile's owners, the ring of fired in-edges and the dense stage.
template <int DP, bool FLAGS>
struct InWarp {
    static constexpr int QCAP = InTile<DP>::QCAP;
    const EpochArgs &a;
    int lane;
    uint32_t *q_q, *q_c;           // in-edge index in the tile, firings
    float *q_pe, *q_is2, *q_ys;    // edge probability, 1/s_src^2 (replay variant only), source row [DP][QCAP]
    float *t_alpha, *t_beta;       // running composite of each owner: y -> t_alpha * y + t_beta ([32], [DP][32])
    float y[DP];                   // the owner's position after its own firings (written by k_epoch_out)
    uint32_t rel_lo;               // first in-edge of this lane's node relative to the tile's first one
    const uint4 *rec0;             // record of the tile's first in-edge
    uint32_t q_head = 0;           // ring start (warp-uniform)
    int q_n = 0;                   // queued entries (warp-uniform)

    __device__ __forceinline__ InWarp(const EpochArgs &a_, unsigned char *smem, int lane_) : a(a_), lane(lane_)
    {
        q_q = reinterpret_cast<uint32_t *>(smem);
        q_c = q_q + QCAP;
        q_pe = reinterpret_cast<float *>(q_c + QCAP);
        q_is2 = q_pe + QCAP;
        q_ys = q_is2 + QCAP;
        t_alpha = q_ys + DP * QCAP;
        t_beta = t_alpha + 32;
    }

    // ---- dense stage: `cnt` queued in-edges starting at ring position `head`, lane = queue entry.  Each becomes the
    // affine map y -> alpha y + beta of its owner (coefficient at the owner's position after k_epoch_out); the maps of
    // one owner are adjacent (the sweep is in in-edge order) and are composed by a segmented scan; the last lane of a
    // segment folds the composite into the owner's running composite.
    __device__ __forceinline__ void dense(uint32_t head, int cnt)
    {
        const bool act = lane < cnt;
        const uint32_t e = (head + (uint32_t)lane) & (QCAP - 1);
        uint32_t own = 32u + (uint32_t)lane;                   // inactive lanes: a segment of their own
        float alpha = 1.0f, beta[DP];
#pragma unroll
        for (int cc = 0; cc < DP; cc++) beta[cc] = 0.0f;
        int c = 0;
        float pe = 0.0f, is2 = 0.0f, ys[DP];
#pragma unroll
        for (int cc = 0; cc < DP; cc++) ys[cc] = 0.0f;
        uint32_t q = 0;
        if (act) {
            q = q_q[e]; c = (int)q_c[e];
            if constexpr (FLAGS) {                             // probability interval and source scale from the record
                const uint4 r = __ldg(rec0 + q);
                pe = F_SUB(__uint_as_float(r.z), __uint_as_float(r.y)); is2 = __uint_as_float(r.w);
            } else {
                pe = q_pe[e]; is2 = q_is2[e];
            }
#pragma unroll
            for (int cc = 0; cc < DP; cc++) ys[cc] = q_ys[cc * QCAP + e];
        }
        {   // owner lane: binary search over the lanes' first in-edge (5 shuffles)
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
        if (tail) {                                            // one tail per owner and stage: no conflicts
            const float at = t_alpha[own];
#pragma unroll
            for (int cc = 0; cc < DP; cc++) t_beta[cc * 32 + own] = F_FMA(alpha, t_beta[cc * 32 + own], beta[cc]);
            t_alpha[own] = F_MUL(alpha, at);
        }
        __syncwarp();
    }

    // the lanes with c > 0 append their in-edge (index q in the tile) to the ring, in lane order; a full warp of queued
    // entries is applied at once
    __device__ __forceinline__ void push(int c, uint32_t q, float pe, float is2, const float (&ys)[DP])
    {
        const unsigned fired = __ballot_sync(0xffffffffu, c > 0);
        if (fired == 0u) return;                               // warp-uniform
        if (c > 0) {
            const uint32_t e = (q_head + (uint32_t)q_n + (uint32_t)__popc(fired & ((1u << lane) - 1u))) & (QCAP - 1);
            q_q[e] = q; q_c[e] = (uint32_t)c;
            if constexpr (!FLAGS) { q_pe[e] = pe; q_is2[e] = is2; }
#pragma unroll
            for (int cc = 0; cc < DP; cc++) q_ys[cc * QCAP + e] = ys[cc];
        }
        q_n += __popc(fired);
        __syncwarp();
        if (q_n >= 32) {
            dense(q_head, 32);
            q_head = (q_head + 32u) & (QCAP - 1);
            q_n -= 32;
        }
    }

    // apply the owners' composites and publish the rows (own replica, then the other ranks' replicas)
    __device__ __forceinline__ void finish(uint32_t node, bool valid)
    {
        if (q_n > 0) dense(q_head, q_n);
        if (valid) {
            const float at = t_alpha[lane];
#pragma unroll
            for (int c = 0; c < DP; c++) y[c] = F_FMA(at, y[c], t_beta[c * 32 + lane]);
            store_row<DP>(a.y_next, node, y);
            // fused exchange: the owner writes its row straight into every peer's replica (NVLink P2P stores, coalesced
            // 32 rows per warp; ONE store when peer_next[0] is a multicast mapping) while other tiles are still computing
            for (uint32_t pr = 0; pr < a.n_peers; pr++) store_row<DP>(a.peer_next[pr], node, y);
        }
    }
};

// common prologue of the in-edge kernels.  Returns false when the tile has no in-edge (rows already final).
template <int DP, bool FLAGS>
__device__ __forceinline__ bool in_tile_begin(const EpochArgs &a, InWarp<DP, FLAGS> &W, uint64_t n0, int nvalid, bool valid,
                                              uint32_t node, uint64_t &Q0, uint32_t &n_in)
{
#pragma unroll
    for (int c = 0; c < DP; c++) W.y[c] = 0.0f;
    if (valid) load_row<DP>(a.y_next, node, W.y);
    const uint64_t my_q0 = a.in_ptr[(valid ? node : (uint32_t)n0) - a.lo];
    const uint64_t Q1 = a.in_ptr[(uint32_t)n0 - a.lo + nvalid];
    Q0 = __shfl_sync(0xffffffffu, my_q0, 0);
    // first in-edge of this lane's node relative to the tile's first one (non-decreasing over the lanes): the owner
    // of in-edge q is the last lane whose value is <= q
    W.rel_lo = valid ? (uint32_t)(my_q0 - Q0) : 0xffffffffu;
    n_in = (uint32_t)(Q1 - Q0);                                // in-edges of the tile
    W.rec0 = a.in_rec + (Q0 - a.in_base);
    if (n_in == 0) {                                           // nothing to apply: y_next already holds the result
        if (valid)
            for (uint32_t pr = 0; pr < a.n_peers; pr++) store_row<DP>(a.peer_next[pr], node, W.y);
        return false;
    }
    W.t_alpha[W.lane] = 1.0f;
#pragma unroll
    for (int c = 0; c < DP; c++) W.t_beta[c * 32 + W.lane] = 0.0f;
    return true;
}

// ---- single rank: the firing counts were pushed by k_epoch_out into the byte map ----------------------------------
template <int DP>
__global__ void __launch_bounds__(ANNEMBED_WARPS_IN * 32, InTile<DP>::MINB)
k_epoch_in_flags(EpochArgs a)
{
    using TL = InTile<DP>;
    constexpr int RC = DP <= 2 ? 8 : (DP <= 4 ? 4 : (DP <= 8 ? 2 : 1));     // rounds of 32 slots in flight per chunk
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * ANNEMBED_WARPS_IN + wib;
    const uint64_t n0 = (uint64_t)a.lo + tile * 32;
    if (n0 >= a.hi) return;
    const int nvalid = (int)min((uint64_t)32, (uint64_t)a.hi - n0);
    const uint32_t node = (uint32_t)n0 + lane;
    const bool valid = lane < nvalid;
    InWarp<DP, true> W(a, smem_raw + (size_t)wib * TL::PER_WARP, lane);
    uint64_t Q0; uint32_t n_in;
    if (!in_tile_begin<DP, true>(a, W, n0, nvalid, valid, node, Q0, n_in)) return;
    unsigned char *fb = a.fired + (Q0 - a.in_base);
    __syncwarp();
    for (uint32_t base = 0; base < n_in; base += 32u * RC) {
        const uint32_t cnt = min(32u * RC, n_in - base);
        uint32_t fl[RC], src[RC];
        float ys[RC][DP];
        // firing counts of RC x 32 slots (coalesced byte loads), cleared for the next mini-epoch
#pragma unroll
        for (int r = 0; r < RC; r++) {
            const uint32_t i = 32u * r + lane;
            fl[r] = i < cnt ? (uint32_t)fb[base + i] : 0u;
        }
#pragma unroll
        for (int r = 0; r < RC; r++)
            if (fl[r]) fb[base + 32u * r + lane] = 0;
        // source of the fired slots, then its row: all RC rounds in flight together
#pragma unroll
        for (int r = 0; r < RC; r++) {
            src[r] = 0u;
            if (fl[r]) src[r] = __ldg(reinterpret_cast<const uint32_t *>(W.rec0 + base + 32u * r + lane));
        }
#pragma unroll
        for (int r = 0; r < RC; r++) {
#pragma unroll
            for (int cc = 0; cc < DP; cc++) ys[r][cc] = 0.0f;
            if (fl[r]) load_row<DP>(a.y_snap, src[r], ys[r]);
        }
#pragma unroll
        for (int r = 0; r < RC; r++) {
            if (32u * r >= cnt) break;                         // warp-uniform
            W.push((int)fl[r], base + 32u * r + lane, 0.0f, 0.0f, ys[r]);
        }
    }
    W.finish(node, valid);
}

// ---- multi rank: the sources' decisions are replayed from the in-edge records ----------------------------------------
template <int DP>
__global__ void __launch_bounds__(ANNEMBED_WARPS_IN * 32, InTile<DP>::MINB)
k_epoch_in(EpochArgs a)
{
    using TL = InTile<DP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * ANNEMBED_WARPS_IN + wib;
    const uint64_t n0 = (uint64_t)a.lo + tile * 32;
    if (n0 >= a.hi) return;
    const int nvalid = (int)min((uint64_t)32, (uint64_t)a.hi - n0);
    const uint32_t node = (uint32_t)n0 + lane;
    const bool valid = lane < nvalid;
    InWarp<DP, false> W(a, smem_raw + (size_t)wib * TL::PER_WARP, lane);
    uint64_t Q0; uint32_t n_in;
    if (!in_tile_begin<DP, false>(a, W, n0, nvalid, valid, node, Q0, n_in)) return;
    const uint4 *recp = W.rec0 + lane;

    // ---- sparse stage: lane = in-edge record.  Replays the source's firing decision; fired in-edges (about
    // kappa * p_e of them) are compacted into the ring, the source-row gather of round r+1 and the records of round
    // r+2 are in flight during round r.
    // Two alternating register sets (rounds are processed in pairs): no copies between the pipeline stages.
    struct Prep {
        int c;                                                 // firings of this lane's in-edge (0: not fired / no in-edge)
        float pe, is2, ys[DP];
    };
    auto prepare = [&](const uint4 &rec, bool have, Prep &P) {
        P.c = 0;
        if (have) {
            const float us = node_uniform(rec.x, a.ukey);
            P.c = cum_ceil(a.kappa, as_float(rec.z), us) - cum_ceil(a.kappa, as_float(rec.y), us);
        }
        P.pe = F_SUB(as_float(rec.z), as_float(rec.y)); P.is2 = as_float(rec.w);
        if (P.c > 0) load_row<DP>(a.y_snap, rec.x, P.ys);
    };
    Prep PA, PB;
    PA.c = PB.c = 0; PA.pe = PB.pe = PA.is2 = PB.is2 = 0.0f;
#pragma unroll
    for (int cc = 0; cc < DP; cc++) PA.ys[cc] = PB.ys[cc] = 0.0f;
    uint4 recA = make_uint4(0, 0, 0, 0), recB = make_uint4(0, 0, 0, 0);
    if (lane < n_in) recA = __ldcs(recp);                      // round 0
    if (32 + lane < n_in) recB = __ldcs(recp + 32);            // round 1
    prepare(recA, lane < n_in, PA);
    const uint4 *recp2 = recp + 32;                            // slot of this lane in the round after next
    int left = (int)n_in - 32 - lane;                          // > 0 iff that slot holds an in-edge
    __syncwarp();
    for (uint32_t base_q = 0; base_q < n_in; base_q += 64) {
        // even round: PA is prepared, recB holds the records of the next round
        recp2 += 32; left -= 32;                               // `left` = in-edges from this lane's slot two rounds ahead
        if (left > 0) recA = __ldcs(recp2);
        prepare(recB, left + 32 > 0, PB);
        W.push(PA.c, base_q + (uint32_t)lane, PA.pe, PA.is2, PA.ys);
        if (base_q + 32 >= n_in) break;
        // odd round: PB is prepared, recA holds the records of the next round
        recp2 += 32; left -= 32;
        if (left > 0) recB = __ldcs(recp2);
        prepare(recA, left + 32 > 0, PA);
        W.push(PB.c, base_q + 32 + (uint32_t)lane, PB.pe, PB.is2, PB.ys);
    }
    W.finish(node, valid);
}

// K5: embedder.rs:1127-1163 + cauchy_edge_weight :1322-1345, fp64 like the reference
template <int DP>
__global__ void __launch_bounds__(256)
k_cross_entropy(uint32_t lo, uint32_t hi, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                const float *__restrict__ p, const float *__restrict__ emb_scale, const float *__restrict__ Y, double b,
                double *__restrict__ partials)
{
    double acc = 0.0;
    for (uint64_t i = (uint64_t)lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
        float yi[DP];
        load_row<DP>(Y, (uint32_t)i, yi);
        const double s = (double)emb_scale[i];
        const double s2 = s * s;
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) {
            float yj[DP];
            load_row<DP>(Y, col[m], yj);
            float ds = 0.0f;
#pragma unroll
            for (int c = 0; c < DP; c++) { const float t = __fsub_rn(yi[c], yj[c]); ds = __fadd_rn(ds, __fmul_rn(t, t)); }
            double x = (double)ds / s2;
            x = (b == 1.0) ? x : pow(x, b);
            float wf = (float)(1.0 / (1.0 + x));
            if (!(wf < 1.0f)) wf = 1.0f - 1.1920929e-7f;                                 // :1338-1341
            const double w = (double)wf, pe = (double)p[m];
            if (w > 0.0) acc += -pe * log(w);
            if (w < 1.0) acc += -(1.0 - pe) * log(1.0 - w);
        }
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double tot = BR(tmp).Sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

template <bool HUB>
__global__ void k_debug_draws(EpochArgs a, uint32_t *__restrict__ counts, uint32_t *__restrict__ negs_out)
{
    const uint64_t node = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= a.n) return;
    const uint64_t r0 = a.row_ptr[node], r1 = a.row_ptr[node + 1];
    const float u = node_uniform((uint32_t)node, a.ukey);
    int c_lo = 0;
    for (uint64_t m = r0; m < r1; m++) {
        const int c_hi = cum_ceil(a.kappa, a.cum[m], u);
        const int c = c_hi - c_lo;
        counts[m] = (uint32_t)c;
        if (negs_out) {
            uint32_t negs[ANNEMBED_NB_NEG] = {ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE, ANNEMBED_NO_NODE};
            if (c > 0) {
                const uint32_t s = (uint32_t)c_lo;              // the node's firing index at which this edge first fires
                const uint32_t nk = neg_stream_key<HUB>(a, (uint32_t)node);
                const Philox4 A = philox4x32_10(nk, s, a.epoch, 1u, a.k0, a.k1);
                const GlobalRowRejector rej{a.col, r0, r1, (uint32_t)node, a.col[m]};
                draw_negatives_v2<HUB>(a, a.epoch, (uint32_t)node, s, A, rej, negs);
            }
            for (int q = 0; q < ANNEMBED_NB_NEG; q++) negs_out[5 * m + q] = negs[q];
        }
        c_lo = c_hi;
    }
}

// inclusive cumulative edge probability along each row, clamped to 1 and exactly 1 on the last edge; written in the
// relabelled edge order (row i of the internal numbering = row old_of_new[i] of the caller's graph)
__global__ void k_row_cumsum(uint64_t n, const uint32_t *__restrict__ old_of_new, const uint64_t *__restrict__ row_ptr,
                             const float *__restrict__ p, const uint64_t *__restrict__ row_ptr2, float *__restrict__ cum)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = old_of_new[i];
    const uint64_t r0 = row_ptr[o], k = row_ptr[o + 1] - r0, w0 = row_ptr2[i];
    float acc = 0.0f;
    for (uint64_t m = 0; m < k; m++) { acc = __fadd_rn(acc, p[r0 + m]); cum[w0 + m] = fminf(acc, 1.0f); }
    cum[w0 + k - 1] = 1.0f;
}

// =====================================================================================================
// host side
// =====================================================================================================
static int sync_stream(annembed_cuda_ctx *ctx)
{
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return ANNEMBED_OK;
}

static int h2d(annembed_cuda_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->st.h2d_bytes += bytes;
    return ANNEMBED_OK;
}
static int d2h(annembed_cuda_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += bytes;
    return ANNEMBED_OK;
}

static int sum_f64(annembed_cuda_ctx *ctx, const float *x, uint64_t n, double *d_out /* device, partials[4095] */)
{
    const unsigned int nb = std::min<unsigned int>(2048u, std::max(1u, nblocks(n, 256)));
    k_partial_sum_f64<float><<<nb, 256, 0, ctx->stream>>>(n, x, ctx->partials.p);
    k_final_sum_f64<<<1, 256, 0, ctx->stream>>>(nb, ctx->partials.p, d_out);
    ctx->st.kernel_launches += 2;
    CU(cudaGetLastError());
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_default_params(annembed_cuda_params *p)
{
    if (!p) return ANNEMBED_ERR_INVALID_ARG;
    memset(p, 0, sizeof(*p));
    p->asked_dim = 2; p->dmap_init = 1; p->beta = 1.0; p->b = 1.0; p->scale_rho = 1.0; p->grad_step = 2.0;   // embedparams.rs:107-132
    p->nb_sampling_by_edge = 10; p->nb_grad_batch = 20; p->grad_factor = 4; p->hierarchy_layer = 0; p->hubness_weighting = 0;
    p->mini_epochs_per_batch = 0; p->seed = 0x5eedULL; p->flags = 0; p->cell_substeps = 0;
    return ANNEMBED_OK;
}

extern "C" const char *annembed_cuda_last_error(const annembed_cuda_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int annembed_cuda_create(annembed_cuda_ctx **out, const annembed_cuda_params *params, int device)
{
    if (!out || !params) { g_create_error = "null argument"; return ANNEMBED_ERR_INVALID_ARG; }
    *out = nullptr;
    if (params->asked_dim == 0 || params->nb_grad_batch == 0 || params->nb_sampling_by_edge == 0 ||
        !(params->b > 0.0) || !(params->beta > 0.0) || !(params->scale_rho > 0.0)) {
        g_create_error = "invalid EmbedderParams (asked_dim, nb_grad_batch, nb_sampling_by_edge, b, beta, scale_rho must be > 0)";
        return ANNEMBED_ERR_INVALID_ARG;
    }
    if (params->asked_dim > 32) { g_create_error = "asked_dim > 32 not supported on device"; return ANNEMBED_ERR_UNSUPPORTED; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return ANNEMBED_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "bad device ordinal"; return ANNEMBED_ERR_INVALID_ARG; }
    annembed_cuda_ctx *ctx = new annembed_cuda_ctx();
    ctx->prm = *params;
    ctx->device = device;
    ctx->DP = pad_dim(params->asked_dim);
    auto fail = [&](const char *what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete ctx;
        return ANNEMBED_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    ctx->launch_stream = ctx->stream;
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    {   // keep freed blocks in the stream-ordered pool (see DevBuf)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaDeviceGetAttribute(&ctx->l2_persist_max, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->l2_window_max, cudaDevAttrMaxAccessPolicyWindowSize, device);
    cudaDeviceGetAttribute(&ctx->smem_optin_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (ctx->l2_persist_max > 0 && !(ctx->prm.flags & ANNEMBED_FLAG_NO_L2_PERSIST))
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)ctx->l2_persist_max);
    if ((e = ctx->partials.alloc(4096)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = ctx->counter.alloc(256)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = ctx->errword.alloc(2)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaEventCreate(&ctx->ev_a)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_b)) != cudaSuccess) return fail("cudaEventCreate", e);
    *out = ctx;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_destroy(annembed_cuda_ctx *ctx)
{
    if (!ctx) return ANNEMBED_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    close_peers(ctx);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (auto ev : ctx->ev) cudaEventDestroy(ev);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_comm_unique_id(uint8_t unique_id[128])
{
    if (!unique_id) return ANNEMBED_ERR_INVALID_ARG;
    if (!g_nccl.load(g_create_error)) return ANNEMBED_ERR_COMM;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return ANNEMBED_ERR_COMM; }
    memcpy(unique_id, &id, 128);
    return ANNEMBED_OK;
}

static void set_shard(annembed_cuda_ctx *ctx)
{
    const uint64_t n = ctx->n;
    // shards are whole warp tiles (32 nodes): every node then sits in the same tile, and its in-edges in the same
    // sweep rounds, for any number of ranks -> bit-identical layouts for any R
    ctx->n_pad = (uint32_t)((((n + ctx->nranks - 1) / ctx->nranks) + 31) / 32 * 32);
    ctx->lo = (uint32_t)std::min<uint64_t>(n, (uint64_t)ctx->rank * ctx->n_pad);
    ctx->hi = (uint32_t)std::min<uint64_t>(n, (uint64_t)(ctx->rank + 1) * ctx->n_pad);
}

extern "C" int annembed_cuda_comm_init(annembed_cuda_ctx *ctx, int rank, int nranks, const uint8_t unique_id[128])
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, ANNEMBED_ERR_INVALID_ARG, "bad rank / nranks");
    REQUIRE(!ctx->have_graph, ANNEMBED_ERR_STATE, "comm_init must precede set_graph_csr");
    CU(cudaSetDevice(ctx->device));
    ctx->rank = rank; ctx->nranks = nranks;
    if (nranks == 1) return ANNEMBED_OK;
    REQUIRE(unique_id != nullptr, ANNEMBED_ERR_INVALID_ARG, "unique_id is null");
    if (!g_nccl.load(ctx->err)) return ANNEMBED_ERR_COMM;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != ncclSuccess) { ctx->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); ctx->comm = nullptr; return ANNEMBED_ERR_COMM; }
    return ANNEMBED_OK;
}

static void close_peers(annembed_cuda_ctx *ctx)
{
    if (!ctx->have_peers) return;
    for (int r = 0; r < ctx->nranks && r < 8; r++)
        for (int b = 0; b < 2; b++)
            if (r != ctx->rank && ctx->peer_y[r][b]) { cudaIpcCloseMemHandle(ctx->peer_y[r][b]); ctx->peer_y[r][b] = nullptr; }
    ctx->have_peers = false;
}

// the two layout buffers + the initial copy; sized for whole tiles per rank so that the all-gather is uniform
static int alloc_layout(annembed_cuda_ctx *ctx)
{
    const uint64_t rows = (uint64_t)ctx->n_pad * ctx->nranks;
    const size_t want = (size_t)rows * ctx->DP;
    if (ctx->y[0].n == want) return ANNEMBED_OK;
    close_peers(ctx);
    // multi-rank: plain cudaMalloc (the fused exchange exports these buffers through CUDA IPC, which pool memory
    // does not support); single rank: pool memory, so that creating and destroying a context per embed() costs no
    // cudaMalloc/cudaFree (a cudaFree of these buffers was measured at 80-900 ms)
    const bool ipc = ctx->nranks > 1;
    CU(ctx->y[0].alloc(want, ipc)); CU(ctx->y[1].alloc(want, ipc)); CU(ctx->y0.alloc(want)); CU(ctx->yapi.alloc(want));
    CU(cudaMemsetAsync(ctx->y[0].p, 0, want * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(ctx->y[1].p, 0, want * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(ctx->y0.p, 0, want * sizeof(float), ctx->stream));
    CU(cudaMemsetAsync(ctx->yapi.p, 0, want * sizeof(float), ctx->stream));
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_set_graph_csr(annembed_cuda_ctx *ctx, uint64_t n, const uint64_t *row_ptr,
                                           const uint32_t *col, const float *dist)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(row_ptr && col && dist, ANNEMBED_ERR_INVALID_ARG, "null graph pointer");
    REQUIRE(n >= 1, ANNEMBED_ERR_INVALID_ARG, "graph has no node");
    REQUIRE(n < 0xFFFFFFFFull, ANNEMBED_ERR_UNSUPPORTED, "n >= 2^32-1 not supported");
    REQUIRE(row_ptr[0] == 0, ANNEMBED_ERR_INVALID_ARG, "row_ptr[0] != 0");
    const uint64_t E = row_ptr[n];
    REQUIRE(E < 0xFFFFFFFFull, ANNEMBED_ERR_UNSUPPORTED, "E >= 2^32-1 not supported");
    REQUIRE(E >= 1, ANNEMBED_ERR_EMPTY_ROW, "graph has no edge (node 0 has no neighbour)");
    CU(cudaSetDevice(ctx->device));
    ctx->have_graph = ctx->have_weights = ctx->have_build = ctx->have_embedding = ctx->have_struct = false;
    ctx->n = n; ctx->E = E;
    set_shard(ctx);
    CU(ctx->row_ptr.alloc(n + 1)); CU(ctx->col.alloc(E)); CU(ctx->dist.alloc(E)); CU(ctx->rho.alloc(n));
    CU(ctx->scale.alloc(n)); CU(ctx->proba.alloc(E));
    int rc;
    if ((rc = alloc_layout(ctx))) return rc;
    if (ctx->nranks > 1 && ctx->comm) {
        // Several ranks: the call is collective and every rank passes the SAME graph.  Each rank uploads one R-th of every
        // array from its host and the parts travel to the other ranks over NVLink (one grouped broadcast per part): the
        // ranks of a node share the host's memory and PCIe root, where R full uploads took 28 ms at 8 ranks against 12 ms alone.
        struct Part { void *dst; const void *src; size_t bytes; };
        const Part parts[3] = {{ctx->row_ptr.p, row_ptr, (n + 1) * sizeof(uint64_t)}, {ctx->col.p, col, E * sizeof(uint32_t)},
                               {ctx->dist.p, dist, E * sizeof(float)}};
        const size_t R = (size_t)ctx->nranks;
        auto cut = [&](size_t bytes, size_t r) { return r >= R ? bytes : (bytes / R * r) & ~(size_t)15; };
        for (const Part &p : parts) {
            const size_t b0 = cut(p.bytes, (size_t)ctx->rank), b1 = cut(p.bytes, (size_t)ctx->rank + 1);
            if (b1 > b0) {
                CU(cudaMemcpyAsync((char *)p.dst + b0, (const char *)p.src + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->stream));
                ctx->st.h2d_bytes += b1 - b0;
            }
        }
        ncclResult_t r = g_nccl.GroupStart();
        for (const Part &p : parts)
            for (size_t k = 0; k < R && r == ncclSuccess; k++) {
                const size_t b0 = cut(p.bytes, k), b1 = cut(p.bytes, k + 1);
                if (b1 > b0) r = g_nccl.Broadcast((char *)p.dst + b0, (char *)p.dst + b0, b1 - b0, ncclUint8, (int)k, ctx->comm, ctx->stream);
            }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { ctx->err = std::string("ncclBroadcast (graph parts): ") + g_nccl.GetErrorString(r); return ANNEMBED_ERR_COMM; }
        if ((rc = sync_stream(ctx))) return rc;
    } else {
        if ((rc = h2d(ctx, ctx->row_ptr.p, row_ptr, (n + 1) * sizeof(uint64_t)))) return rc;
        if ((rc = h2d(ctx, ctx->col.p, col, E * sizeof(uint32_t)))) return rc;
        if ((rc = h2d(ctx, ctx->dist.p, dist, E * sizeof(float)))) return rc;
    }
    // validate on the device
    unsigned long long init[2] = {~0ull, 0xFFFFFFFF00000000ull};   // error word; {kmax (low), kmin (high)}
    CU(cudaMemcpyAsync(ctx->errword.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    k_validate_graph<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, E, ctx->row_ptr.p, ctx->col.p, ctx->dist.p, ctx->errword.p,
                                                               (unsigned int *)(ctx->errword.p + 1), (unsigned int *)(ctx->errword.p + 1) + 1);
    ctx->st.kernel_launches++;
    unsigned long long res[2];
    CU(cudaMemcpyAsync(res, ctx->errword.p, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    if (res[0] != ~0ull) {
        const unsigned long long node = res[0] >> 4;
        const unsigned int code = (unsigned int)(res[0] & 15);
        char buf[160];
        switch (code) {
        case GERR_EMPTY: snprintf(buf, sizeof buf, "node rank %llu has no neighbour (kdumap.rs:75-85)", node); ctx->err = buf; return ANNEMBED_ERR_EMPTY_ROW;
        case GERR_UNSORTED: snprintf(buf, sizeof buf, "row %llu: distances not ascending / not finite (kgraph.rs:508-509)", node); ctx->err = buf; return ANNEMBED_ERR_UNSORTED_ROW;
        case GERR_COL: snprintf(buf, sizeof buf, "row %llu: neighbour index >= n", node); break;
        case GERR_SELF: snprintf(buf, sizeof buf, "row %llu: self edge (embedder.rs:1201)", node); break;
        default: snprintf(buf, sizeof buf, "row %llu: row_ptr not monotone", node); break;
        }
        ctx->err = buf;
        return ANNEMBED_ERR_INVALID_ARG;
    }
    ctx->kmax = (uint32_t)(res[1] & 0xFFFFFFFFu);
    ctx->kmin = (uint32_t)(res[1] >> 32);
    if (n < (uint64_t)ctx->kmax + 3) {
        ctx->err = "graph too small: no node acceptable as negative sample (embedder.rs:1241-1252 would not terminate)";
        return ANNEMBED_ERR_NO_NEGATIVE;
    }
    k_first_dist<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->row_ptr.p, ctx->dist.p, ctx->rho.p);
    ctx->st.kernel_launches++;
    if ((rc = sync_stream(ctx))) return rc;
    ctx->have_graph = true;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_comm_export_layout(annembed_cuda_ctx *ctx, uint8_t handles[128])
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(handles, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "comm_export_layout: graph not set (the layout buffers are sized by it)");
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    for (int b = 0; b < 2; b++) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ctx->y[b].p));
        memcpy(handles + 64 * b, &h, 64);
    }
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_comm_import_layouts(annembed_cuda_ctx *ctx, const uint8_t *all_handles)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(all_handles, ANNEMBED_ERR_INVALID_ARG, "null handles");
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "comm_import_layouts: graph not set");
    REQUIRE(ctx->nranks > 1 && ctx->comm, ANNEMBED_ERR_STATE, "comm_import_layouts: comm_init with nranks > 1 first");
    REQUIRE(ctx->nranks <= 8, ANNEMBED_ERR_UNSUPPORTED, "fused exchange supports up to 8 ranks (one NVSwitch domain)");
    CU(cudaSetDevice(ctx->device));
    close_peers(ctx);
    for (int r = 0; r < ctx->nranks; r++)
        for (int b = 0; b < 2; b++) {
            if (r == ctx->rank) { ctx->peer_y[r][b] = ctx->y[b].p; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + (size_t)r * 128 + 64 * b, 64);
            void *p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->peer_y[r][b] = (float *)p;
        }
    CU(ctx->barrier_buf.alloc(4));
    CU(cudaMemsetAsync(ctx->barrier_buf.p, 0, 4 * sizeof(float), ctx->stream));
    ctx->have_peers = true;
    ctx->have_struct = false; ctx->have_build = false;      // the internal numbering depends on the form of K4 (use_async)
    return sync_stream(ctx);
}

extern "C" int annembed_cuda_edge_weights(annembed_cuda_ctx *ctx, float *scale_out, float *proba_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "edge_weights: graph not set");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    k_edge_weights<<<nblocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->row_ptr.p, ctx->col.p, ctx->dist.p, ctx->rho.p,
                                                                  (float)ctx->prm.scale_rho, (float)ctx->prm.beta,
                                                                  ctx->scale.p, ctx->proba.p);
    ctx->st.kernel_launches++;
    CU(cudaEventRecord(ctx->ev_b, ctx->stream));
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    ctx->st.edge_weights_ms = ms;
    ctx->have_weights = true; ctx->have_build = false;
    if (scale_out && (rc = d2h(ctx, scale_out, ctx->scale.p, ctx->n * sizeof(float)))) return rc;
    if (proba_out && (rc = d2h(ctx, proba_out, ctx->proba.p, ctx->E * sizeof(float)))) return rc;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_edge_weights_umap(annembed_cuda_ctx *ctx, float norm, float *scale_out, float *weight_out,
                                               uint8_t *status_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "edge_weights_umap: graph not set");
    REQUIRE(scale_out && weight_out, ANNEMBED_ERR_INVALID_ARG, "null output");
    CU(cudaSetDevice(ctx->device));
    DevBuf<float> s, w; DevBuf<uint8_t> st;
    CU(s.alloc(ctx->n)); CU(w.alloc(ctx->E)); CU(st.alloc(ctx->n));
    k_edge_weights_umap<<<nblocks(ctx->n, 128), 128, 0, ctx->stream>>>(ctx->n, ctx->row_ptr.p, ctx->dist.p, norm, s.p, w.p, st.p);
    ctx->st.kernel_launches++;
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    if ((rc = d2h(ctx, scale_out, s.p, ctx->n * sizeof(float)))) return rc;
    if ((rc = d2h(ctx, weight_out, w.p, ctx->E * sizeof(float)))) return rc;
    if (status_out && (rc = d2h(ctx, status_out, st.p, ctx->n))) return rc;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_set_edge_weights(annembed_cuda_ctx *ctx, const float *scale, const float *proba)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "set_edge_weights: graph not set");
    REQUIRE(scale && proba, ANNEMBED_ERR_INVALID_ARG, "null weights");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = h2d(ctx, ctx->scale.p, scale, ctx->n * sizeof(float)))) return rc;
    if ((rc = h2d(ctx, ctx->proba.p, proba, ctx->E * sizeof(float)))) return rc;
    ctx->have_weights = true; ctx->have_build = false;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_get_perplexity(annembed_cuda_ctx *ctx, float *out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_weights, ANNEMBED_ERR_STATE, "get_perplexity: edge weights not computed");
    REQUIRE(out, ANNEMBED_ERR_INVALID_ARG, "null output");
    CU(cudaSetDevice(ctx->device));
    DevBuf<float> t; CU(t.alloc(ctx->n));
    k_perplexity<<<nblocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->row_ptr.p, ctx->proba.p, t.p);
    ctx->st.kernel_launches++;
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    return d2h(ctx, out, t.p, ctx->n * sizeof(float));
}

// Vose alias table on the host (setup, O(n)); ≙ WeightedAliasIndex::new in NodeSampler::new (embedder.rs:916-919)
extern "C" int annembed_cuda_set_neg_weights(annembed_cuda_ctx *ctx, const float *w)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "set_neg_weights: graph not set");
    CU(cudaSetDevice(ctx->device));
    if (!w) { ctx->have_alias = false; ctx->neg_alias.release(); ctx->neg_alias_old.release(); ctx->sec_alias.release(); ctx->line_t1.release(); ctx->line_t2.release(); ctx->neg_w_host.clear(); return ANNEMBED_OK; }
    const uint64_t n = ctx->n;
    double tot = 0.0;
    for (uint64_t i = 0; i < n; i++) {
        REQUIRE(w[i] >= 0.0f && std::isfinite(w[i]), ANNEMBED_ERR_INVALID_ARG, "negative / non-finite sampling weight");
        tot += w[i];
    }
    REQUIRE(tot > 0.0, ANNEMBED_ERR_INVALID_ARG, "all sampling weights are zero");
    std::vector<uint2> tab;
    annembed_host::build_node_alias_table(n, [&](uint64_t i) { return (double)w[i]; }, tot, tab);
    CU(ctx->neg_alias_old.alloc(n));
    int rc;
    if ((rc = h2d(ctx, ctx->neg_alias_old.p, tab.data(), n * sizeof(uint2)))) return rc;
    ctx->neg_w_host.assign(w, w + n);
    ctx->have_alias = true; ctx->alias_dirty = true;
    return ANNEMBED_OK;
}

// Internal numbering of the nodes and its cells.
//  1. locality order: identity (ANNEMBED_FLAG_NO_RELABEL, tiny graphs) or the nested min-label order of k_lp_round;
//  2. cells: consecutive ranges of at most cell_nodes nodes of that order, starting at multiples of 32 (whole warp tiles),
//     cut at the boundaries of the top-level label's segments where possible (a connected component of at most
//     cell_nodes nodes is never split): greedy packing of the segments, on the host (the segment list is short);
//  3. the order inside every cell is then randomised (k_cell_sort_keys explains why);
//  4. ranks own whole cells (balanced node counts).  Cells do not depend on the number of ranks.
// Graph only; built once per set_graph_csr.
struct ShardBounds { uint32_t lo[9]; uint32_t nranks; };
// sort key of position `pos` of the locality order: (the rank that owns it, a hash of the node)
__global__ void k_shard_sort_keys(uint64_t n, ShardBounds sb, const uint32_t *__restrict__ order, unsigned long long *__restrict__ key)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = 0;
    for (uint32_t q = 1; q < sb.nranks; q++) r += (uint32_t)i >= sb.lo[q] ? 1u : 0u;
    key[i] = ((unsigned long long)r << 32) | mix32(order[i] * 0x9E3779B1u + 0x7F4A7C15u);
}
__global__ void __launch_bounds__(256)
k_count_cross_rank(uint64_t n, ShardBounds sb, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, unsigned long long *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int c = 0;
    if (i < n) {
        uint32_t ri = 0;
        for (uint32_t q = 1; q < sb.nranks; q++) ri += (uint32_t)i >= sb.lo[q] ? 1u : 0u;
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) {
            uint32_t rj = 0;
            for (uint32_t q = 1; q < sb.nranks; q++) rj += col[m] >= sb.lo[q] ? 1u : 0u;
            c += rj != ri ? 1u : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}
__global__ void k_hash_keys(uint64_t n, uint32_t salt, uint32_t *__restrict__ key)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = mix32((uint32_t)i * 0x9E3779B1u + salt);
}
static bool use_async(const annembed_cuda_ctx *ctx);
static int build_relabelling(annembed_cuda_ctx *ctx)
{
    const uint64_t n = ctx->n;
    int rc;
    const uint32_t CELL = ctx->cell_nodes = (uint32_t)(ctx->DP <= 2 ? CellCfg<2>::CELL : CellCfg<4>::CELL);
    CU(ctx->new_of_old.alloc(n)); CU(ctx->old_of_new.alloc(n));
    k_iota<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p);
    ctx->st.kernel_launches++;
    std::vector<uint32_t> seg;                       // segment starts of the locality order (empty: none known)
    bool relabel = !(ctx->prm.flags & ANNEMBED_FLAG_NO_RELABEL) && n >= 1024;
    DevBuf<uint32_t> key, key_s, ord;
    DevBuf<unsigned char> tmp;
    size_t tmp_bytes = 0;
    const bool async_form = use_async(ctx);
    if (relabel && async_form && ctx->nranks == 1) {
        // Asynchronous form: a RANDOM internal order (sort by a hash of the caller's id).  The nodes a warp visits together
        // (a tile), the 4 nodes that share a negative stream and the 4 nodes of a negative's sector must all be unrelated
        // nodes, whatever structure the caller's numbering has: neighbours that move at the same moment, or against the
        // same negatives, move coherently and shift the layout statistics.  The whole layout is the working
        // set of the negatives anyway, so a locality order would buy no cache hits.
        CU(key.alloc(n)); CU(key_s.alloc(n)); CU(ord.alloc(n));
        k_hash_keys<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, 0x7F4A7C15u, key.p);
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key_s.p, ctx->old_of_new.p, ord.p, (int64_t)n, 0, 32, ctx->stream));
        CU(tmp.alloc(tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key_s.p, ctx->old_of_new.p, ord.p, (int64_t)n, 0, 32, ctx->stream));
        CU(cudaMemcpyAsync(ctx->old_of_new.p, ord.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->st.kernel_launches += 2;
        if ((rc = sync_stream(ctx))) return rc;
        relabel = false;                             // no locality order, no segments: fixed grid of cells below
    }
    if (relabel) {
        DevBuf<uint32_t> L[2];
        CU(L[0].alloc(n)); CU(L[1].alloc(n)); CU(key.alloc(n)); CU(key_s.alloc(n)); CU(ord.alloc(n));
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key_s.p, ctx->old_of_new.p, ord.p, (int64_t)n, 0, 32, ctx->stream));
        size_t tb2 = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb2, (unsigned long long *)nullptr, (unsigned long long *)nullptr, ctx->old_of_new.p, ord.p,
                                           (int64_t)n, 0, 64, ctx->stream));
        tmp_bytes = std::max(tmp_bytes, tb2);
        CU(tmp.alloc(tmp_bytes));
        k_lp_init<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, L[0].p);
        int cur = 0;
        uint32_t *order = ctx->old_of_new.p, *order2 = ord.p;
        for (int round = 1; round <= 8; round++) {
            CU(cudaMemcpyAsync(L[cur ^ 1].p, L[cur].p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
            k_lp_round<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->row_ptr.p, ctx->col.p, L[cur].p, L[cur ^ 1].p);
            cur ^= 1;
            ctx->st.kernel_launches += 1;
            if (round == 1 || round == 2 || round == 4 || round == 8) {
                // stable sort of the current order by this level's label (least significant level first)
                k_gather_u32<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, order, L[cur].p, key.p);
                size_t tb = tmp_bytes;
                CU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key_s.p, order, order2, (int64_t)n, 0, 32, ctx->stream));
                std::swap(order, order2);
                ctx->st.kernel_launches += 2;
            }
        }
        if (order != ctx->old_of_new.p)
            CU(cudaMemcpyAsync(ctx->old_of_new.p, order, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        // segments of the top-level label (key_s holds the sorted labels of the last pass)
        DevBuf<unsigned char> flag; DevBuf<uint32_t> seg_d; DevBuf<unsigned long long> nsel;
        CU(flag.alloc(n)); CU(seg_d.alloc(n)); CU(nsel.alloc(1));
        k_seg_flags<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, key_s.p, flag.p);
        size_t tb3 = 0;
        thrust::counting_iterator<uint32_t> pos(0);
        CU(cub::DeviceSelect::Flagged(nullptr, tb3, pos, flag.p, seg_d.p, nsel.p, (int64_t)n, ctx->stream));
        DevBuf<unsigned char> tmp3; CU(tmp3.alloc(tb3));
        CU(cub::DeviceSelect::Flagged(tmp3.p, tb3, pos, flag.p, seg_d.p, nsel.p, (int64_t)n, ctx->stream));
        ctx->st.kernel_launches += 2;
        unsigned long long nseg = 0;
        CU(cudaMemcpyAsync(&nseg, nsel.p, sizeof(nseg), cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_stream(ctx))) return rc;
        if (nseg <= (4u << 20)) {                    // a longer list means no useful segment structure: fixed grid
            seg.resize(nseg);
            CU(cudaMemcpyAsync(seg.data(), seg_d.p, nseg * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            if ((rc = sync_stream(ctx))) return rc;
        }
    }
    // ---- cells: greedy packing of the segments (fixed grid when there is no segment list)
    std::vector<uint32_t> &cs = ctx->cell_start_h;
    cs.clear(); cs.push_back(0);
    {
        uint64_t cur = 0;
        for (size_t i = 0; i < seg.size(); i++) {
            const uint64_t s0 = seg[i], e0 = i + 1 < seg.size() ? seg[i + 1] : n;
            if (e0 - cur <= CELL) continue;          // the segment fits in the open cell
            const uint64_t c = s0 & ~31ull;          // close the open cell in front of it (whole tiles)
            if (c > cur) { cs.push_back((uint32_t)c); cur = c; }
            while (e0 - cur > CELL) { cur += CELL; cs.push_back((uint32_t)cur); }
        }
        while (n - cur > CELL) { cur += CELL; cs.push_back((uint32_t)cur); }
        cs.push_back((uint32_t)n);
    }
    ctx->n_cells = (uint32_t)cs.size() - 1;
    CU(ctx->cell_start.alloc(cs.size()));
    CU(cudaMemcpyAsync(ctx->cell_start.p, cs.data(), cs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (relabel) {
        // random order inside every cell: sort the positions by (cell, hash of the node)
        DevBuf<unsigned long long> k64, k64s;
        CU(k64.alloc(n)); CU(k64s.alloc(n));
        k_cell_sort_keys<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->n_cells, ctx->cell_start.p, ctx->old_of_new.p, k64.p);
        size_t tb = tmp_bytes;
        CU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k64.p, k64s.p, ctx->old_of_new.p, ord.p, (int64_t)n, 0, 64, ctx->stream));
        CU(cudaMemcpyAsync(ctx->old_of_new.p, ord.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->st.kernel_launches += 2;
        if ((rc = sync_stream(ctx))) return rc;     // the scratch buffers are released on return
    }
    // ---- shards: whole cells, balanced by node count
    ctx->shard_lo.assign(ctx->nranks + 1, (uint32_t)n);
    std::vector<uint32_t> shard_cell(ctx->nranks + 1, ctx->n_cells);
    shard_cell[0] = 0; ctx->shard_lo[0] = 0;
    for (int r = 1; r < ctx->nranks; r++) {
        const uint64_t target = n * (uint64_t)r / (uint64_t)ctx->nranks;
        uint32_t c = (uint32_t)(std::lower_bound(cs.begin(), cs.end(), (uint32_t)target) - cs.begin());
        c = std::max(c, shard_cell[r - 1]);
        c = std::min(c, ctx->n_cells);
        shard_cell[r] = c; ctx->shard_lo[r] = cs[c];
    }
    if (async_form && ctx->nranks > 1) {
        // asynchronous form: equal parts of the locality order (whole tiles), so that the row exchange is ONE in-place
        // ncclAllGather; a part boundary may cut one cell
        for (int r = 1; r < ctx->nranks; r++) ctx->shard_lo[r] = (uint32_t)std::min<uint64_t>(n, (uint64_t)r * ctx->n_pad);
    }
    ctx->cell_lo = shard_cell[ctx->rank]; ctx->cell_hi = shard_cell[ctx->rank + 1];
    ctx->lo = ctx->shard_lo[ctx->rank]; ctx->hi = ctx->shard_lo[ctx->rank + 1];
    if (async_form && ctx->nranks > 1 && n >= 1024 && !(ctx->prm.flags & ANNEMBED_FLAG_NO_RELABEL)) {
        // Asynchronous form on several ranks: the ranks keep the graph-local parts of the locality order (few edges
        // cross ranks), the order INSIDE a rank's part is random (see the single-rank case above)
        ShardBounds sb;
        memset(&sb, 0, sizeof sb);
        sb.nranks = (uint32_t)ctx->nranks;
        for (int r = 0; r <= ctx->nranks && r <= 8; r++) sb.lo[r] = ctx->shard_lo[r];
        DevBuf<unsigned long long> k64, k64s;
        DevBuf<uint32_t> ord2;
        DevBuf<unsigned char> tmp2;
        CU(k64.alloc(n)); CU(k64s.alloc(n)); CU(ord2.alloc(n));
        k_shard_sort_keys<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, sb, ctx->old_of_new.p, k64.p);
        size_t tb = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, k64.p, k64s.p, ctx->old_of_new.p, ord2.p, (int64_t)n, 0, 64, ctx->stream));
        CU(tmp2.alloc(tb));
        CU(cub::DeviceRadixSort::SortPairs(tmp2.p, tb, k64.p, k64s.p, ctx->old_of_new.p, ord2.p, (int64_t)n, 0, 64, ctx->stream));
        CU(cudaMemcpyAsync(ctx->old_of_new.p, ord2.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->st.kernel_launches += 2;
        if ((rc = sync_stream(ctx))) return rc;
    }
    k_invert_perm<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p, ctx->new_of_old.p);
    ctx->st.kernel_launches++;
    CU(cudaGetLastError());
    return sync_stream(ctx);
}

// device context build, graph part (≙ nothing in the reference: its symmetric move writes y_j under a lock,
// embedder.rs:1239): internal numbering, relabelled CSR, transposed index of the owned nodes, out-edge -> in-edge slot
// map.  Depends on the graph and the shard only; built once per set_graph_csr.
static int ensure_struct(annembed_cuda_ctx *ctx)
{
    if (ctx->have_struct) return ANNEMBED_OK;
    const uint64_t n = ctx->n, E = ctx->E;
    int rc;
    if ((rc = build_relabelling(ctx))) return rc;
    // relabelled CSR
    CU(ctx->row_ptr2.alloc(n + 1)); CU(ctx->col2.alloc(E));
    {
        DevBuf<uint64_t> deg2; DevBuf<unsigned char> tmp;
        CU(deg2.alloc(n + 1));
        k_relabel_degrees<<<nblocks(n + 1, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p, ctx->row_ptr.p, deg2.p);
        size_t tmp_bytes = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg2.p, ctx->row_ptr2.p, (int64_t)(n + 1), ctx->stream));
        CU(tmp.alloc(tmp_bytes));
        CU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, deg2.p, ctx->row_ptr2.p, (int64_t)(n + 1), ctx->stream));
        k_relabel_rows<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p, ctx->new_of_old.p, ctx->row_ptr.p, ctx->col.p,
                                                                 ctx->row_ptr2.p, ctx->col2.p);
        ctx->st.kernel_launches += 3;
        if ((rc = sync_stream(ctx))) return rc;
        ctx->cross_rank_edges = 0;
        if (ctx->nranks > 1) {
            ShardBounds sb;
            memset(&sb, 0, sizeof sb);
            sb.nranks = (uint32_t)ctx->nranks;
            for (int r = 0; r <= ctx->nranks && r <= 8; r++) sb.lo[r] = ctx->shard_lo[r];
            unsigned long long zero = 0ull, res = 0ull;
            CU(cudaMemcpyAsync(ctx->errword.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream));
            k_count_cross_rank<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, sb, ctx->row_ptr2.p, ctx->col2.p, ctx->errword.p);
            CU(cudaMemcpyAsync(&res, ctx->errword.p, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
            if ((rc = sync_stream(ctx))) return rc;
            ctx->cross_rank_edges = res;
            ctx->st.kernel_launches++;
        }
    }
    ctx->KP = ctx->kmax <= 6 ? 6 : (ctx->kmax <= 8 ? 8 : (ctx->kmax <= 10 ? 10 : (ctx->kmax <= 16 ? 16 : 0)));
    if (use_async(ctx)) {
        // the asynchronous form publishes the destination's move with an atomic: no transposed index, no in-edge records,
        // no slot maps -- only the padded rows (filled by ensure_build)
        ctx->rowpack.release(); ctx->erank.release(); ctx->fired.release(); ctx->in_rec.release(); ctx->in_src.release(); ctx->in_eid.release();
        ctx->in_ptr_all.release(); ctx->ext_ptr.release(); ctx->ext_slot.release();
        ctx->in_cnt = 0; ctx->in_base = 0; ctx->ext_cnt = 0; ctx->ext_cnt_global = 0; ctx->cell_map_bytes = 0;
        if (ctx->KP) {
            CU(ctx->rowpack.alloc((n + 32) * (uint64_t)ctx->KP));
            CU(cudaMemsetAsync(ctx->rowpack.p, 0xff, (n + 32) * (uint64_t)ctx->KP * sizeof(uint2), ctx->stream));
        }
        if ((rc = sync_stream(ctx))) return rc;
        ctx->have_struct = true;
        ctx->struct_async = true;
        ctx->alias_dirty = ctx->have_alias;
        return ANNEMBED_OK;
    }
    ctx->struct_async = false;
    CU(ctx->in_ptr_all.alloc(n + 2));
    DevBuf<uint32_t> eid, dst_sorted, eid_sorted;
    DevBuf<unsigned char> tmp;
    CU(eid.alloc(E)); CU(dst_sorted.alloc(E)); CU(eid_sorted.alloc(E));
    k_iota<<<nblocks(E, 256), 256, 0, ctx->stream>>>(E, eid.p);
    int bits = 1; while (bits < 32 && (1ull << bits) < n) bits++;
    size_t tmp_bytes = 0;
    // stable radix sort of (dst, edge id): in-edges of a node stay in edge-id order
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->col2.p, dst_sorted.p, eid.p, eid_sorted.p, (int64_t)E, 0, bits, ctx->stream));
    CU(tmp.alloc(tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, ctx->col2.p, dst_sorted.p, eid.p, eid_sorted.p, (int64_t)E, 0, bits, ctx->stream));
    k_in_ptr<<<nblocks(E + 1, 256), 256, 0, ctx->stream>>>(E, n, dst_sorted.p, ctx->in_ptr_all.p);
    ctx->st.kernel_launches += 3;
    uint64_t qr[2];
    CU(cudaMemcpyAsync(&qr[0], ctx->in_ptr_all.p + ctx->lo, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&qr[1], ctx->in_ptr_all.p + ctx->hi, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    ctx->in_base = qr[0];
    const uint64_t cnt = qr[1] - qr[0];
    ctx->in_cnt = cnt;
    CU(ctx->in_rec.alloc(std::max<uint64_t>(cnt, 1)));
    CU(ctx->in_src.alloc(std::max<uint64_t>(cnt, 1)));
    CU(ctx->in_eid.alloc(std::max<uint64_t>(cnt, 1)));
    if (cnt) {
        k_in_struct<<<nblocks(cnt, 256), 256, 0, ctx->stream>>>(qr[0], qr[1], n, eid_sorted.p, dst_sorted.p, ctx->lo, ctx->row_ptr2.p,
                                                                ctx->in_src.p, ctx->in_eid.p);
        ctx->st.kernel_launches++;
    }
    // rows of the tiled kernels (rows of at most 16 neighbours): padded length, slot map, byte map of firing counts
    ctx->rowpack.release(); ctx->erank.release(); ctx->fired.release();
    ctx->ext_ptr.release(); ctx->ext_slot.release(); ctx->ext_cnt = 0; ctx->cell_map_bytes = 0;
    if (ctx->KP) {
        // + 32 nodes of padding: the cell kernel copies whole tiles (TMA bulk copies) even for the last, partial one
        CU(ctx->rowpack.alloc((n + 32) * (uint64_t)ctx->KP));
        CU(ctx->erank.alloc((n + 32) * (uint64_t)ctx->KP));
        CU(cudaMemsetAsync(ctx->rowpack.p, 0xff, (n + 32) * (uint64_t)ctx->KP * sizeof(uint2), ctx->stream));
        CU(cudaMemsetAsync(ctx->erank.p, 0xff, (n + 32) * (uint64_t)ctx->KP * sizeof(uint32_t), ctx->stream));
        if (cnt) {
            k_erank<<<nblocks(cnt, 256), 256, 0, ctx->stream>>>(cnt, qr[0], ctx->KP, ctx->in_src.p, ctx->in_eid.p, ctx->row_ptr2.p, ctx->erank.p);
            ctx->st.kernel_launches++;
        }
        if (ctx->nranks == 1) {
            CU(ctx->fired.alloc(E + 64));
            CU(cudaMemsetAsync(ctx->fired.p, 0, E + 64, ctx->stream));
        }
        // cell kernel: in-edges whose source lies outside the destination's cell, grouped by cell; largest byte map
        CU(ctx->ext_ptr.alloc((uint64_t)ctx->n_cells + 1));
        CU(ctx->ext_slot.alloc(std::max<uint64_t>(cnt, 1)));
        if (cnt) {
            DevBuf<unsigned char> flag, tmp2; DevBuf<unsigned long long> nsel;
            CU(flag.alloc(cnt)); CU(nsel.alloc(1));
            k_ext_flags<<<nblocks(cnt, 256), 256, 0, ctx->stream>>>(cnt, qr[0], dst_sorted.p, ctx->in_src.p, ctx->n_cells, ctx->cell_start.p, flag.p);
            size_t tb = 0;
            thrust::counting_iterator<uint32_t> pos(0);
            CU(cub::DeviceSelect::Flagged(nullptr, tb, pos, flag.p, ctx->ext_slot.p, nsel.p, (int64_t)cnt, ctx->stream));
            CU(tmp2.alloc(tb));
            CU(cub::DeviceSelect::Flagged(tmp2.p, tb, pos, flag.p, ctx->ext_slot.p, nsel.p, (int64_t)cnt, ctx->stream));
            unsigned long long nx = 0;
            CU(cudaMemcpyAsync(&nx, nsel.p, sizeof(nx), cudaMemcpyDeviceToHost, ctx->stream));
            if ((rc = sync_stream(ctx))) return rc;
            ctx->ext_cnt = nx;
            ctx->st.kernel_launches += 2;
        }
        k_ext_ptr<<<nblocks((uint64_t)ctx->n_cells + 1, 256), 256, 0, ctx->stream>>>(ctx->n_cells, ctx->cell_start.p, ctx->in_ptr_all.p, qr[0], cnt,
                                                                                   ctx->ext_cnt, ctx->ext_slot.p, ctx->ext_ptr.p);
        // whole-graph figures (identical on every rank: they decide the kernel and the launch length)
        unsigned long long zero2[2] = {0ull, 0ull}, res2[2] = {0ull, 0ull};
        CU(cudaMemcpyAsync(ctx->errword.p, zero2, sizeof(zero2), cudaMemcpyHostToDevice, ctx->stream));
        k_cell_in_max<<<nblocks(ctx->n_cells, 256), 256, 0, ctx->stream>>>(0, ctx->n_cells, ctx->cell_start.p, ctx->in_ptr_all.p,
                                                                        (unsigned int *)ctx->errword.p);
        k_count_cross_cell<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->row_ptr2.p, ctx->col2.p, ctx->n_cells, ctx->cell_start.p, ctx->errword.p + 1);
        CU(cudaMemcpyAsync(res2, ctx->errword.p, sizeof(res2), cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_stream(ctx))) return rc;
        ctx->cell_map_bytes = ((unsigned int)res2[0] + 15u) / 16u * 16u + 16u;
        ctx->ext_cnt_global = res2[1];
        ctx->st.kernel_launches += 3;
    }
    if ((rc = sync_stream(ctx))) return rc;
    ctx->have_struct = true;
    ctx->alias_dirty = ctx->have_alias;
    return ANNEMBED_OK;
}

// Sector-level alias table of the hubness sampler (embedder.rs:909-931 NodeSampler, restated one level up): Vose's alias
// method over the SECTORS of 4 consecutive nodes of the internal numbering (weight = the sum of the 4 node weights), plus the 3
// cumulative thresholds that pick a node inside a sector.  P(node) = P(sector) * P(node | sector): exactly the node law.
// The 4 lanes of a group share the sector draw and its accept / alias decision (ONE coalesced table sector, then ONE
// coalesced row sector of the layout) and each picks its row from a rotation of one shared uniform
// (tests/studies/sector_alias_study.py: chi-square per lane; tests/test_gpu_parity.py on the device).
// Entry of sector s (32 bytes, one memory sector, so that the draw is ONE gather and branch-free):
// uint4 {bits(prob), alias sector, bits(thr0), bits(thr1)}, uint4 {bits(thr2), bits(thr0), bits(thr1), bits(thr2) of the ALIAS sector}.
static int build_sector_alias(annembed_cuda_ctx *ctx)
{
    const uint64_t n = ctx->n, nsec = (n + 3) / 4;
    if (ctx->neg_w_host.size() != n) { ctx->sec_alias.release(); return ANNEMBED_OK; }
    std::vector<uint32_t> old_of_new(n);
    CU(cudaMemcpyAsync(old_of_new.data(), ctx->old_of_new.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    std::vector<uint4> tab;
    auto weight = [&](uint64_t i) { return (double)ctx->neg_w_host[old_of_new[i]]; };
    // Line-level tables of the event kernels (layouts of dimension <= 4): the G = 16 (dimension 2) or 8 (dimension 3-4)
    // nodes whose rows fill a 128-byte line of the layout share the line draw.  Two alias methods, one inside the other:
    // T1 over the LINES (weight = sum of the line's node weights): {prob, alias line};  T2 inside every line over its G
    // rows (conditional law w_i / W_line): per row the accept threshold in units of 2^-24 and the alias row.  A lane reads
    // T1[line] (8 bytes, the same for the whole group), then ITS column of T2 of the final line (the group's columns are
    // distinct: one coalesced read), then its row of that line of the layout (one line request per group).  The T2 entry
    // of a line carries the inner table of its ALIAS line behind its own (2 G words = one 128-byte line in dimension 2), so
    // that T1 and both candidate columns are read at once: the draw is one round of table reads, then the row.
    // P(node) = P(line) * P(row | line): exactly the node law (embedder.rs:909-931 restated two levels up).
    // Built on a second host thread while this one builds and uploads the sector table (both are host work, 0.2-0.3 s each
    // at 11M nodes).
    const bool want_lines = ctx->DP <= 4;
    const uint32_t G = ctx->DP == 2 ? 16u : 8u;
    const uint64_t nl = (n + G - 1) / G;
    std::vector<uint2> t1;
    std::vector<uint32_t> t2;
    struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } lines;   // joined on every return path
    if (want_lines) lines.t = std::thread([&]() { annembed_host::build_line_alias_tables(n, G, weight, t1, t2); });
    annembed_host::build_sector_alias_table(n, weight, tab);
    CU(ctx->sec_alias.alloc(2 * nsec));
    if ((rc = h2d(ctx, ctx->sec_alias.p, tab.data(), 2 * nsec * sizeof(uint4)))) return rc;
    if (!want_lines) { ctx->line_t1.release(); ctx->line_t2.release(); return ANNEMBED_OK; }
    lines.t.join();
    CU(ctx->line_t1.alloc(nl)); CU(ctx->line_t2.alloc(nl * 2 * G));
    if ((rc = h2d(ctx, ctx->line_t1.p, t1.data(), nl * sizeof(uint2)))) return rc;
    return h2d(ctx, ctx->line_t2.p, t2.data(), nl * 2 * G * sizeof(uint32_t));
}

// device context build, weights part: K2 + cumulative row probabilities + rows + in-edge payloads
// (≙ EntropyOptim::new, embedder.rs:964-1025)
static int ensure_build(annembed_cuda_ctx *ctx)
{
    int rc;
    if (ctx->have_build) {
        if (ctx->alias_dirty && ctx->have_alias) goto alias;
        return ANNEMBED_OK;
    }
    REQUIRE(ctx->have_weights, ANNEMBED_ERR_STATE, "edge weights not computed (embedder.rs:802-808: initial_space not constructed)");
    {
        const uint64_t n = ctx->n, E = ctx->E;
        CU(cudaEventRecord(ctx->ev_a, ctx->stream));
        if ((rc = ensure_struct(ctx))) return rc;
        if (ctx->emb_scale.n != n) { CU(ctx->emb_scale.alloc(n)); CU(ctx->inv_s2.alloc(n)); CU(ctx->inv_s2n.alloc(n + 32)); CU(cudaMemsetAsync(ctx->inv_s2n.p, 0, (n + 32) * sizeof(float), ctx->stream)); }
        if (ctx->cum.n != E) CU(ctx->cum.alloc(E));
        k_row_cumsum<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p, ctx->row_ptr.p, ctx->proba.p, ctx->row_ptr2.p, ctx->cum.p);
        ctx->st.kernel_launches++;
        // K2
        if ((rc = sum_f64(ctx, ctx->scale.p, n, ctx->partials.p + 4095))) return rc;
        k_embedded_scales<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->scale.p, ctx->partials.p + 4095, ctx->emb_scale.p, ctx->inv_s2.p);
        k_gather_f32<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->old_of_new.p, ctx->inv_s2.p, ctx->inv_s2n.p);
        ctx->st.kernel_launches += 2;
        if (ctx->KP) {
            // node-major rows for the bulk-synchronous kernels (the cell kernel copies whole tiles), tile-interleaved 16-byte
            // chunks for the asynchronous form (async_sweep.cuh async_row_ptr)
            k_rowpack<<<nblocks(n * (uint64_t)ctx->KP, 256), 256, 0, ctx->stream>>>(n, ctx->KP, ctx->struct_async ? 1 : 0, ctx->row_ptr2.p, ctx->col2.p,
                                                                                    ctx->cum.p, ctx->rowpack.p);
            ctx->st.kernel_launches++;
        }
        if (ctx->in_cnt) {
            k_in_rec<<<nblocks(ctx->in_cnt, 256), 256, 0, ctx->stream>>>(ctx->in_cnt, ctx->in_src.p, ctx->in_eid.p, ctx->row_ptr2.p, ctx->cum.p,
                                                                         ctx->inv_s2n.p, ctx->in_rec.p);
            ctx->st.kernel_launches++;
        }
        CU(cudaEventRecord(ctx->ev_b, ctx->stream));
        if ((rc = sync_stream(ctx))) return rc;
        float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        ctx->st.build_ms = ms;
        ctx->have_build = true;
    }
alias:
    if (ctx->alias_dirty && ctx->have_alias) {
        CU(ctx->neg_alias.alloc(ctx->n));
        k_relabel_alias<<<nblocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->new_of_old.p, ctx->neg_alias_old.p, ctx->neg_alias.p);
        ctx->st.kernel_launches++;
        if ((rc = sync_stream(ctx))) return rc;
        if ((rc = build_sector_alias(ctx))) return rc;
        ctx->alias_dirty = false;
    }
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_get_hubness_counts(annembed_cuda_ctx *ctx, uint32_t *counts)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(counts, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "get_hubness_counts: graph not set");
    CU(cudaSetDevice(ctx->device));
    // in-degree of every node (fromhnsw/hubness.rs:39-76): a histogram of the neighbour lists, caller's numbering
    DevBuf<uint32_t> t; CU(t.alloc(ctx->n));
    CU(cudaMemsetAsync(t.p, 0, ctx->n * sizeof(uint32_t), ctx->stream));
    k_in_degree<<<nblocks(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->col.p, t.p);
    ctx->st.kernel_launches += 1;
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    return d2h(ctx, counts, t.p, ctx->n * sizeof(uint32_t));
}

extern "C" int annembed_cuda_set_embedding(annembed_cuda_ctx *ctx, const float *y)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "set_embedding: graph not set");
    REQUIRE(y, ANNEMBED_ERR_INVALID_ARG, "null embedding");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n;
    const int d = (int)ctx->prm.asked_dim, DP = ctx->DP;
    int rc;
    if ((rc = alloc_layout(ctx))) return rc;
    if (d == DP) {
        if ((rc = h2d(ctx, ctx->y0.p, y, n * d * sizeof(float)))) return rc;
    } else {
        DevBuf<float> stage; CU(stage.alloc(n * d));
        if ((rc = h2d(ctx, stage.p, y, n * d * sizeof(float)))) return rc;
        k_pad_rows<<<nblocks(n * DP, 256), 256, 0, ctx->stream>>>(n, d, DP, stage.p, ctx->y0.p);
        ctx->st.kernel_launches++;
        if ((rc = sync_stream(ctx))) return rc;
    }
    ctx->have_embedding = true;
    return annembed_cuda_reset_embedding(ctx);
}

extern "C" int annembed_cuda_set_embedding_from_projection(annembed_cuda_ctx *ctx, uint64_t n_small, const float *first,
                                                           const uint32_t *proj_node, const float *proj_dist, float median_dist)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "set_embedding_from_projection: graph not set");
    REQUIRE(first && proj_node && proj_dist, ANNEMBED_ERR_INVALID_ARG, "null argument");
    REQUIRE(n_small >= 1 && n_small <= ctx->n, ANNEMBED_ERR_INVALID_ARG, "n_small must be in 1..n");
    REQUIRE(median_dist > 0.0f && std::isfinite(median_dist), ANNEMBED_ERR_INVALID_ARG, "median projection distance must be > 0");
    for (uint64_t i = n_small; i < ctx->n; i++)
        REQUIRE(proj_node[i] < n_small, ANNEMBED_ERR_INVALID_ARG, "projection target is not a node of the small graph");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n;
    const int d = (int)ctx->prm.asked_dim, DP = ctx->DP;
    int rc;
    if ((rc = alloc_layout(ctx))) return rc;
    DevBuf<float> dfirst, ddist; DevBuf<uint32_t> dnode;
    CU(dfirst.alloc(n_small * d)); CU(ddist.alloc(n)); CU(dnode.alloc(n));
    if ((rc = h2d(ctx, dfirst.p, first, n_small * d * sizeof(float)))) return rc;
    if ((rc = h2d(ctx, dnode.p, proj_node, n * sizeof(uint32_t)))) return rc;
    if ((rc = h2d(ctx, ddist.p, proj_dist, n * sizeof(float)))) return rc;
    k_project_init<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, n_small, d, DP, dfirst.p, dnode.p, ddist.p, median_dist,
                                                             (uint32_t)(ctx->prm.seed & 0xFFFFFFFFu), (uint32_t)(ctx->prm.seed >> 32), ctx->y0.p);
    ctx->st.kernel_launches++;
    if ((rc = sync_stream(ctx))) return rc;
    ctx->have_embedding = true;
    return annembed_cuda_reset_embedding(ctx);
}

// =====================================================================================================
// N2: diffusion-map initial layout (embedder.rs:308-345) -- kernels in dmap.cuh
// =====================================================================================================
namespace {
// small dense fp64 helpers for the DMAP_RANK x DMAP_RANK projected problems (host side)
bool cholesky_upper(const double *G, int r, double *R)          // G = R^T R, R upper triangular (row-major)
{
    for (int i = 0; i < r * r; i++) R[i] = 0.0;
    for (int j = 0; j < r; j++) {
        double s = G[j * r + j];
        for (int k = 0; k < j; k++) s -= R[k * r + j] * R[k * r + j];
        if (!(s > 0.0)) return false;
        const double d = std::sqrt(s);
        R[j * r + j] = d;
        for (int i = j + 1; i < r; i++) {
            double t = G[j * r + i];
            for (int k = 0; k < j; k++) t -= R[k * r + j] * R[k * r + i];
            R[j * r + i] = t / d;
        }
    }
    return true;
}
void upper_inverse(const double *R, int r, double *Ri)           // Ri = R^-1 (upper triangular)
{
    for (int i = 0; i < r * r; i++) Ri[i] = 0.0;
    for (int j = 0; j < r; j++) {
        Ri[j * r + j] = 1.0 / R[j * r + j];
        for (int i = j - 1; i >= 0; i--) {
            double s = 0.0;
            for (int k = i + 1; k <= j; k++) s += R[i * r + k] * Ri[k * r + j];
            Ri[i * r + j] = -s / R[i * r + i];
        }
    }
}
// cyclic Jacobi for a symmetric matrix: A = V diag(w) V^T, eigenvalues sorted in decreasing order, V column j <-> w[j]
void jacobi_eigh(const double *A_in, int r, double *w, double *V)
{
    std::vector<double> A(A_in, A_in + r * r);
    for (int i = 0; i < r; i++) for (int j = 0; j < r; j++) V[i * r + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < r; i++) for (int j = 0; j < r; j++) (i == j ? dia : off) += A[i * r + j] * A[i * r + j];
        if (off <= 1e-30 * dia) break;
        for (int p = 0; p < r - 1; p++)
            for (int q = p + 1; q < r; q++) {
                const double apq = A[p * r + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * r + q] - A[p * r + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < r; k++) {
                    const double akp = A[k * r + p], akq = A[k * r + q];
                    A[k * r + p] = c * akp - s * akq; A[k * r + q] = s * akp + c * akq;
                }
                for (int k = 0; k < r; k++) {
                    const double apk = A[p * r + k], aqk = A[q * r + k];
                    A[p * r + k] = c * apk - s * aqk; A[q * r + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < r; k++) {
                    const double vkp = V[k * r + p], vkq = V[k * r + q];
                    V[k * r + p] = c * vkp - s * vkq; V[k * r + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(r);
    for (int i = 0; i < r; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return A[a * r + a] > A[b * r + b]; });
    std::vector<double> Vs(r * r);
    for (int j = 0; j < r; j++) {
        w[j] = A[order[j] * r + order[j]];
        for (int k = 0; k < r; k++) Vs[k * r + j] = V[k * r + order[j]];
    }
    std::copy(Vs.begin(), Vs.end(), V);
}
} // namespace

static int host_sum(annembed_cuda_ctx *ctx, const float *x, uint64_t n, double *out)
{
    int rc;
    if ((rc = sum_f64(ctx, x, n, ctx->partials.p + 4095))) return rc;
    CU(cudaMemcpyAsync(out, ctx->partials.p + 4095, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_stream(ctx);
}

// the symmetric normalised kernel of the loaded graph:  diag + A + A^T, A = one value per directed edge
struct DmapKernel {
    DevBuf<uint64_t> tp_ptr;                 // transposed index of the whole graph
    DevBuf<uint32_t> tp_src, tp_eid;
    DevBuf<float> normed, val, diag, sw;     // normed first-pass scales, A values, diagonal, sqrt(degrees)
    double sw_sum = 0.0;
};

static int dmap_build_kernel(annembed_cuda_ctx *ctx, uint32_t gnbn, DmapKernel &K)
{
    const uint64_t n = ctx->n, E = ctx->E;
    const float alfa = 0.5f, beta = -0.1f, epsil = 2.0f;            // embedder.rs:319-320, diffmaps.rs:100
    int rc;
    cudaStream_t st = ctx->stream;
    const unsigned int gn = nblocks(n, 256);
    DevBuf<uint64_t> &tp_ptr = K.tp_ptr;
    DevBuf<uint32_t> &tp_src = K.tp_src, &tp_eid = K.tp_eid;
    DevBuf<float> &normed = K.normed, &val = K.val, &diag = K.diag, &sw = K.sw;
    // ---- transposed index of the whole graph (every rank computes the whole layout: it is replicated anyway)
    {
        DevBuf<uint32_t> eid, dst_sorted, eid_sorted;
        DevBuf<unsigned char> tmp;
        CU(tp_ptr.alloc(n + 2)); CU(tp_src.alloc(E)); CU(tp_eid.alloc(E));
        CU(eid.alloc(E)); CU(dst_sorted.alloc(E)); CU(eid_sorted.alloc(E));
        k_iota<<<nblocks(E, 256), 256, 0, st>>>(E, eid.p);
        int bits = 1; while (bits < 32 && (1ull << bits) < n) bits++;
        size_t tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->col.p, dst_sorted.p, eid.p, eid_sorted.p, (int64_t)E, 0, bits, st));
        CU(tmp.alloc(tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, ctx->col.p, dst_sorted.p, eid.p, eid_sorted.p, (int64_t)E, 0, bits, st));
        k_in_ptr<<<nblocks(E + 1, 256), 256, 0, st>>>(E, n, dst_sorted.p, tp_ptr.p);
        k_in_struct<<<nblocks(E, 256), 256, 0, st>>>(0, E, n, eid_sorted.p, dst_sorted.p, 0u, ctx->row_ptr.p, tp_src.p, tp_eid.p);
        ctx->st.kernel_launches += 4;
        if ((rc = sync_stream(ctx))) return rc;
    }

    // ---- node kernels: two passes of scales (diffmaps.rs:752-849)
    DevBuf<float> scale, w_self, q;
    CU(scale.alloc(n)); CU(normed.alloc(n)); CU(w_self.alloc(n)); CU(val.alloc(E)); CU(q.alloc(n)); CU(diag.alloc(n));
    const uint32_t nbgh = std::min<uint32_t>(gnbn, ctx->kmax);
    k_dmap_local_scale<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, ctx->dist.p, nbgh, scale.p);
    double ssum = 0.0;
    if ((rc = host_sum(ctx, scale.p, n, &ssum))) return rc;
    const float mean_scale = (float)(ssum / (double)n);
    REQUIRE(mean_scale > 0.0f && std::isfinite(mean_scale), ANNEMBED_ERR_INVALID_ARG, "dmap_init: all neighbour distances are zero");
    k_dmap_fix_scale<<<gn, 256, 0, st>>>(n, mean_scale, scale.p, normed.p);
    const float sqrt_epsil = std::sqrt(epsil);
    auto kernel_pass = [&](const float *scales) -> int {             // w_self, val = symmetrised weights, q = row sums
        DevBuf<float> w;
        CU(w.alloc(E));
        k_dmap_kernel_weights<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, ctx->col.p, ctx->dist.p, scales, sqrt_epsil, w_self.p, w.p);
        k_dmap_symmetrise<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, ctx->col.p, w.p, val.p);
        k_dmap_rowsum<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, tp_ptr.p, tp_eid.p, val.p, w_self.p, 2.0f, q.p);
        ctx->st.kernel_launches += 3;
        return sync_stream(ctx);
    };
    if ((rc = kernel_pass(scale.p))) return rc;
    double qsum = 0.0;
    if ((rc = host_sum(ctx, q.p, n, &qsum))) return rc;
    {   // kernel0_to_density (diffmaps.rs:852-942): scales <- (q / mean q)^beta * mean_scale
        DevBuf<float> scale2;
        CU(scale2.alloc(n));
        k_dmap_beta_scales<<<gn, 256, 0, st>>>(n, q.p, qsum / (double)n, beta, mean_scale, scale2.p);
        if ((rc = kernel_pass(scale2.p))) return rc;
    }
    // ---- compute_laplacian, sparse branch (diffmaps.rs:504-587)
    if ((rc = host_sum(ctx, q.p, n, &qsum))) return rc;
    k_dmap_alpha<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, ctx->col.p, q.p, qsum / (double)ctx->kmax, alfa, val.p, w_self.p, diag.p);
    CU(sw.alloc(n));
    k_dmap_rowsum<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, tp_ptr.p, tp_eid.p, val.p, diag.p, 1.0f, sw.p);
    k_dmap_sqrt<<<gn, 256, 0, st>>>(n, sw.p);
    k_dmap_normalise<<<gn, 256, 0, st>>>(n, ctx->row_ptr.p, ctx->col.p, sw.p, val.p, diag.p);
    ctx->st.kernel_launches += 7;
    return host_sum(ctx, sw.p, n, &K.sw_sum);
}


extern "C" int annembed_cuda_dmap_init(annembed_cuda_ctx *ctx, uint32_t gnbn, float diffusion_time, float *y_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "dmap_init: graph not set");
    constexpr int R = DMAP_RANK;
    const uint64_t n = ctx->n;
    const int d = (int)ctx->prm.asked_dim, DP = ctx->DP;
    REQUIRE(d <= R - 1, ANNEMBED_ERR_UNSUPPORTED, "dmap_init: asked_dim must be <= 19 (rank-20 range finder, graphlaplace.rs:113)");
    REQUIRE(n >= 4 * R, ANNEMBED_ERR_UNSUPPORTED, "dmap_init: graph too small for the rank-20 range finder");
    if (gnbn == 0) gnbn = 12;                                       // embedder.rs:317
    if (!(diffusion_time > 0.0f)) diffusion_time = 5.0f;            // embedder.rs:316
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = alloc_layout(ctx))) return rc;
    cudaStream_t st = ctx->stream;
    const unsigned int gn = nblocks(n, 256);
    DmapKernel K;
    if ((rc = dmap_build_kernel(ctx, gnbn, K))) return rc;
    DevBuf<uint64_t> &tp_ptr = K.tp_ptr;
    DevBuf<uint32_t> &tp_src = K.tp_src, &tp_eid = K.tp_eid;
    DevBuf<float> &normed = K.normed, &val = K.val, &diag = K.diag, &sw = K.sw;
    const double sw_sum = K.sw_sum;

    // ---- range finder: rank-20 subspace iteration, 5 iterations (svdapprox.rs:343-410), QR by two Cholesky passes
    DevBuf<float> Ya, Yb;
    DevBuf<double> gpart, dM;
    CU(Ya.alloc(n * R)); CU(Yb.alloc(n * R)); CU(dM.alloc(R * R + R * R));
    const unsigned int gram_blocks = std::min<unsigned int>(nblocks(n, DMAP_GRAM_ROWS), (unsigned int)ctx->sm_count * 4u);
    CU(gpart.alloc((size_t)gram_blocks * R * R));
    std::vector<double> G(R * R), Rm(R * R), Ri(R * R);
    auto gram = [&](const float *Y) -> int {
        k_dmap_gram<<<gram_blocks, R * R, 0, st>>>(n, Y, gpart.p);
        k_dmap_gram_final<<<1, R * R, 0, st>>>(gram_blocks, gpart.p, dM.p + R * R);
        ctx->st.kernel_launches += 2;
        CU(cudaMemcpyAsync(G.data(), dM.p + R * R, sizeof(double) * R * R, cudaMemcpyDeviceToHost, st));
        return sync_stream(ctx);
    };
    auto right_multiply = [&](const double *M, int ncols, const float *Yin, float *out, int stride) -> int {
        CU(cudaMemcpyAsync(dM.p, M, sizeof(double) * R * ncols, cudaMemcpyHostToDevice, st));
        k_dmap_right_multiply<<<nblocks(n, 128), 128, 0, st>>>(n, ncols, dM.p, Yin, out, stride);
        ctx->st.kernel_launches++;
        return sync_stream(ctx);
    };
    auto orthonormalise = [&](float *Y) -> int {                     // Y <- Q of its QR factorisation (in place)
        for (int pass = 0; pass < 2; pass++) {
            if ((rc = gram(Y))) return rc;
            REQUIRE(cholesky_upper(G.data(), R, Rm.data()), ANNEMBED_ERR_UNSUPPORTED, "dmap_init: range finder lost rank (graph kernel has fewer than 20 significant directions)");
            upper_inverse(Rm.data(), R, Ri.data());
            if ((rc = right_multiply(Ri.data(), R, Y, Y, R))) return rc;
        }
        return ANNEMBED_OK;
    };
    auto spmm = [&](const float *X, float *Y) {
        k_dmap_spmm<<<nblocks(n, 128), 128, 0, st>>>(n, ctx->row_ptr.p, ctx->col.p, tp_ptr.p, tp_src.p, tp_eid.p, val.p, diag.p, X, Y);
        ctx->st.kernel_launches++;
    };
    if (ctx->dmap_omega.size() == (size_t)n * R) {
        if ((rc = h2d(ctx, Yb.p, ctx->dmap_omega.data(), (size_t)n * R * sizeof(float)))) return rc;
    } else {
        k_dmap_gaussian<<<gn, 256, 0, st>>>(n, (uint32_t)(ctx->prm.seed & 0xFFFFFFFFu), (uint32_t)(ctx->prm.seed >> 32), Yb.p);
    }
    spmm(Yb.p, Ya.p);
    if ((rc = orthonormalise(Ya.p))) return rc;
    for (int it = 1; it < DMAP_ITERS; it++) {
        spmm(Ya.p, Yb.p);                                           // K^T Q (the kernel is symmetric)
        if ((rc = orthonormalise(Yb.p))) return rc;
        spmm(Yb.p, Ya.p);
        if ((rc = orthonormalise(Ya.p))) return rc;
    }
    // ---- direct SVD (svdapprox.rs:721-801): B = Q^T K; B B^T = (K Q)^T (K Q); U = Q U_B, sigma = sqrt(eig)
    spmm(Ya.p, Yb.p);
    if ((rc = gram(Yb.p))) return rc;
    std::vector<double> ev(R), V(R * R);
    jacobi_eigh(G.data(), R, ev.data(), V.data());
    REQUIRE(ev[0] > 0.0, ANNEMBED_ERR_CUDA, "dmap_init: projected kernel has no positive eigenvalue");
    std::vector<double> Vd((size_t)R * d);
    std::vector<float> lam_t(32, 0.0f);
    for (int c = 0; c < d; c++) {
        for (int k = 0; k < R; k++) Vd[(size_t)k * d + c] = V[k * R + (c + 1)];      // skip the first (stationary) vector
        const double ratio = std::sqrt(std::max(ev[c + 1], 0.0) / ev[0]);           // lambda_{c+1} / lambda_0 (diffmaps.rs:1208)
        lam_t[c] = (float)std::pow(ratio, (double)diffusion_time);
    }
    ctx->dmap_sigma.assign(R, 0.0);
    for (int k = 0; k < R; k++) ctx->dmap_sigma[k] = std::sqrt(std::max(ev[k], 0.0));
    DevBuf<float> U, dlam;
    CU(U.alloc(n * d)); CU(dlam.alloc(32));
    if ((rc = right_multiply(Vd.data(), d, Ya.p, U.p, d))) return rc;
    CU(cudaMemcpyAsync(dlam.p, lam_t.data(), 32 * sizeof(float), cudaMemcpyHostToDevice, st));
    // ---- coordinates (diffmaps.rs:1219-1236) into the padded layout, then set_data_box(., 10) (embedder.rs:345)
    CU(cudaMemsetAsync(ctx->y0.p, 0, ctx->y0.n * sizeof(float), st));
    k_dmap_coordinates<<<gn, 256, 0, st>>>(n, d, DP, U.p, dlam.p, normed.p, sw.p, (float)(sw_sum / (double)n), ctx->y0.p);
    const unsigned int rb = std::min<unsigned int>(gn, 1024u);
    DevBuf<double> cpart;
    DevBuf<float> bmax, dmeans;
    CU(cpart.alloc((size_t)rb * 32)); CU(bmax.alloc(rb)); CU(dmeans.alloc(32));
    k_dmap_colsum<<<rb, 256, 0, st>>>(n, d, DP, ctx->y0.p, cpart.p);
    std::vector<double> hpart((size_t)rb * 32);
    CU(cudaMemcpyAsync(hpart.data(), cpart.p, hpart.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    if ((rc = sync_stream(ctx))) return rc;
    std::vector<float> means(32, 0.0f);
    for (int c = 0; c < d; c++) {
        double s = 0.0;
        for (unsigned int b = 0; b < rb; b++) s += hpart[(size_t)b * 32 + c];
        means[c] = (float)(s / (double)n);
    }
    CU(cudaMemcpyAsync(dmeans.p, means.data(), 32 * sizeof(float), cudaMemcpyHostToDevice, st));
    k_dmap_center_max<<<rb, 256, 0, st>>>(n, d, DP, dmeans.p, ctx->y0.p, bmax.p);
    std::vector<float> hmax(rb);
    CU(cudaMemcpyAsync(hmax.data(), bmax.p, rb * sizeof(float), cudaMemcpyDeviceToHost, st));
    if ((rc = sync_stream(ctx))) return rc;
    float mx = 0.0f;
    for (float v : hmax) mx = std::max(mx, v);
    REQUIRE(mx > 0.0f && std::isfinite(mx), ANNEMBED_ERR_CUDA, "dmap_init: degenerate layout");
    k_dmap_scale<<<nblocks(n * DP, 256), 256, 0, st>>>(n * DP, 1.0f / (mx / 5.0f), ctx->y0.p);       // box_size 10 / 2
    ctx->st.kernel_launches += 6;
    if ((rc = sync_stream(ctx))) return rc;
    ctx->have_embedding = true;
    if ((rc = annembed_cuda_reset_embedding(ctx))) return rc;
    if (y_out) return annembed_cuda_get_embedding(ctx, y_out);
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_dmap_singular_values(const annembed_cuda_ctx *ctx, double *sigma_out, uint32_t count)
{
    if (!ctx || !sigma_out) return ANNEMBED_ERR_INVALID_ARG;
    if (ctx->dmap_sigma.empty()) return ANNEMBED_ERR_STATE;
    for (uint32_t k = 0; k < count; k++) sigma_out[k] = k < ctx->dmap_sigma.size() ? ctx->dmap_sigma[k] : 0.0;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_dmap_set_test_matrix(annembed_cuda_ctx *ctx, const float *omega, uint64_t rows)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    if (!omega) { ctx->dmap_omega.clear(); return ANNEMBED_OK; }
    REQUIRE(ctx->have_graph && rows == ctx->n, ANNEMBED_ERR_INVALID_ARG, "dmap_set_test_matrix: needs the graph and n rows of 20 columns");
    ctx->dmap_omega.assign(omega, omega + (size_t)rows * DMAP_RANK);
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_dmap_kernel(annembed_cuda_ctx *ctx, uint32_t gnbn, float *diag_out, float *val_out,
                                        float *sw_out, float *normed_scale_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "dmap_kernel: graph not set");
    if (gnbn == 0) gnbn = 12;
    CU(cudaSetDevice(ctx->device));
    DmapKernel K;
    int rc;
    if ((rc = dmap_build_kernel(ctx, gnbn, K))) return rc;
    if (diag_out && (rc = d2h(ctx, diag_out, K.diag.p, ctx->n * sizeof(float)))) return rc;
    if (val_out && (rc = d2h(ctx, val_out, K.val.p, ctx->E * sizeof(float)))) return rc;
    if (sw_out && (rc = d2h(ctx, sw_out, K.sw.p, ctx->n * sizeof(float)))) return rc;
    if (normed_scale_out && (rc = d2h(ctx, normed_scale_out, K.normed.p, ctx->n * sizeof(float)))) return rc;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_reset_embedding(annembed_cuda_ctx *ctx)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_embedding, ANNEMBED_ERR_STATE, "reset_embedding: embedding not set");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->yapi.p, ctx->y0.p, ctx->y0.n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return sync_stream(ctx);
}

extern "C" int annembed_cuda_get_embedded_scales(annembed_cuda_ctx *ctx, float *out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(out, ANNEMBED_ERR_INVALID_ARG, "null output");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_build(ctx))) return rc;
    return d2h(ctx, out, ctx->emb_scale.p, ctx->n * sizeof(float));
}

// ---- schedule -------------------------------------------------------------------------------------------------------
// Asynchronous form (one rank, default): a batch is `M` sweeps over the nodes, every node fires kappa = nb_sampling_by_edge
// * mean degree / M times per sweep.  Default kappa = ANNEMBED_ASYNC_KAPPA = 1: one sample per visit of a node, as spread
// out in time as the reference's independent edge draws.  Measured on the BASELINE.json configs against the serial
// reference loop (tools/gpu_fidelity_probe.py): at 1 firing per visit every layout statistic is within
// 0.5 % of the reference's, at 2 (the node's samples come in bursts of two) kNN preservation is 6 % off, at 4, 15 %.
#ifndef ANNEMBED_ASYNC_KAPPA
#define ANNEMBED_ASYNC_KAPPA 0.25
#endif
#ifndef ANNEMBED_ASYNC_THIN
#define ANNEMBED_ASYNC_THIN 4u           // thinned sub-sweeps per launch (kappa = 1 per launch in total)
#endif
// several ranks: needs the peers' replicas (annembed_cuda_comm_import_layouts); without them the bulk-synchronous form
// with the NCCL all-gather runs
static bool use_async(const annembed_cuda_ctx *ctx)
{
    return (ctx->nranks == 1 || ctx->have_peers) &&
           !(ctx->prm.flags & (ANNEMBED_FLAG_BULK_SYNCHRONOUS | ANNEMBED_FLAG_REPLAY_IN_EDGES | ANNEMBED_FLAG_LEGACY_EPOCH_KERNELS));
}
// Bulk-synchronous form (ANNEMBED_FLAG_BULK_SYNCHRONOUS, and every multi-rank run): mini-epochs per reference batch.
// What governs the fidelity of the bulk-synchronous loop is how much of a batch is applied against one snapshot, i.e.
// samples per edge per mini-epoch (nb_sampling_by_edge / M).  Measured against the serial reference loop on the
// BASELINE.json configs (tools/gpu_schedule_probe.py):
//  * the neighbourhood statistics of the final layout (get_quality_estimate_from_edge_length, kNN preservation) are set by
//    the LAST batches: with the last third at 0.075 samples per edge per mini-epoch they are within 1-2 % of the
//    reference's whatever the first two thirds use (0.6 .. 0.15 samples per edge gave the same statistics);
//  * the first two thirds only move the final cross entropy (coarser early mini-epochs end 10-17 % below the reference's).
// The first two thirds run at kappa = 1 firing per node and mini-epoch (the coarsest level the event form of the cell
// kernel serves, 0.17 samples per edge at k = 6), the last third at 0.075 samples per edge (and kappa <= 1).
#ifndef ANNEMBED_SAMPLES_PER_EDGE_LATE
#define ANNEMBED_SAMPLES_PER_EDGE_LATE 0.075
#endif
static uint32_t mini_epochs_kappa(const annembed_cuda_ctx *ctx, double kappa)
{
    const double deg = ctx->n ? (double)ctx->E / (double)ctx->n : 1.0;
    return (uint32_t)std::max<double>(1.0, std::ceil((double)ctx->prm.nb_sampling_by_edge * deg / kappa - 1e-9));
}
static uint32_t eff_mini_epochs(const annembed_cuda_ctx *ctx)           // the finest level of the schedule
{
    if (ctx->prm.mini_epochs_per_batch) return ctx->prm.mini_epochs_per_batch;
    if (use_async(ctx)) return mini_epochs_kappa(ctx, ANNEMBED_ASYNC_KAPPA);
    const uint32_t late = (uint32_t)std::max<double>(1.0, std::ceil((double)ctx->prm.nb_sampling_by_edge / ANNEMBED_SAMPLES_PER_EDGE_LATE));
    return std::max(late, mini_epochs_kappa(ctx, 1.0));
}
static uint32_t mini_epochs_of_batch(const annembed_cuda_ctx *ctx, uint32_t iter)
{
    if (ctx->prm.mini_epochs_per_batch) return ctx->prm.mini_epochs_per_batch;      // explicit: every batch
    if (use_async(ctx)) return eff_mini_epochs(ctx);
    const uint32_t nb = ctx->prm.nb_grad_batch;
    return (3 * iter > 2 * nb) ? eff_mini_epochs(ctx) : mini_epochs_kappa(ctx, 1.0);
}
// global index of the first mini-epoch of batch `iter` (counter word of the Philox streams)
static uint32_t first_epoch_of_batch(const annembed_cuda_ctx *ctx, uint32_t iter)
{
    uint32_t e = 0;
    for (uint32_t i = 1; i < iter; i++) e += mini_epochs_of_batch(ctx, i);
    return e;
}

#ifndef ANNEMBED_CELL_SUBSTEPS_LOCAL
#define ANNEMBED_CELL_SUBSTEPS_LOCAL 16   // sub-steps per launch when (nearly) every edge has both ends in one cell
#endif
#ifndef ANNEMBED_FUSED_CHUNKS
#define ANNEMBED_FUSED_CHUNKS 1   // >1 staggers sub-ranges on two streams (measured: no gain at 8 GPUs, profiles/r01_bench_8gpu_*)
#endif

static SgdConst make_const(const annembed_cuda_ctx *ctx, double grad_step)
{
    SgdConst K;
    K.gamma = (float)grad_step;
    K.b = (float)ctx->prm.b;
    K.two_b = (float)(2.0 * ctx->prm.b);
    K.b_is_one = ctx->prm.b == 1.0;
    return K;
}

extern "C" int annembed_cuda_step_fixed(annembed_cuda_ctx *ctx, uint64_t n_samples, const uint64_t *edge_idx,
                                        const uint32_t *neg_idx, double grad_step)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_embedding, ANNEMBED_ERR_STATE, "step_fixed: embedding not set");
    REQUIRE(ctx->nranks == 1, ANNEMBED_ERR_UNSUPPORTED, "step_fixed is single-GPU (test mode)");
    if (n_samples == 0) return ANNEMBED_OK;
    REQUIRE(edge_idx && neg_idx, ANNEMBED_ERR_INVALID_ARG, "null sample list");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_build(ctx))) return rc;
    for (uint64_t s = 0; s < n_samples; s++) {
        REQUIRE(edge_idx[s] < ctx->E, ANNEMBED_ERR_INVALID_ARG, "edge index out of range");
        for (int q = 0; q < 5; q++) REQUIRE(neg_idx[5 * s + q] < ctx->n, ANNEMBED_ERR_INVALID_ARG, "negative index out of range");
    }
    DevBuf<uint64_t> de; DevBuf<uint32_t> dn;
    CU(de.alloc(n_samples)); CU(dn.alloc(n_samples * 5));
    if ((rc = h2d(ctx, de.p, edge_idx, n_samples * sizeof(uint64_t)))) return rc;
    if ((rc = h2d(ctx, dn.p, neg_idx, n_samples * 5 * sizeof(uint32_t)))) return rc;
    const SgdConst K = make_const(ctx, grad_step);
    float *Y = ctx->yapi.p;
#define LAUNCH_FIXED(DPV) k_step_fixed<DPV><<<1, 32, 0, ctx->stream>>>(Y, ctx->n, ctx->row_ptr.p, ctx->col.p, ctx->proba.p, ctx->inv_s2.p, K, n_samples, de.p, dn.p)
    switch (ctx->DP) {
    case 2: LAUNCH_FIXED(2); break;
    case 4: LAUNCH_FIXED(4); break;
    case 8: LAUNCH_FIXED(8); break;
    case 16: LAUNCH_FIXED(16); break;
    default: LAUNCH_FIXED(32); break;
    }
#undef LAUNCH_FIXED
    ctx->st.kernel_launches++;
    return sync_stream(ctx);
}

static EpochArgs make_epoch_args(annembed_cuda_ctx *ctx, uint32_t epoch, double grad_step, uint32_t M = 0)
{
    if (M == 0) M = eff_mini_epochs(ctx);
    EpochArgs a;
    memset(&a, 0, sizeof a);
    a.y_snap = ctx->y[ctx->cur].p;
    a.y_next = ctx->y[ctx->cur ^ 1].p;
    // everything below is in the internal numbering
    a.row_ptr = ctx->row_ptr2.p; a.col = ctx->col2.p; a.p = nullptr; a.inv_s2 = ctx->inv_s2n.p;
    a.in_ptr = ctx->in_ptr_all.p ? ctx->in_ptr_all.p + ctx->lo : nullptr; a.in_rec = ctx->in_rec.p; a.in_base = ctx->in_base;
    a.neg_alias = ctx->neg_alias.p;
    a.sec_alias = (ctx->prm.flags & ANNEMBED_FLAG_NODE_ALIAS) ? nullptr : ctx->sec_alias.p;
    a.line_t1 = nullptr; a.line_t2 = nullptr;         // set by the event-kernel launches (set_event_negative_groups)
    a.cum = ctx->cum.p;
    a.rowpack = ctx->rowpack.p; a.erank = ctx->erank.p;
    a.fired = (ctx->nranks == 1 && !(ctx->prm.flags & ANNEMBED_FLAG_REPLAY_IN_EDGES)) ? ctx->fired.p : nullptr;
    a.n_peers = 0;
    for (int r = 0; r < 7; r++) a.peer_next[r] = nullptr;
    a.k2 = (uint32_t)(ctx->prm.seed & 0xFFFFFFFFu) ^ ((uint32_t)(ctx->prm.seed >> 32) * 0x85EBCA6Bu);
    a.n = (uint32_t)ctx->n; a.lo = ctx->lo; a.hi = ctx->hi;
    a.epoch = epoch;
    a.ukey = epoch_ukey(epoch, a.k2);
    a.k0 = (uint32_t)(ctx->prm.seed & 0xFFFFFFFFu); a.k1 = (uint32_t)(ctx->prm.seed >> 32);
    // expected firings of edge e per mini-epoch: nb_sampling_by_edge * E * (p_e / n) / M   (embedder.rs:858,987)
    a.kappa = (float)((double)ctx->prm.nb_sampling_by_edge * ((double)ctx->E / (double)ctx->n) / (double)M);
    a.K = make_const(ctx, grad_step);
    return a;
}

// The gathers of the layout are the only reuse in the epoch kernel (7 random rows per sample); everything else
// streams.  Pin the snapshot in L2 (persisting access-policy window) so that the streaming arrays do not evict it.
static void set_l2_window(annembed_cuda_ctx *ctx, const void *ptr, size_t bytes)
{
    if (ctx->l2_persist_max <= 0 || ctx->l2_window_max <= 0 || (ctx->prm.flags & ANNEMBED_FLAG_NO_L2_PERSIST)) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    const size_t win = std::min(bytes, (size_t)ctx->l2_window_max);
    v.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = win ? (float)std::min(1.0, (double)ctx->l2_persist_max / (double)win) : 0.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaStreamSetAttribute(ctx->stream2, cudaStreamAttributeAccessPolicyWindow, &v);
}

// The tiled kernels are specialised for b == 1, keep the per-edge firing counts of a node in bytes and hold rows of
// at most 16 neighbours in registers; everything else runs the thread-per-node kernel.  ONE predicate decides both the
// kernel and (multi-rank) whether the exchange is fused into it.
static bool use_tiled(const annembed_cuda_ctx *ctx, float kappa)
{
    return !(ctx->prm.flags & ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL) && ctx->prm.b == 1.0 && ctx->KP != 0 &&
           kappa + 2.0f < (float)EpochTile<2, 6>::MAX_FIRINGS;
}

template <int DP, bool HUB, int KP>
static cudaError_t launch_tiled(annembed_cuda_ctx *ctx, const EpochArgs &a)
{
    using TL = EpochTile<DP, KP>;
    const uint64_t tiles = ((uint64_t)(a.hi - a.lo) + 31) / 32;
    const unsigned int nb = (unsigned int)((tiles + TL::WARPS - 1) / TL::WARPS);
    k_epoch_out<DP, HUB, KP><<<nb, TL::WARPS * 32, 0, ctx->launch_stream>>>(a, ctx->counter.p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const unsigned int nb2 = (unsigned int)((tiles + ANNEMBED_WARPS_IN - 1) / ANNEMBED_WARPS_IN);
    if (InTile<DP>::SMEM > 48 * 1024) {
        e = cudaFuncSetAttribute(k_epoch_in<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, InTile<DP>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_epoch_in_flags<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, InTile<DP>::SMEM);
        if (e != cudaSuccess) return e;
    }
    if (a.fired) k_epoch_in_flags<DP><<<nb2, ANNEMBED_WARPS_IN * 32, InTile<DP>::SMEM, ctx->launch_stream>>>(a);
    else k_epoch_in<DP><<<nb2, ANNEMBED_WARPS_IN * 32, InTile<DP>::SMEM, ctx->launch_stream>>>(a);
    return cudaGetLastError();
}

template <int DP, bool HUB>
static cudaError_t launch_epoch_dp(annembed_cuda_ctx *ctx, const EpochArgs &a)
{
    if (a.hi <= a.lo) return cudaSuccess;
    const bool tiled = use_tiled(ctx, a.kappa);
    ctx->last_epoch_kernels = tiled ? 2 : 1;
    if (tiled) {
        switch (ctx->KP) {
        case 6: return launch_tiled<DP, HUB, 6>(ctx, a);      // row length of the reference's examples
        case 8: return launch_tiled<DP, HUB, 8>(ctx, a);
        case 10: return launch_tiled<DP, HUB, 10>(ctx, a);
        default: return launch_tiled<DP, HUB, 16>(ctx, a);
        }
    }
    k_epoch_generic<DP, HUB><<<nblocks(a.hi - a.lo, 256), 256, 0, ctx->launch_stream>>>(a, ctx->counter.p);
    return cudaGetLastError();
}

template <bool HUB>
static cudaError_t launch_epoch(annembed_cuda_ctx *ctx, const EpochArgs &a)
{
    switch (ctx->DP) {
    case 2: return launch_epoch_dp<2, HUB>(ctx, a);
    case 4: return launch_epoch_dp<4, HUB>(ctx, a);
    case 8: return launch_epoch_dp<8, HUB>(ctx, a);
    case 16: return launch_epoch_dp<16, HUB>(ctx, a);
    default: return launch_epoch_dp<32, HUB>(ctx, a);
    }
}

static int use_hubness(annembed_cuda_ctx *ctx, bool *hub)
{
    *hub = ctx->prm.hubness_weighting != 0;
    if (*hub) REQUIRE(ctx->have_alias, ANNEMBED_ERR_STATE, "hubness_weighting set but set_neg_weights was not called");
    return ANNEMBED_OK;
}

// closes a mini-epoch across the ranks once the peer stores of k_epoch_in are issued (kernel completion makes them visible)
static int rank_barrier(annembed_cuda_ctx *ctx)
{
    ncclResult_t r = g_nccl.AllReduce(ctx->barrier_buf.p, ctx->barrier_buf.p + 1, 1, ncclFloat, ncclSum, ctx->comm, ctx->stream);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce (barrier): ") + g_nccl.GetErrorString(r); return ANNEMBED_ERR_COMM; }
    return ANNEMBED_OK;
}

// ---- the cell-resident form of K4 (cell_epoch.cuh) ---------------------------------------------------------------
// shared memory of one launch; `events`: the event form of phase A (kappa <= 1), which has no staging buffers
static size_t cell_smem_bytes(const annembed_cuda_ctx *ctx, bool events)
{
    const int W = (events ? ANNEMBED_CELL_THREADS_EVENTS : ANNEMBED_CELL_THREADS) / 32;
    const int kp = events ? 32 : ctx->KP;
    size_t fixed;
    if (ctx->DP <= 2) fixed = kp == 6 ? cell_smem_fixed_bytes<2, 6>(W) : (kp == 8 ? cell_smem_fixed_bytes<2, 8>(W) : cell_smem_fixed_bytes<2, 16>(W));
    else fixed = kp == 6 ? cell_smem_fixed_bytes<4, 6>(W) : (kp == 8 ? cell_smem_fixed_bytes<4, 8>(W) : cell_smem_fixed_bytes<4, 16>(W));
    return fixed + (events ? ctx->cell_map_bytes / 8 + 32 : ctx->cell_map_bytes);   // bitmap / byte map of the firing counts
}
static bool cell_events(float kappa) { return kappa <= 1.0f; }
// The cell kernel serves what the tiled pair serves (b == 1, rows of at most 16 neighbours, byte firing counts) for
// layouts of dimension <= 4, when the largest in-edge byte map of a cell fits in shared memory beside the positions.
static bool use_cells(const annembed_cuda_ctx *ctx, float kappa)
{
    if (ctx->prm.flags & (ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL | ANNEMBED_FLAG_LEGACY_EPOCH_KERNELS | ANNEMBED_FLAG_REPLAY_IN_EDGES)) return false;
    if (ctx->prm.b != 1.0 || ctx->KP == 0 || ctx->DP > 4 || !(kappa + 2.0f < (float)EpochTile<2, 6>::MAX_FIRINGS)) return false;
    return cell_smem_bytes(ctx, cell_events(kappa)) <= (size_t)ctx->smem_optin_max;
}
// Sub-steps (mini-epochs) per launch.  Inside a launch the partners of the same cell are read at their current
// positions; partners in other cells and the negatives are as old as the launch.  With every edge inside its cell a
// whole batch can run in one launch; the more in-edges cross cells, the shorter the launches.
static uint32_t cell_substeps(const annembed_cuda_ctx *ctx, uint32_t M)
{
    uint32_t S = ctx->prm.cell_substeps;
    if (S == 0) {
        const double f = ctx->E ? (double)ctx->ext_cnt_global / (double)ctx->E : 0.0;
        S = f <= 0.01 ? ANNEMBED_CELL_SUBSTEPS_LOCAL : (f <= 0.10 ? 4u : (f <= 0.30 ? 2u : 1u));
    }
    return std::max(1u, std::min(std::min(S, M), (uint32_t)ANNEMBED_CELL_MAX_SUBSTEPS));
}

template <int DP, bool HUB, int KP>
static cudaError_t launch_cells_kp(annembed_cuda_ctx *ctx, const CellArgs &A)
{
    const bool events = cell_events(A.e.kappa);
    const size_t smem = cell_smem_bytes(ctx, events);
    cudaError_t e;
    if (events) {
        e = cudaFuncSetAttribute(k_cell_epochs<DP, HUB, KP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_cell_epochs<DP, HUB, KP, true><<<ctx->cell_hi - ctx->cell_lo, ANNEMBED_CELL_THREADS_EVENTS, smem, ctx->launch_stream>>>(A, ctx->counter.p);
    } else {
        e = cudaFuncSetAttribute(k_cell_epochs<DP, HUB, KP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_cell_epochs<DP, HUB, KP, false><<<ctx->cell_hi - ctx->cell_lo, ANNEMBED_CELL_THREADS, smem, ctx->launch_stream>>>(A, ctx->counter.p);
    }
    return cudaGetLastError();
}
template <int DP, bool HUB>
static cudaError_t launch_cells_dp(annembed_cuda_ctx *ctx, const CellArgs &A)
{
    switch (ctx->KP) {
    case 6: return launch_cells_kp<DP, HUB, 6>(ctx, A);
    case 8: return launch_cells_kp<DP, HUB, 8>(ctx, A);
    case 10: return launch_cells_kp<DP, HUB, 10>(ctx, A);
    default: return launch_cells_kp<DP, HUB, 16>(ctx, A);
    }
}
static cudaError_t launch_cells(annembed_cuda_ctx *ctx, const CellArgs &A, bool hub)
{
    if (ctx->cell_hi <= ctx->cell_lo) return cudaSuccess;
    if (ctx->DP <= 2) return hub ? launch_cells_dp<2, true>(ctx, A) : launch_cells_dp<2, false>(ctx, A);
    return hub ? launch_cells_dp<4, true>(ctx, A) : launch_cells_dp<4, false>(ctx, A);
}

// ---- the asynchronous form of K4 (async_sweep.cuh) ---------------------------------------------------------------
static bool async_tiled(const annembed_cuda_ctx *ctx, float kappa)
{
    return !(ctx->prm.flags & ANNEMBED_FLAG_GENERIC_EPOCH_KERNEL) && ctx->prm.b == 1.0 && ctx->KP != 0 &&
           kappa + 2.0f < (float)EpochTile<2, 6>::MAX_FIRINGS;
}
// Nodes in flight are capped at a fraction of n (persistent warps; large graphs fill the machine): what a sample misses of
// the other nodes' moves is the samples in flight, and the layout statistics feel it when that is more than a few per
// cent of the nodes (measured against the serial reference, DESIGN.md 4: the kNN-preservation rate moves by -0.09 %
// per per cent of the nodes in flight in dimension 2, by +0.35 % per per cent in dimension 15).
static unsigned int async_blocks(const annembed_cuda_ctx *ctx, uint64_t units, unsigned int nodes_in_flight_per_block, int DP)
{
    const uint64_t div = DP <= 4 ? ANNEMBED_ASYNC_WINDOW_DIV : ANNEMBED_ASYNC_WINDOW_DIV_WIDE;
    const uint64_t window = std::max<uint64_t>(nodes_in_flight_per_block, ctx->n / div);
    const uint64_t cap = std::max<uint64_t>(1, window / nodes_in_flight_per_block);
    return (unsigned int)std::min<uint64_t>(units, cap);
}
// visiting order of a sweep's tiles (TileOrder, async_sweep.cuh): multiplier ~ tiles / golden ratio, coprime with tiles
static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; }
static PeerMap peer_map(const annembed_cuda_ctx *ctx)
{
    PeerMap pm;
    memset(&pm, 0, sizeof pm);
    pm.nranks = (uint32_t)ctx->nranks;
    for (int r = 0; r < ctx->nranks && r < 8; r++) { pm.y[r] = ctx->nranks > 1 ? ctx->peer_y[r][0] : ctx->y[0].p; pm.lo[r] = ctx->shard_lo[r]; }
    pm.lo[std::min(ctx->nranks, 8)] = (uint32_t)ctx->n;
    return pm;
}
static TileOrder tile_order(uint64_t tiles, uint64_t warps_total)
{
    TileOrder o;
    o.tiles = (uint32_t)tiles;
    uint32_t mul = (uint32_t)std::max<double>(1.0, std::floor((double)tiles * 0.6180339887498949));
    while (mul > 1 && gcd_u32(mul, o.tiles) != 1) mul--;
    o.mul = tiles > 1 ? mul : 0u;
    o.step = tiles ? (uint32_t)((warps_total * (uint64_t)o.mul) % tiles) : 0u;
    return o;
}
// event kernels (kappa <= 1), uniform sampler: the nodes whose rows share a 128-byte line of the layout share their
// negative streams (sgd_core.cuh neg_stream_key); ANNEMBED_FLAG_SECTOR_NEGATIVES keeps the 4-node groups (A/B)
static uint32_t event_neg_group_shift(const annembed_cuda_ctx *ctx)
{
    if (ctx->prm.flags & ANNEMBED_FLAG_SECTOR_NEGATIVES) return 0u;
    return ctx->DP == 2 ? 4u : (ctx->DP == 4 ? 3u : 0u);
}
static void set_event_negative_groups(const annembed_cuda_ctx *ctx, EpochArgs &a)
{
    a.neg_group_shift = event_neg_group_shift(ctx);
    if (a.neg_group_shift > 2 && a.sec_alias && ctx->line_t1.p && ctx->line_t2.p) { a.line_t1 = ctx->line_t1.p; a.line_t2 = ctx->line_t2.p; }
}
template <int DP, bool HUB, int KP>
static cudaError_t launch_async_kp(annembed_cuda_ctx *ctx, const EpochArgs &a_in, uint32_t subs)
{
    EpochArgs a = a_in;
    const uint64_t owned = (uint64_t)(a.hi - a.lo);
#ifndef ANNEMBED_ASYNC_POISSON
    if (a.kappa <= 1.0f) {
        set_event_negative_groups(ctx, a);
        // persistent warps: at most the resident blocks, the in-flight window, and ~8 firing tiles per warp and launch
        const uint64_t tiles = (owned + 31) / 32;
        if constexpr (DP <= 4) {
            // Hubness sampler: the pipeline through cp.async groups (the table reads are a third dependent round of memory
            // accesses inside the gather stage, and the register pipeline exposes every round: 101 G against 82 G edge
            // updates/s).  Uniform sampler: the two pipelines perform the same (199 / 200 G: the memory system is the bound);
            // the register kernel is the default, ANNEMBED_FLAG_CP_ASYNC_PIPELINE selects the other one.
            if (HUB || (ctx->prm.flags & ANNEMBED_FLAG_CP_ASYNC_PIPELINE)) {
                using TC = EventCp<DP, KP>;
                static int resident_per_sm = 0;
                if (!resident_per_sm) {
                    cudaError_t e = cudaFuncSetAttribute(k_sweep_events_cp<DP, HUB, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::SMEM);
                    if (e != cudaSuccess) return e;
                    int nbsm = 0;
                    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, k_sweep_events_cp<DP, HUB, KP>, TC::THREADS, TC::SMEM);
                    if (e != cudaSuccess) return e;
                    resident_per_sm = std::max(1, nbsm);
                }
                const unsigned int resident = (unsigned int)(ctx->sm_count * resident_per_sm);
                const uint64_t want = std::max<uint64_t>(1, (uint64_t)((double)tiles * a.kappa * (double)subs / 8.0) / TC::WARPS);
                const unsigned int nb = (unsigned int)std::min<uint64_t>(std::min<uint64_t>(want, resident), async_blocks(ctx, tiles, TC::THREADS * TC::VISITS, DP));
                k_sweep_events_cp<DP, HUB, KP><<<nb, TC::THREADS, TC::SMEM, ctx->launch_stream>>>(a, ctx->y[0].p, peer_map(ctx), tile_order(tiles, (uint64_t)nb * TC::WARPS),
                                                                                               subs, ctx->counter.p);
                return cudaGetLastError();
            }
        }
      if constexpr (!(DP <= 4 && HUB)) {
        using TE = EventTile<DP, KP>;
        const unsigned int resident = (unsigned int)(ctx->sm_count * TE::MINB);
        const uint64_t want = std::max<uint64_t>(1, (uint64_t)((double)tiles * a.kappa * (double)subs / 8.0) / TE::WARPS);
        const unsigned int nb = (unsigned int)std::min<uint64_t>(std::min<uint64_t>(want, resident), async_blocks(ctx, tiles, TE::WARPS * 32 * TE::VISITS, DP));
        k_sweep_events<DP, HUB, KP><<<nb, TE::WARPS * 32, 0, ctx->launch_stream>>>(a, ctx->y[0].p, peer_map(ctx), tile_order(tiles, (uint64_t)nb * TE::WARPS), subs,
                                                                                ctx->counter.p);
        return cudaGetLastError();
      }
    }
#endif
    (void)subs;                                      // several firings per visit: one sweep per launch
    using TL = EpochTile<DP, KP>;
    const uint64_t tiles = (owned + 31) / 32;
    const unsigned int nb = async_blocks(ctx, (tiles + TL::WARPS - 1) / TL::WARPS, TL::WARPS * 32, DP);
    k_sweep_async<DP, HUB, KP><<<nb, TL::WARPS * 32, 0, ctx->launch_stream>>>(a, ctx->y[0].p, peer_map(ctx), tile_order(tiles, (uint64_t)nb * TL::WARPS), ctx->counter.p);
    return cudaGetLastError();
}
template <int DP, bool HUB>
static cudaError_t launch_async_dp(annembed_cuda_ctx *ctx, const EpochArgs &a, uint32_t subs)
{
    if (a.hi <= a.lo) return cudaSuccess;
    if (async_tiled(ctx, a.kappa)) {
        switch (ctx->KP) {
        case 6: return launch_async_kp<DP, HUB, 6>(ctx, a, subs);
        case 8: return launch_async_kp<DP, HUB, 8>(ctx, a, subs);
        case 10: return launch_async_kp<DP, HUB, 10>(ctx, a, subs);
        default: return launch_async_kp<DP, HUB, 16>(ctx, a, subs);
        }
    }
    const uint64_t tiles = ((uint64_t)(a.hi - a.lo) + 31) / 32;
    const unsigned int nb = async_blocks(ctx, (tiles + 3) / 4, 128, DP);
    k_sweep_async_generic<DP, HUB><<<nb, 128, 0, ctx->launch_stream>>>(a, ctx->y[0].p, peer_map(ctx), tile_order(tiles, (uint64_t)nb * 4), ctx->counter.p);
    return cudaGetLastError();
}
template <bool HUB>
static cudaError_t launch_async(annembed_cuda_ctx *ctx, const EpochArgs &a, uint32_t subs)
{
    switch (ctx->DP) {
    case 2: return launch_async_dp<2, HUB>(ctx, a, subs);
    case 4: return launch_async_dp<4, HUB>(ctx, a, subs);
    case 8: return launch_async_dp<8, HUB>(ctx, a, subs);
    case 16: return launch_async_dp<16, HUB>(ctx, a, subs);
    default: return launch_async_dp<32, HUB>(ctx, a, subs);
    }
}
// Several ranks, asynchronous form: the owners' rows are copied to all replicas every X launches (X samples per node).
// In between a rank reads the other ranks' nodes as of the last exchange: harmless for the negatives (measured with the
// cell kernel: up to 2 samples per edge), a loss of fidelity for the positive edges that cross ranks -- the fewer of
// those, the longer the period.
static uint32_t async_exchange_every(const annembed_cuda_ctx *ctx)
{
    const double f = ctx->E ? (double)ctx->cross_rank_edges / (double)ctx->E : 0.0;
    return f <= 0.02 ? 8u : (f <= 0.10 ? 4u : (f <= 0.30 ? 2u : 1u));
}
// sub-sweeps per launch of the asynchronous form: the default schedule runs its ANNEMBED_ASYNC_THIN thinned sub-sweeps of a
// sweep in one launch (k_sweep_events); an explicit mini_epochs_per_batch makes every sweep its own launch
// One launch runs several sweeps' worth of sub-sweeps back to back inside its persistent warps (a warp always visits the
// same tiles, so the order of a node's samples is kept; the warps drift against each other, which the asynchronous form
// does not mind): ANNEMBED_ASYNC_SWEEPS_PER_LAUNCH sweeps on one rank, the sweeps between two row exchanges on several.
#ifndef ANNEMBED_ASYNC_SWEEPS_PER_LAUNCH
#define ANNEMBED_ASYNC_SWEEPS_PER_LAUNCH 4u
#endif
static uint32_t async_subs(const annembed_cuda_ctx *ctx, float kappa)
{
    if (ctx->prm.mini_epochs_per_batch || !async_tiled(ctx, kappa) || kappa > 1.0f) return 1u;
    return ANNEMBED_ASYNC_THIN * (ctx->nranks > 1 ? async_exchange_every(ctx) : ANNEMBED_ASYNC_SWEEPS_PER_LAUNCH);
}

// replicate the rows every rank owns (rank-dependent counts): one broadcast per rank, grouped
static int exchange_rows_nccl(annembed_cuda_ctx *ctx, float *buf)
{
    ncclResult_t r = g_nccl.GroupStart();
    for (int k = 0; k < ctx->nranks && r == ncclSuccess; k++) {
        const size_t off = (size_t)ctx->shard_lo[k] * ctx->DP, cnt = (size_t)(ctx->shard_lo[k + 1] - ctx->shard_lo[k]) * ctx->DP;
        if (cnt) r = g_nccl.Broadcast(buf + off, buf + off, cnt, ncclFloat, k, ctx->comm, ctx->stream);
    }
    const ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { ctx->err = std::string("ncclBroadcast (row exchange): ") + g_nccl.GetErrorString(r); return ANNEMBED_ERR_COMM; }
    return ANNEMBED_OK;
}

// asynchronous form: the parts are equal (n_pad rows per rank): one in-place all-gather of the owners' rows
static int exchange_rows_allgather(annembed_cuda_ctx *ctx, float *buf, cudaStream_t stream)
{
    const size_t cnt = (size_t)ctx->n_pad * ctx->DP;
    ncclResult_t r = g_nccl.AllGather(buf + (size_t)ctx->rank * cnt, buf, cnt, ncclFloat, ctx->comm, stream);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllGather (row exchange): ") + g_nccl.GetErrorString(r); return ANNEMBED_ERR_COMM; }
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_optimize_batches(annembed_cuda_ctx *ctx, uint32_t first_batch, uint32_t n_batches)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(ctx->have_graph, ANNEMBED_ERR_STATE, "optimize: graph not set");
    REQUIRE(ctx->have_embedding, ANNEMBED_ERR_STATE, "optimize: initial embedding not set");
    REQUIRE(first_batch >= 1, ANNEMBED_ERR_INVALID_ARG, "batches are numbered from 1 (embedder.rs:873)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    bool hub;
    if ((rc = use_hubness(ctx, &hub))) return rc;
    if ((rc = ensure_build(ctx))) return rc;
    const uint32_t nb = ctx->prm.nb_grad_batch;
    const uint32_t last = std::min<uint64_t>((uint64_t)first_batch + n_batches, (uint64_t)nb + 1);   // exclusive
    size_t n_launch = 0;
    for (uint32_t iter = first_batch; iter < last; iter++) {
        const uint32_t M = mini_epochs_of_batch(ctx, iter);
        const float kappa = (float)((double)ctx->prm.nb_sampling_by_edge * ((double)ctx->E / (double)ctx->n) / (double)M);
        const uint32_t S = use_async(ctx) ? async_subs(ctx, kappa) : (use_cells(ctx, kappa) ? cell_substeps(ctx, M) : 1u);
        n_launch += (M + S - 1) / S;
    }
    while (ctx->ev.size() < 2 * n_launch + 2 * n_launch * (ctx->nranks > 1)) {
        cudaEvent_t e; CU(cudaEventCreate(&e)); ctx->ev.push_back(e);
    }
    CU(cudaMemsetAsync(ctx->counter.p, 0, 256 * sizeof(unsigned long long), ctx->stream));
    CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    // the caller's layout, in the internal node order, becomes the first snapshot
    const uint64_t nrow = ctx->n * (uint64_t)ctx->DP;
    if (ctx->nranks > 1 && ctx->have_peers && (rc = rank_barrier(ctx))) return rc;    // no peer still reads or writes y[]
    ctx->cur = 0;
    k_rows_to_internal<<<nblocks(nrow, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->DP, ctx->old_of_new.p, ctx->yapi.p, ctx->y[0].p);
    ctx->st.kernel_launches++;
    if (ctx->nranks > 1 && ctx->have_peers && (rc = rank_barrier(ctx))) return rc;    // every replica holds the snapshot
    size_t li = 0;
    ctx->last_substeps = 0;
    const size_t xoff = 2 * n_launch;
    size_t n_kernels = 0, n_exchanges = 0;
    bool forked = false;
    for (uint32_t iter = first_batch; iter < last; iter++) {
        const double grad_step = ctx->prm.grad_step * (1.0 - (double)iter / (double)nb);   // embedder.rs:875
        const uint32_t M = mini_epochs_of_batch(ctx, iter), e0 = first_epoch_of_batch(ctx, iter);
        uint32_t S = 1;
        const bool async = use_async(ctx);
        {
            const EpochArgs probe = make_epoch_args(ctx, e0, grad_step, M);
            if (async) S = async_subs(ctx, probe.kappa);
            else if (use_cells(ctx, probe.kappa)) S = cell_substeps(ctx, M);
        }
        for (uint32_t m = 0; m < M; m += S, li++) {
            EpochArgs a = make_epoch_args(ctx, e0 + m, grad_step, M);
            if (async) {                                       // one sweep over the single layout buffer y[0]
                a.y_snap = ctx->y[0].p; a.y_next = ctx->y[0].p; a.fired = nullptr;
                set_l2_window(ctx, a.y_snap, (size_t)ctx->n * ctx->DP * sizeof(float));
                CU(cudaEventRecord(ctx->ev[2 * li], ctx->stream));
                CU(hub ? launch_async<true>(ctx, a, std::min(S, M - m)) : launch_async<false>(ctx, a, std::min(S, M - m)));
                CU(cudaEventRecord(ctx->ev[2 * li + 1], ctx->stream));
                n_kernels += 1;
                if (ctx->nranks > 1) {
                    const bool last_launch = iter + 1 == last && m + S >= M;
                    // default schedule: a launch holds the sweeps between two exchanges; explicit schedule (one sweep per
                    // launch): every async_exchange_every launches.  No barrier: the owner's replica accumulates every
                    // reduction whenever it lands, a late one simply travels with the next exchange.  The exchange runs on
                    // the second stream, UNDER the next launch: a sweep reads the other ranks' rows at whatever age they
                    // have and only ever writes rows through reductions on their owners, so rows that change under it (or
                    // travel while they are reduced into) are part of the asynchronous semantics.
                    const uint32_t every = S > 1 ? 1u : async_exchange_every(ctx);
                    const bool xchg = (li + 1) % every == 0 || last_launch;
                    if (xchg && !last_launch) {
                        n_exchanges++;
                        CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
                        CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
                        CU(cudaEventRecord(ctx->ev[xoff + 2 * li], ctx->stream2));
                        if ((rc = exchange_rows_allgather(ctx, ctx->y[0].p, ctx->stream2))) return rc;
                        CU(cudaEventRecord(ctx->ev[xoff + 2 * li + 1], ctx->stream2));
                        forked = true;
                    } else {
                        if (last_launch && forked) {               // join the exchanges in flight
                            CU(cudaEventRecord(ctx->ev_join, ctx->stream2));
                            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
                        }
                        CU(cudaEventRecord(ctx->ev[xoff + 2 * li], ctx->stream));
                        if (xchg) {
                            // the last exchange closes the run: every rank's reductions must have landed before the rows travel
                            n_exchanges++;
                            if ((rc = rank_barrier(ctx))) return rc;
                            if ((rc = exchange_rows_allgather(ctx, ctx->y[0].p, ctx->stream))) return rc;
                        }
                        CU(cudaEventRecord(ctx->ev[xoff + 2 * li + 1], ctx->stream));
                    }
                }
                continue;
            }
            const bool cells = use_cells(ctx, a.kappa);
            // the fused exchange lives in the tiled in-edge kernel / the cell kernel: the same predicate picks both
            const bool fused = ctx->nranks > 1 && ctx->have_peers && (cells || use_tiled(ctx, a.kappa));
            if (fused) {
                for (int r = 0; r < ctx->nranks; r++)
                    if (r != ctx->rank) a.peer_next[a.n_peers++] = ctx->peer_y[r][ctx->cur ^ 1];
            }
            set_l2_window(ctx, a.y_snap, (size_t)ctx->n * ctx->DP * sizeof(float));
            CU(cudaEventRecord(ctx->ev[2 * li], ctx->stream));
            if (cells) {
                CellArgs A;
                A.e = a;
                A.e.in_ptr = ctx->in_ptr_all.p;                // indexed by node id
                A.e.fired = nullptr;
                A.cell_start = ctx->cell_start.p; A.cell_lo = ctx->cell_lo;
                A.substeps = std::min(S, M - m);
                A.ext_ptr = ctx->ext_ptr.p; A.ext_slot = ctx->ext_slot.p;
                CU(launch_cells(ctx, A, hub));
                n_kernels += 1;
                ctx->last_substeps = S;
            } else {
                CU(hub ? launch_epoch<true>(ctx, a) : launch_epoch<false>(ctx, a));
                n_kernels += ctx->last_epoch_kernels;
            }
            CU(cudaEventRecord(ctx->ev[2 * li + 1], ctx->stream));
            if (ctx->nranks > 1) {
                CU(cudaEventRecord(ctx->ev[xoff + 2 * li], ctx->stream));
                n_exchanges++;
                if (fused) {
                    // the kernel already stored the owned rows into every replica: only a barrier is left
                    if ((rc = rank_barrier(ctx))) return rc;
                } else {
                    if ((rc = exchange_rows_nccl(ctx, ctx->y[ctx->cur ^ 1].p))) return rc;
                }
                CU(cudaEventRecord(ctx->ev[xoff + 2 * li + 1], ctx->stream));
            }
            ctx->cur ^= 1;
        }
    }
    set_l2_window(ctx, nullptr, 0);
    k_rows_from_internal<<<nblocks(nrow, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->DP, ctx->old_of_new.p, ctx->y[ctx->cur].p, ctx->yapi.p);
    ctx->st.kernel_launches++;
    CU(cudaEventRecord(ctx->ev_b, ctx->stream));
    unsigned long long cnts[256];
    CU(cudaMemcpyAsync(cnts, ctx->counter.p, sizeof(cnts), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    unsigned long long cnt = 0;
    for (int i = 0; i < 256; i++) cnt += cnts[i];
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    double kms = 0, xms = 0;
    for (size_t i = 0; i < n_launch; i++) {
        float t = 0; CU(cudaEventElapsedTime(&t, ctx->ev[2 * i], ctx->ev[2 * i + 1])); kms += t;
        if (ctx->nranks > 1) { CU(cudaEventElapsedTime(&t, ctx->ev[xoff + 2 * i], ctx->ev[xoff + 2 * i + 1])); xms += t; }
    }
    ctx->st.optimize_ms = ms; ctx->st.epoch_kernel_ms = kms; ctx->st.exchange_ms = xms;
    ctx->st.epoch_launches = n_launch; ctx->st.kernel_launches += n_kernels; ctx->st.exchanges = n_exchanges;
    ctx->st.positive_samples = cnt; ctx->st.edge_updates = 6 * cnt;
    ctx->st.model_bytes = (double)cnt * (12.0 + 36.0 * (double)ctx->prm.asked_dim);
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_cross_entropy(annembed_cuda_ctx *ctx, double *out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(out, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_embedding, ANNEMBED_ERR_STATE, "cross_entropy: embedding not set");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_build(ctx))) return rc;
    CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    const unsigned int nb = std::min<unsigned int>(2048u, std::max(1u, nblocks(ctx->hi - ctx->lo, 256)));
    const float *Y = ctx->yapi.p;
#define LAUNCH_CE(DPV) k_cross_entropy<DPV><<<nb, 256, 0, ctx->stream>>>(ctx->lo, ctx->hi, ctx->row_ptr.p, ctx->col.p, ctx->proba.p, ctx->emb_scale.p, Y, ctx->prm.b, ctx->partials.p)
    switch (ctx->DP) {
    case 2: LAUNCH_CE(2); break;
    case 4: LAUNCH_CE(4); break;
    case 8: LAUNCH_CE(8); break;
    case 16: LAUNCH_CE(16); break;
    default: LAUNCH_CE(32); break;
    }
#undef LAUNCH_CE
    k_final_sum_f64<<<1, 256, 0, ctx->stream>>>(nb, ctx->partials.p, ctx->partials.p + 4095);
    ctx->st.kernel_launches += 2;
    if (ctx->nranks > 1) {
        ncclResult_t r = g_nccl.AllReduce(ctx->partials.p + 4095, ctx->partials.p + 4095, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream);
        if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return ANNEMBED_ERR_COMM; }
    }
    CU(cudaEventRecord(ctx->ev_b, ctx->stream));
    double v = 0;
    CU(cudaMemcpyAsync(&v, ctx->partials.p + 4095, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    ctx->st.cross_entropy_ms = ms;
    ctx->st.d2h_bytes += sizeof(double);
    *out = v;
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_optimize(annembed_cuda_ctx *ctx, double *ce_initial, double *ce_final)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    int rc;
    if (ce_initial && (rc = annembed_cuda_cross_entropy(ctx, ce_initial))) return rc;      // embedder.rs:846
    if ((rc = annembed_cuda_optimize_batches(ctx, 1, ctx->prm.nb_grad_batch))) return rc;  // :873-879
    if (ce_final && (rc = annembed_cuda_cross_entropy(ctx, ce_final))) return rc;          // :885
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_get_embedding(annembed_cuda_ctx *ctx, float *y_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(y_out, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_embedding, ANNEMBED_ERR_STATE, "get_embedding: embedding not set (embedder.rs:384 panics before embed)");
    CU(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n;
    const int d = (int)ctx->prm.asked_dim, DP = ctx->DP;
    if (d == DP) return d2h(ctx, y_out, ctx->yapi.p, n * d * sizeof(float));
    DevBuf<float> stage; CU(stage.alloc(n * d));
    k_unpad_rows<<<nblocks(n * d, 256), 256, 0, ctx->stream>>>(n, d, DP, ctx->yapi.p, stage.p);
    ctx->st.kernel_launches++;
    int rc;
    if ((rc = sync_stream(ctx))) return rc;
    return d2h(ctx, y_out, stage.p, n * d * sizeof(float));
}

extern "C" int annembed_cuda_get_stats(annembed_cuda_ctx *ctx, annembed_cuda_stats *stats)
{
    if (!ctx || !stats) return ANNEMBED_ERR_INVALID_ARG;
    *stats = ctx->st;
    stats->mini_epochs_per_batch = ctx->have_graph ? eff_mini_epochs(ctx) : ctx->prm.mini_epochs_per_batch;
    stats->l2_persist_max_bytes = (uint64_t)std::max(ctx->l2_persist_max, 0);
    stats->l2_window_max_bytes = (uint64_t)std::max(ctx->l2_window_max, 0);
    stats->n_cells = ctx->have_struct ? ctx->n_cells : 0;
    stats->cell_nodes = ctx->cell_nodes;
    stats->cell_substeps = ctx->last_substeps;
    stats->cross_cell_edges = ctx->have_struct ? ctx->ext_cnt_global : 0;
    stats->cross_rank_edges = ctx->have_struct ? ctx->cross_rank_edges : 0;
    return ANNEMBED_OK;
}
extern "C" int annembed_cuda_reset_stats(annembed_cuda_ctx *ctx)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    memset(&ctx->st, 0, sizeof(ctx->st));
    return ANNEMBED_OK;
}

// =====================================================================================================
// N3: quality estimate (embedder.rs:620-753) on the device
// =====================================================================================================
__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
static inline float ordered_to_float(int i) { i ^= ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &i, 4); return f; }

__global__ void k_bbox(uint64_t n, int DP, const float *__restrict__ Y, int *__restrict__ box /* xmin, xmax, ymin, ymax (ordered ints) */)
{
    int xmin = 0x7fffffff, xmax = -0x7fffffff - 1, ymin = 0x7fffffff, ymax = -0x7fffffff - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int x = float_to_ordered(Y[i * DP]), y = float_to_ordered(DP > 1 ? Y[i * DP + 1] : 0.0f);
        xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(box, xmin); atomicMax(box + 1, xmax); atomicMin(box + 2, ymin); atomicMax(box + 3, ymax); }
}

// per-block partials {sum of matches, nodes without match} over nodes and {sum of finite ratios, number of finite ratios} over edges
__global__ void __launch_bounds__(256)
k_quality_reduce(uint64_t n, uint64_t E, const uint32_t *__restrict__ nodes_match, const float *__restrict__ ratio,
                 double *__restrict__ partials /* [gridDim.x][4] */)
{
    double a = 0, b = 0, c = 0, d = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        a += (double)nodes_match[i]; b += nodes_match[i] == 0 ? 1.0 : 0.0;
    }
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (uint64_t)gridDim.x * blockDim.x) {
        const float r = ratio[e];
        if (isfinite(r)) { c += (double)r; d += 1.0; }
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    a = BR(tmp).Sum(a); __syncthreads();
    b = BR(tmp).Sum(b); __syncthreads();
    c = BR(tmp).Sum(c); __syncthreads();
    d = BR(tmp).Sum(d);
    if (threadIdx.x == 0) { partials[4 * blockIdx.x] = a; partials[4 * blockIdx.x + 1] = b; partials[4 * blockIdx.x + 2] = c; partials[4 * blockIdx.x + 3] = d; }
}

// quantiles (numpy "linear" definition) of the first m entries of an ascending device array
static int device_quantiles(annembed_cuda_ctx *ctx, const float *sorted, uint64_t m, double out[6])
{
    static const double qs[6] = {0.05, 0.25, 0.5, 0.75, 0.85, 0.95};
    for (int i = 0; i < 6; i++) out[i] = std::nan("");
    if (m == 0) return ANNEMBED_OK;
    for (int i = 0; i < 6; i++) {
        const double pos = qs[i] * (double)(m - 1);
        const uint64_t lo = (uint64_t)std::floor(pos), hi = std::min<uint64_t>(lo + 1, m - 1);
        float v[2];
        CU(cudaMemcpyAsync(&v[0], sorted + lo, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&v[1], sorted + hi, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        out[i] = (double)v[0] + (pos - (double)lo) * ((double)v[1] - (double)v[0]);
    }
    return ANNEMBED_OK;
}

template <int DP>
static int quality_impl(annembed_cuda_ctx *ctx, uint32_t nbng, annembed_cuda_quality *out, float *radius_out,
                        float *first_dist_out, float *node_ratio_out)
{
    const uint64_t n = ctx->n, E = ctx->E;
    const float *Y = ctx->yapi.p;
    int rc;
    DevBuf<float> t, radius, ratio, node_ratio, first_dist, sorted;
    DevBuf<uint32_t> nodes_match, cell, idx, cell_s, idx_s, cell_start;
    DevBuf<int> box;
    DevBuf<unsigned char> tmp;
    CU(t.alloc(E)); CU(radius.alloc(n)); CU(ratio.alloc(E)); CU(node_ratio.alloc(n)); CU(first_dist.alloc(n));
    CU(nodes_match.alloc(n)); CU(cell.alloc(n)); CU(idx.alloc(n)); CU(cell_s.alloc(n)); CU(idx_s.alloc(n)); CU(box.alloc(4));
    // embedder.rs:478-522
    k_transformed_kgraph<DP><<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->row_ptr.p, ctx->col.p, Y, t.p);
    // bounding box of coordinates (0,1) and the grid: about 2 points per cell
    const int init[4] = {0x7fffffff, -0x7fffffff - 1, 0x7fffffff, -0x7fffffff - 1};
    CU(cudaMemcpyAsync(box.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    k_bbox<<<std::min<unsigned int>(1024u, nblocks(n, 256)), 256, 0, ctx->stream>>>(n, DP, Y, box.p);
    int hb[4];
    CU(cudaMemcpyAsync(hb, box.p, sizeof(hb), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    const float xmin = ordered_to_float(hb[0]), xmax = ordered_to_float(hb[1]), ymin = ordered_to_float(hb[2]), ymax = ordered_to_float(hb[3]);
    REQUIRE(std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax), ANNEMBED_ERR_INVALID_ARG,
            "quality_estimate: the embedding holds non-finite coordinates");
    GridParams gp;
    gp.G = (int)std::min<double>(8192.0, std::max(1.0, std::ceil(std::sqrt((double)n / 2.0))));
    const float width = std::max(std::max(xmax - xmin, ymax - ymin), 1e-30f);
    gp.h = width / (float)gp.G * 1.0001f;
    gp.inv_h = 1.0f / gp.h;
    gp.x0 = xmin; gp.y0 = ymin;
    const uint64_t ncell = (uint64_t)gp.G * gp.G;
    CU(cell_start.alloc(ncell + 2));
    k_cell_of_point<DP><<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, Y, gp, cell.p, idx.p);
    int bits = 1; while (bits < 32 && (1ull << bits) < ncell) bits++;
    size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cell.p, cell_s.p, idx.p, idx_s.p, (int64_t)n, 0, bits, ctx->stream));
    CU(sorted.alloc(std::max(n, E)));
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tb2, radius.p, sorted.p, (int64_t)n, 0, 32, ctx->stream));
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tb3, ratio.p, sorted.p, (int64_t)E, 0, 32, ctx->stream));
    CU(tmp.alloc(std::max(tmp_bytes, std::max(tb2, tb3))));
    tmp_bytes = tmp.n;
    CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, cell.p, cell_s.p, idx.p, idx_s.p, (int64_t)n, 0, bits, ctx->stream));
    k_cell_start<<<nblocks(n + 1, 256), 256, 0, ctx->stream>>>(n, ncell, cell_s.p, cell_start.p);
    // embedder.rs:527-554 (exact instead of HNSW): distance to the nbng-th nearest embedded neighbour
    k_knn_radius<DP><<<nblocks(n, 4), 128, 0, ctx->stream>>>(n, nbng, Y, gp, idx_s.p, cell_s.p, cell_start.p, radius.p);
    // embedder.rs:646-674
    k_quality_per_node<<<nblocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->row_ptr.p, t.p, radius.p, nodes_match.p, ratio.p, node_ratio.p, first_dist.p);
    const unsigned int nb = std::min<unsigned int>(1023u, std::max(1u, nblocks(std::max(n, E), 256)));
    k_quality_reduce<<<nb, 256, 0, ctx->stream>>>(n, E, nodes_match.p, ratio.p, ctx->partials.p);
    ctx->st.kernel_launches += 8;
    std::vector<double> hp(4 * nb);
    CU(cudaMemcpyAsync(hp.data(), ctx->partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = sync_stream(ctx))) return rc;
    double sum_match = 0, n_zero = 0, sum_ratio = 0, n_finite = 0;
    for (unsigned int i = 0; i < nb; i++) { sum_match += hp[4 * i]; n_zero += hp[4 * i + 1]; sum_ratio += hp[4 * i + 2]; n_finite += hp[4 * i + 3]; }
    out->nb_without_match = (uint64_t)n_zero;                                                   // :676-678
    out->mean_nbmatch = sum_match / std::max(1.0, (double)n - n_zero);                          // :679-680
    out->knn_preservation = sum_match / (double)E;
    out->mean_ratio = n_finite > 0 ? sum_ratio / n_finite : std::nan("");                       // :721-726
    size_t tbx = tmp.n;
    CU(cub::DeviceRadixSort::SortKeys(tmp.p, tbx, radius.p, sorted.p, (int64_t)n, 0, 32, ctx->stream));
    if ((rc = device_quantiles(ctx, sorted.p, n, out->radius_quantiles))) return rc;           // :681-690
    tbx = tmp.n;
    CU(cub::DeviceRadixSort::SortKeys(tmp.p, tbx, ratio.p, sorted.p, (int64_t)E, 0, 32, ctx->stream));
    if ((rc = device_quantiles(ctx, sorted.p, (uint64_t)n_finite, out->ratio_quantiles))) return rc;   // :695-714
    ctx->st.kernel_launches += 2;
    if (radius_out && (rc = d2h(ctx, radius_out, radius.p, n * sizeof(float)))) return rc;
    if (first_dist_out && (rc = d2h(ctx, first_dist_out, first_dist.p, n * sizeof(float)))) return rc;   // first_dist.csv :729-735
    if (node_ratio_out && (rc = d2h(ctx, node_ratio_out, node_ratio.p, n * sizeof(float)))) return rc;   // continuity_ratio.csv :737-743
    return ANNEMBED_OK;
}

extern "C" int annembed_cuda_quality_estimate(annembed_cuda_ctx *ctx, uint32_t nbng, annembed_cuda_quality *out,
                                              float *radius_out, float *first_dist_out, float *node_ratio_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(out, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_graph && ctx->have_embedding, ANNEMBED_ERR_STATE, "quality_estimate: graph / embedding not set (embedder.rs:629-632)");
    REQUIRE(nbng >= 1 && (uint64_t)nbng < ctx->n, ANNEMBED_ERR_INVALID_ARG, "nbng must be in 1..n-1");
    REQUIRE(ctx->nranks == 1, ANNEMBED_ERR_UNSUPPORTED, "quality_estimate runs on one GPU");
    CU(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    switch (ctx->DP) {
    case 2: return quality_impl<2>(ctx, nbng, out, radius_out, first_dist_out, node_ratio_out);
    case 4: return quality_impl<4>(ctx, nbng, out, radius_out, first_dist_out, node_ratio_out);
    case 8: return quality_impl<8>(ctx, nbng, out, radius_out, first_dist_out, node_ratio_out);
    case 16: return quality_impl<16>(ctx, nbng, out, radius_out, first_dist_out, node_ratio_out);
    default: return quality_impl<32>(ctx, nbng, out, radius_out, first_dist_out, node_ratio_out);
    }
}

extern "C" int annembed_cuda_debug_draws(annembed_cuda_ctx *ctx, uint32_t epoch, uint32_t *counts_out, uint32_t *neg_out)
{
    if (!ctx) return ANNEMBED_ERR_INVALID_ARG;
    REQUIRE(counts_out, ANNEMBED_ERR_INVALID_ARG, "null output");
    REQUIRE(ctx->have_graph && ctx->have_weights, ANNEMBED_ERR_STATE, "debug_draws: graph / weights not set");
    REQUIRE(ctx->prm.flags & ANNEMBED_FLAG_NO_RELABEL, ANNEMBED_ERR_STATE,
            "debug_draws reports draws by edge of the caller's graph: create the context with ANNEMBED_FLAG_NO_RELABEL");
    CU(cudaSetDevice(ctx->device));
    int rc;
    bool hub;
    if ((rc = use_hubness(ctx, &hub))) return rc;
    DevBuf<uint32_t> dc, dn;
    CU(dc.alloc(ctx->E));
    if (neg_out) CU(dn.alloc(ctx->E * 5));
    if ((rc = ensure_build(ctx))) return rc;
    EpochArgs a = make_epoch_args(ctx, epoch, 0.0);
    // the draws of the kernel that would run: the event kernels of the default schedule widen the uniform sampler's groups
    if (use_async(ctx) && !ctx->prm.mini_epochs_per_batch && a.kappa <= 1.0f) set_event_negative_groups(ctx, a);
    if (hub) k_debug_draws<true><<<nblocks(ctx->n, 128), 128, 0, ctx->stream>>>(a, dc.p, dn.p);
    else k_debug_draws<false><<<nblocks(ctx->n, 128), 128, 0, ctx->stream>>>(a, dc.p, dn.p);
    ctx->st.kernel_launches++;
    if ((rc = sync_stream(ctx))) return rc;
    if ((rc = d2h(ctx, counts_out, dc.p, ctx->E * sizeof(uint32_t)))) return rc;
    if (neg_out && (rc = d2h(ctx, neg_out, dn.p, ctx->E * 5 * sizeof(uint32_t)))) return rc;
    return ANNEMBED_OK;
}
