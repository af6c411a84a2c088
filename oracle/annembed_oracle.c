/*
 * annembed_oracle.c -- CPU restatement of annembed's cross-entropy embedding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under annembed_b200/ (the product) may include, link,
 * import or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * PARITY UNPINNED: the reference (pure Rust, /root/reference) cannot be compiled here (no cargo /
 * rustc, unpinned deps, no network) and its own tests hold no golden vectors for this path
 * (embedder.rs:1435-1467 only asserts `embed().is_ok()`).  The only reference known-answer tests
 * on the path are the two bisection tests (tools/dichotomy.rs:75-91), which test_oracle.py checks
 * against oracle_dichotomy_*.  Everything else is pinned by hand-derived vectors in tests/.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference/src).
 * Written from the behaviour of those lines; no reference source is copied (the reference is Rust).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdatomic.h>
#include <sched.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PROBA_MIN 1.0e-4f /* embedder.rs:50 */

static double wall_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------
 * Edge weights.  tools/kdumap.rs:132-235 (get_scale_from_proba_normalisation) mapped over nodes
 * as tools/kdumap.rs:57-60 does.  All arithmetic f32, sums sequential left-to-right.
 * Returns 0, or -(i+1) if node i has an empty neighbourhood (kdumap.rs:75-85 exits the process).
 * ------------------------------------------------------------------------------------------ */
static void weights_one_node(const uint64_t *row_ptr, const uint32_t *col, const float *dist,
                             uint64_t i, float scale_rho, float beta, float *scale_out,
                             float *p_out /* row_ptr[i]-based */)
{
    const uint64_t lo = row_ptr[i], hi = row_ptr[i + 1];
    const uint64_t k = hi - lo;
    /* kdumap.rs:146-155: mean of first-neighbour distance of every neighbour, then of x itself */
    const float rho_x = dist[lo];
    float sum = 0.0f;
    for (uint64_t m = lo; m < hi; m++) sum += dist[row_ptr[col[m]]];
    sum += rho_x;
    const float mean_rho = sum / (float)(k + 1);
    const float scale = scale_rho * mean_rho; /* :159 */
    *scale_out = scale;
    /* :163-170 last strictly positive distance, scanning from the end */
    int all_equal = 1;
    float last_dist = 0.0f;
    for (uint64_t m = hi; m > lo; m--) {
        if (dist[m - 1] > 0.0f) { last_dist = dist[m - 1]; all_equal = 0; break; }
    }
    if (!all_equal && last_dist > rho_x) {
        /* :172-188  w = max(exp(-(max(d-d0,0)/scale)^beta), PROBA_MIN); f32::max ignores NaN */
        float wsum = 0.0f;
        for (uint64_t m = lo; m < hi; m++) {
            float a = fmaxf(dist[m] - rho_x, 0.0f) / scale;
            float w = expf(-powf(a, beta));
            w = fmaxf(w, PROBA_MIN); /* fmaxf(NaN, x) == x, like Rust's f32::max */
            p_out[m] = w;
        }
        for (uint64_t m = lo; m < hi; m++) wsum += p_out[m]; /* :215 */
        for (uint64_t m = lo; m < hi; m++) p_out[m] /= wsum; /* :216-218 */
    } else {
        /* :224-230 all neighbours at the same distance */
        for (uint64_t m = lo; m < hi; m++) p_out[m] = 1.0f / (float)k;
    }
}

int oracle_edge_weights(uint64_t n, const uint64_t *row_ptr, const uint32_t *col,
                        const float *dist, float scale_rho, float beta, float *scale_out,
                        float *p_out)
{
    for (uint64_t i = 0; i < n; i++)
        if (row_ptr[i + 1] == row_ptr[i]) return -(int)(i + 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++)
        weights_one_node(row_ptr, col, dist, (uint64_t)i, scale_rho, beta, &scale_out[i], p_out);
    return 0;
}

/* tools/nodeparam.rs:88-91  perplexity = exp(-sum p ln p), f32 */
void oracle_perplexity(uint64_t n, const uint64_t *row_ptr, const float *p, float *out)
{
    for (uint64_t i = 0; i < n; i++) {
        float h = 0.0f;
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) h += -p[m] * logf(p[m]);
        out[i] = expf(h);
    }
}

/* ------------------------------------------------------------------------------------------
 * Bisection (dead code in the reference, optional kernel here).
 * tools/dichotomy.rs:4-65.  Returns 0 ok / 1 not converged (Err) / -1 would panic.
 * f is given as a callback so the reference's own two tests can be restated.
 * ------------------------------------------------------------------------------------------ */
typedef float (*oracle_fn)(float, const void *);
int oracle_dichotomy(int increasing, oracle_fn f, const void *ud, float lower_r, float upper_r,
                     float target, float *root)
{
    if (lower_r >= upper_r) return -1;
    float fl = f(lower_r, ud), fu = f(upper_r, ud);
    if (fmaxf(fl, fu) < target || fminf(fu, fl) > target) return -1;
    if (fu < fl && increasing) return -1;
    if (fu > fl && !increasing) return -1;
    float middle = 1.0f, upper = upper_r, lower = lower_r;
    int nbiter = 0;
    while (fabsf(target - f(middle, ud)) > 1.0e-5f) {
        if (increasing) {
            if (f(middle, ud) > target) upper = middle; else lower = middle;
        } else {
            if (f(middle, ud) > target) lower = middle; else upper = middle;
        }
        middle = (lower + upper) * 0.5f;
        if (++nbiter > 100) { *root = middle; return 1; }
    }
    *root = middle;
    return 0;
}
static float f_square(float x, const void *ud) { (void)ud; return x * x; }
static float f_invsquare(float x, const void *ud) { (void)ud; return 1.0f / (x * x); }
/* the reference's two tests, tools/dichotomy.rs:75-91 */
int oracle_dichotomy_test_inc(float *root) { return oracle_dichotomy(1, f_square, 0, 0.f, 5.f, 2.f, root); }
int oracle_dichotomy_test_dec(float *root) { return oracle_dichotomy(0, f_invsquare, 0, 0.2f, 5.f, 0.5f, root); }

struct umap_row { const float *d; uint64_t k; };
static float f_umap(float beta, const void *ud)
{
    const struct umap_row *r = (const struct umap_row *)ud;
    float s = 0.0f;
    for (uint64_t m = 0; m < r->k; m++) s += expf(-(r->d[m] - r->d[0]) * beta);
    return s;
}
/* embedder.rs:760-783 get_scale_from_umap: returns (1/beta, un-normalised exp(-(d-d0) beta)) */
int oracle_scale_from_umap(const float *d, uint64_t k, float norm, float *scale, float *w)
{
    struct umap_row r = { d, k };
    float beta;
    int rc = oracle_dichotomy(0, f_umap, &r, 0.0f, FLT_MAX, norm, &beta);
    if (rc < 0) return rc;
    for (uint64_t m = 0; m < k; m++) w[m] = expf(-(d[m] - d[0]) * beta);
    *scale = 1.0f / beta;
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Embedded scales.  embedder.rs:1356-1373.  f32 sequential sum (use_f64_sum=0, the reference's
 * arithmetic) or f64 sum (use_f64_sum=1; what a more accurate reduction gives).
 * ------------------------------------------------------------------------------------------ */
void oracle_embedded_scales(uint64_t n, const float *scale, int use_f64_sum, float *out)
{
    float mean;
    if (use_f64_sum) {
        double s = 0.0;
        for (uint64_t i = 0; i < n; i++) s += scale[i];
        mean = (float)(s / (double)n);
    } else {
        float s = 0.0f;
        for (uint64_t i = 0; i < n; i++) s += scale[i];
        mean = s / (float)n;
    }
    const float sup = 4.0f, inf = 1.0f / sup, width = 0.2f;
    for (uint64_t i = 0; i < n; i++) out[i] = width * fmaxf(fminf(scale[i] / mean, sup), inf);
}

/* ------------------------------------------------------------------------------------------
 * One SGD sample.  embedder.rs:1167-1302 (ce_optim_edge_shannon) with F = f32 coordinates and
 * f64 coefficients, the negatives given explicitly (already accepted).
 * y: n x d row-major, modified in place.  g is the per-sample gradient scratch (d floats).
 * ------------------------------------------------------------------------------------------ */
static inline double common_coeff(double u, double s2, double b)
{
    /* :1216-1222 and :1276-1282 */
    if (b != 1.0) {
        double cw = 1.0 / (1.0 + pow(u, b));
        return 2.0 * b * cw * pow(u, b - 1.0) / s2;
    }
    return 2.0 * b * (1.0 / (1.0 + u)) / s2;
}

static inline void sgd_sample(float *y, uint32_t d, uint64_t i, uint64_t j, double p, double scale,
                              double b, double grad_step, const uint32_t *negs, int nneg, float *g)
{
    float *yi_g = y + i * d, *yj_g = y + j * d;
    float yi[64], yj[64];
    for (uint32_t c = 0; c < d; c++) { yi[c] = yi_g[c]; yj[c] = yj_g[c]; g[c] = 0.0f; } /* :1186-1199 */
    const double s2 = scale * scale;
    float dsum = 0.0f; /* summed in F, :1206-1211 */
    for (uint32_t c = 0; c < d; c++) dsum += (yi[c] - yj[c]) * (yi[c] - yj[c]);
    const double u = (double)dsum / s2;
    const double coeff = common_coeff(u, s2, b);
    if (u > 0.0) {
        const double alfa = (double)(1.0f / PROBA_MIN);                 /* :1225 */
        const double rep = 1.0 / fmax(u * u, alfa);                     /* :1226 */
        double cij = grad_step * coeff * (-p + (1.0 - p) * rep);
        cij = fmax(cij, -0.49);                                         /* :1228-1229 */
        const float cf = (float)cij;
        for (uint32_t c = 0; c < d; c++) g[c] = (yj[c] - yi[c]) * cf;   /* :1230 */
    }
    for (uint32_t c = 0; c < d; c++) { yi[c] -= g[c]; yj[c] += g[c]; }  /* :1237-1238 */
    for (uint32_t c = 0; c < d; c++) yj_g[c] = yj[c];                   /* :1239 publish y_j */
    for (int q = 0; q < nneg; q++) {                                    /* :1241-1299 */
        const float *yk = y + (uint64_t)negs[q] * d;
        float dk = 0.0f;
        for (uint32_t c = 0; c < d; c++) dk += (yi[c] - yk[c]) * (yi[c] - yk[c]);
        const double dik = (double)dk;
        const double uk = dik / s2;
        const double ck = common_coeff(uk, s2, b);
        if (dik > 0.0) {
            const double rep = 1.0 / fmax(uk * uk, 1.0 / 16.0);         /* :1286-1288 */
            const double cik = fmin(grad_step * ck * rep, 2.0);
            const float cf = (float)cik;
            for (uint32_t c = 0; c < d; c++) g[c] = (yk[c] - yi[c]) * cf;
        } /* else: g keeps its previous value (stale-gradient quirk) */
        for (uint32_t c = 0; c < d; c++) yi[c] -= g[c];                 /* :1297 */
    }
    for (uint32_t c = 0; c < d; c++) yi_g[c] = yi[c];                   /* :1301 publish y_i */
}

/* Apply a fixed list of samples strictly in order (the reference's serial gradient_iteration,
 * embedder.rs:1304-1308, with the random draws replaced by the given list).
 * edge_idx indexes the flat edge list (row_ptr order, embedder.rs:975-984); neg: n_samples x 5. */
int oracle_step_fixed(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col,
                      const float *p, const float *emb_scale, float *y, double b, double grad_step,
                      uint64_t n_samples, const uint64_t *edge_idx, const uint32_t *neg)
{
    if (d > 64) return -1;
    float g[64];
    for (uint64_t s = 0; s < n_samples; s++) {
        const uint64_t e = edge_idx[s];
        /* source node of edge e: binary search in row_ptr */
        uint64_t lo = 0, hi = n;
        while (hi - lo > 1) { uint64_t mid = (lo + hi) / 2; if (row_ptr[mid] <= e) lo = mid; else hi = mid; }
        const uint64_t i = lo, j = col[e];
        sgd_sample(y, d, i, j, (double)p[e], (double)emb_scale[i], b, grad_step, neg + 5 * s, 5, g);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Cross entropy.  embedder.rs:1127-1163 + cauchy_edge_weight :1322-1345 (F = f32).
 * ------------------------------------------------------------------------------------------ */
double oracle_cross_entropy(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col,
                            const float *p, const float *emb_scale, const float *y, double b)
{
    double ce = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ce)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double scale = (double)emb_scale[i];
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) {
            const float *yi = y + (uint64_t)i * d, *yj = y + (uint64_t)col[m] * d;
            float ds = 0.0f;
            for (uint32_t c = 0; c < d; c++) ds += (yi[c] - yj[c]) * (yi[c] - yj[c]);
            double x = (double)ds / (scale * scale);
            x = pow(x, b);
            float wf = (float)(1.0 / (1.0 + x));
            if (!(wf < 1.0f)) wf = 1.0f - FLT_EPSILON;
            const double w = (double)wf, pe = (double)p[m];
            double term = 0.0;
            if (w > 0.0) term += -pe * log(w);
            if (w < 1.0) term += -(1.0 - pe) * log(1.0 - w);
            ce += term;
        }
    }
    return ce;
}

/* ------------------------------------------------------------------------------------------
 * RNG + alias tables for the Hogwild loop.  The reference uses rand::rng() (thread-local ChaCha,
 * OS seeded: not reproducible, embedder.rs:1182,1121) and rand_distr::WeightedAliasIndex
 * (crate not vendored).  Restated with xoshiro256++ per thread and Vose's alias method, which
 * is the published algorithm WeightedAliasIndex implements.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline void rng_seed(rng_t *r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed); }
static inline uint64_t rng_next(rng_t *r)
{
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return (uint64_t)(((__uint128_t)rng_next(r) * n) >> 64); }
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { uint64_t n; float *prob; uint64_t *alias; } alias_t;
static int alias_build(alias_t *a, const float *w, uint64_t n)
{
    a->n = n;
    a->prob = (float *)malloc(n * sizeof(float));
    a->alias = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *small = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *large = (uint64_t *)malloc(n * sizeof(uint64_t));
    double *q = (double *)malloc(n * sizeof(double));
    if (!a->prob || !a->alias || !small || !large || !q) return -1;
    double tot = 0.0;
    for (uint64_t i = 0; i < n; i++) tot += w[i];
    uint64_t ns = 0, nl = 0;
    for (uint64_t i = 0; i < n; i++) {
        q[i] = (double)w[i] * (double)n / tot;
        if (q[i] < 1.0) small[ns++] = i; else large[nl++] = i;
    }
    while (ns > 0 && nl > 0) {
        uint64_t s = small[--ns], l = large[--nl];
        a->prob[s] = (float)q[s]; a->alias[s] = l;
        q[l] = (q[l] + q[s]) - 1.0;
        if (q[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    while (nl > 0) { uint64_t l = large[--nl]; a->prob[l] = 1.0f; a->alias[l] = l; }
    while (ns > 0) { uint64_t s = small[--ns]; a->prob[s] = 1.0f; a->alias[s] = s; }
    free(small); free(large); free(q);
    return 0;
}
static inline uint64_t alias_draw(const alias_t *a, rng_t *r)
{
    uint64_t i = rng_below(r, a->n);
    return (rng_unit(r) < (double)a->prob[i]) ? i : a->alias[i];
}
static void alias_free(alias_t *a) { free(a->prob); free(a->alias); }

/* ------------------------------------------------------------------------------------------
 * The optimizer loop.  embedder.rs:794-904 (driver + schedule), :964-1025 (flat edge list, alias
 * over all E weights), :1117-1123 + :909-931 (negative draw, uniform or alias over neg_w/mean),
 * :1311-1315 (threaded Hogwild iteration).  Rows are plain arrays: the reference's per-row
 * RwLocks only make row reads/writes atomic, which does not change the arithmetic.
 *
 * neg_w: NULL -> uniform negatives; else N weights, already clamp(count,1,N) (embedder.rs:826-833).
 * first_batch/n_batches_to_run let a caller run a bounded slice of the schedule (bench sample):
 * batches iter = first_batch .. first_batch+n_batches_to_run-1 of 1..=nb_grad_batch.
 * sample_fraction scales the samples per batch (1.0 = the reference's nb_sampling_by_edge * E).
 * Returns the number of positive samples processed, or <0 on error.
 * ------------------------------------------------------------------------------------------ */
int64_t oracle_optimize(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col,
                        const float *p, const float *emb_scale, float *y, double b,
                        double grad_step_init, uint32_t nb_sampling_by_edge, uint32_t nb_grad_batch,
                        const float *neg_w, uint64_t seed, uint32_t first_batch,
                        uint32_t n_batches_to_run, double sample_fraction, int n_threads,
                        double *loop_seconds /* nullable: wall time of the sampling loop only */)
{
    if (d > 64) return -1;
    const uint64_t E = row_ptr[n];
    /* flat edge list: source node per edge (embedder.rs:975-984) */
    uint32_t *src = (uint32_t *)malloc(E * sizeof(uint32_t));
    if (!src) return -2;
    for (uint64_t i = 0; i < n; i++)
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) src[m] = (uint32_t)i;
    alias_t pos, negt;
    if (alias_build(&pos, p, E)) return -2;                 /* :987 */
    int have_neg = neg_w != NULL;
    if (have_neg) {
        /* NodeSampler::new normalises by the mean (:916-919); the alias method is scale free */
        if (alias_build(&negt, neg_w, n)) return -2;
    }
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    const uint64_t nb_sample = (uint64_t)((double)nb_sampling_by_edge * (double)E * sample_fraction); /* :858 */
    int64_t done = 0;
    double t_loop = 0.0;
    for (uint32_t iter = first_batch; iter < first_batch + n_batches_to_run && iter <= nb_grad_batch; iter++) {
        const double grad_step = grad_step_init * (1.0 - (double)iter / (double)nb_grad_batch); /* :875 */
        const double t0 = wall_seconds();
#pragma omp parallel
        {
            rng_t rng;
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            rng_seed(&rng, seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)iter << 20) + (uint64_t)tid);
            float g[64];
#pragma omp for schedule(static)
            for (int64_t s = 0; s < (int64_t)nb_sample; s++) {
                const uint64_t e = alias_draw(&pos, &rng);  /* :1182 */
                const uint64_t i = src[e], j = col[e];
                uint32_t negs[5];
                int got = 0;
                while (got < 5) {                           /* :1241-1252 */
                    uint64_t k = have_neg ? alias_draw(&negt, &rng) : rng_below(&rng, n);
                    if (k == i || k == j) continue;
                    int in_row = 0;
                    for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) /* nodeparam.rs:83-85 */
                        if (col[m] == k) { in_row = 1; break; }
                    if (in_row) continue;
                    negs[got++] = (uint32_t)k;
                }
                sgd_sample(y, d, i, j, (double)p[e], (double)emb_scale[i], b, grad_step, negs, 5, g);
            }
        }
        t_loop += wall_seconds() - t0;
        done += (int64_t)nb_sample;
    }
    if (loop_seconds) *loop_seconds = t_loop;
    alias_free(&pos);
    if (have_neg) alias_free(&negt);
    free(src);
    return done;
}

/* ------------------------------------------------------------------------------------------
 * The same loop on the reference's DATA LAYOUT: the timing twin SURVEY.md 8(d) asks for beside the
 * plain-array loop above ("per-node heap rows + reader/writer locks").  Same arithmetic and the
 * same random draws as oracle_optimize (single-threaded the two produce bit-identical layouts:
 * tests/test_oracle.py), but every memory operation of ce_optim_edge_shannon is performed the
 * way the Rust code performs it:
 *   - embedded: Vec<Arc<RwLock<Array1<F>>>> (embedder.rs:941,994-998): one heap block per node
 *     holding the Arc counters, the lock word and the Array1 header, whose data is another heap
 *     block;  get_embedded_data = Arc::clone (:1071-1073): one atomic increment + one decrement
 *     of the row's strong count per access;
 *   - .read().to_owned() (:1186-1187): shared lock (one compare-and-swap), a fresh heap copy of
 *     the row, unlock;  *(...write()) = y (:1239,1301): exclusive lock, the copy BECOMES the row's
 *     storage, the old block is freed;  negatives: try_read (:1257), a failed lock redraws;
 *   - gradient: Array1::zeros(dim) (:1200), each `gradient = (&y - &y_i) * c` (:1230,1290)
 *     allocates the difference, scales it in place and drops the previous gradient;
 *   - edges: Vec<(NodeIdx, OutEdge<f32>)>, 24 bytes per edge (:939,975-984);  the rejection scan
 *     get_edge (nodeparam.rs:83-85) walks the node's own Vec<OutEdge<f32>> (16 bytes per edge, one
 *     heap block per node, kdumap.rs:63-87).
 * The lock is parking_lot's word lock restated (uncontended paths only: one CAS to lock, one
 * atomic to unlock; contention spins with sched_yield).  The generator stays xoshiro256++ (the
 * reference's thread-local ChaCha12 costs more per draw): the twin is still an optimistic
 * stand-in for the Rust loop, by less than the plain-array loop is.
 * Checked once with gcc -fsanitize=address,undefined (4 threads, both samplers: clean, no leak) and -fsanitize=thread
 * (no report between two worker threads; the remaining ones pair the set-up writes of the main thread with libgomp's
 * uninstrumented barrier).
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t node; float weight; } rl_outedge_t;              /* OutEdge<f32> */
typedef struct { uint64_t src; rl_outedge_t e; } rl_edge_t;                /* (NodeIdx, OutEdge<f32>) */
typedef struct { rl_outedge_t *edges; uint64_t len, cap; float scale; } rl_nodeparam_t;   /* NodeParam, nodeparam.rs:72-76 */
typedef struct {
    atomic_long strong, weak;             /* ArcInner */
    atomic_ulong lock;                    /* RawRwLock state */
    float *data; uint64_t len, cap;       /* Array1<F>: OwnedRepr<F> ... */
    float *ptr; uint64_t dim; int64_t stride; /* ... + view pointer, shape, stride */
} rl_row_t;
#define RL_WRITER 8ul
#define RL_READER 16ul

static inline int rl_try_read(rl_row_t *r)
{
    unsigned long st = atomic_load_explicit(&r->lock, memory_order_relaxed);
    while (!(st & RL_WRITER))
        if (atomic_compare_exchange_weak_explicit(&r->lock, &st, st + RL_READER, memory_order_acquire, memory_order_relaxed)) return 1;
    return 0;
}
static inline void rl_read_unlock(rl_row_t *r) { atomic_fetch_sub_explicit(&r->lock, RL_READER, memory_order_release); }
static inline void rl_write_lock(rl_row_t *r)
{
    for (;;) {
        unsigned long zero = 0ul;
        if (atomic_compare_exchange_weak_explicit(&r->lock, &zero, RL_WRITER, memory_order_acquire, memory_order_relaxed)) return;
        if (zero != 0ul) sched_yield();
    }
}
static inline void rl_write_unlock(rl_row_t *r) { atomic_store_explicit(&r->lock, 0ul, memory_order_release); }
static inline rl_row_t *rl_get_embedded_data(rl_row_t **rows, uint64_t node)          /* :1071-1073 */
{
    rl_row_t *r = rows[node];
    atomic_fetch_add_explicit(&r->strong, 1, memory_order_relaxed);
    return r;
}
static inline void rl_drop(rl_row_t *r) { atomic_fetch_sub_explicit(&r->strong, 1, memory_order_release); }
static inline float *rl_to_owned(const rl_row_t *r, uint32_t d)
{
    float *c = (float *)malloc(d * sizeof(float));
    memcpy(c, r->ptr, d * sizeof(float));
    return c;
}
static inline float *rl_read_owned(rl_row_t **rows, uint64_t node, uint32_t d)        /* get(..).read().to_owned() */
{
    rl_row_t *r = rl_get_embedded_data(rows, node);
    while (!rl_try_read(r)) sched_yield();
    float *c = rl_to_owned(r, d);
    rl_read_unlock(r);
    rl_drop(r);
    return c;
}
static inline void rl_write_move(rl_row_t **rows, uint64_t node, float *y)            /* *(get(..).write()) = y */
{
    rl_row_t *r = rl_get_embedded_data(rows, node);
    rl_write_lock(r);
    float *old = r->data;
    r->data = y; r->ptr = y;
    rl_write_unlock(r);
    rl_drop(r);
    free(old);
}
static inline int rl_get_edge(const rl_nodeparam_t *np, uint64_t k)                   /* nodeparam.rs:83-85 */
{
    for (uint64_t m = 0; m < np->len; m++) if (np->edges[m].node == k) return 1;
    return 0;
}

int64_t oracle_optimize_reference_layout(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col,
                                         const float *p, const float *emb_scale, float *y, double b,
                                         double grad_step_init, uint32_t nb_sampling_by_edge, uint32_t nb_grad_batch,
                                         const float *neg_w, uint64_t seed, uint32_t first_batch,
                                         uint32_t n_batches_to_run, double sample_fraction, int n_threads,
                                         double *loop_seconds)
{
    if (d > 64) return -1;
    const uint64_t E = row_ptr[n];
    rl_edge_t *edges = (rl_edge_t *)malloc(E * sizeof(rl_edge_t));                    /* :975-984 */
    rl_nodeparam_t *node_params = (rl_nodeparam_t *)malloc(n * sizeof(rl_nodeparam_t));
    rl_row_t **rows = (rl_row_t **)malloc(n * sizeof(rl_row_t *));
    if (!edges || !node_params || !rows) return -2;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t k = row_ptr[i + 1] - row_ptr[i];
        node_params[i].edges = (rl_outedge_t *)malloc((k ? k : 1) * sizeof(rl_outedge_t));
        node_params[i].len = node_params[i].cap = k; node_params[i].scale = 0.0f;
        rl_row_t *r = (rl_row_t *)malloc(sizeof(rl_row_t));
        float *data = (float *)malloc(d * sizeof(float));
        if (!node_params[i].edges || !r || !data) return -2;
        for (uint64_t m = 0; m < k; m++) {
            const uint64_t e = row_ptr[i] + m;
            node_params[i].edges[m].node = col[e]; node_params[i].edges[m].weight = p[e];
            edges[e].src = i; edges[e].e = node_params[i].edges[m];
        }
        atomic_init(&r->strong, 1); atomic_init(&r->weak, 1); atomic_init(&r->lock, 0ul);
        memcpy(data, y + i * d, d * sizeof(float));
        r->data = r->ptr = data; r->len = r->cap = r->dim = d; r->stride = 1;
        rows[i] = r;
    }
    alias_t pos, negt;
    if (alias_build(&pos, p, E)) return -2;                 /* :987 */
    const int have_neg = neg_w != NULL;
    if (have_neg && alias_build(&negt, neg_w, n)) return -2;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    const uint64_t nb_sample = (uint64_t)((double)nb_sampling_by_edge * (double)E * sample_fraction); /* :858 */
    int64_t done = 0;
    double t_loop = 0.0;
    for (uint32_t iter = first_batch; iter < first_batch + n_batches_to_run && iter <= nb_grad_batch; iter++) {
        const double grad_step = grad_step_init * (1.0 - (double)iter / (double)nb_grad_batch); /* :875 */
        const double t0 = wall_seconds();
#pragma omp parallel
        {
            rng_t rng;
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            rng_seed(&rng, seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)iter << 20) + (uint64_t)tid);
#pragma omp for schedule(static)
            for (int64_t s = 0; s < (int64_t)nb_sample; s++) {
                const uint64_t e = alias_draw(&pos, &rng);                           /* :1182 */
                const uint64_t node_i = edges[e].src, node_j = edges[e].e.node;
                float *y_i = rl_read_owned(rows, node_i, d);                         /* :1186 */
                float *y_j = rl_read_owned(rows, node_j, d);                         /* :1187 */
                float *gradient = (float *)calloc(d, sizeof(float));                 /* :1200 */
                const double weight = (double)edges[e].e.weight;
                const double scale = (double)emb_scale[node_i];
                const double s2 = scale * scale;
                float dsum = 0.0f;
                for (uint32_t c = 0; c < d; c++) dsum += (y_i[c] - y_j[c]) * (y_i[c] - y_j[c]);
                const double u = (double)dsum / s2;
                const double coeff = common_coeff(u, s2, b);
                if (u > 0.0) {
                    const double alfa = (double)(1.0f / PROBA_MIN);
                    const double rep = 1.0 / fmax(u * u, alfa);
                    double cij = grad_step * coeff * (-weight + (1.0 - weight) * rep);
                    cij = fmax(cij, -0.49);
                    const float cf = (float)cij;
                    float *t = (float *)malloc(d * sizeof(float));                   /* &y_j - &y_i */
                    for (uint32_t c = 0; c < d; c++) t[c] = y_j[c] - y_i[c];
                    for (uint32_t c = 0; c < d; c++) t[c] = t[c] * cf;               /* * coeff, in place */
                    free(gradient); gradient = t;
                }
                for (uint32_t c = 0; c < d; c++) y_i[c] -= gradient[c];              /* :1237 */
                for (uint32_t c = 0; c < d; c++) y_j[c] += gradient[c];              /* :1238 */
                rl_write_move(rows, node_j, y_j);                                    /* :1239 */
                int got = 0;
                while (got < 5) {                                                    /* :1241-1299 */
                    const uint64_t k = have_neg ? alias_draw(&negt, &rng) : rng_below(&rng, n);
                    if (k == node_i || k == node_j || rl_get_edge(&node_params[node_i], k)) continue;
                    rl_row_t *r = rl_get_embedded_data(rows, k);
                    if (!rl_try_read(r)) { rl_drop(r); continue; }                   /* :1257-1265 */
                    float *y_k = rl_to_owned(r, d);
                    rl_read_unlock(r);
                    rl_drop(r);
                    got++;
                    float dk = 0.0f;
                    for (uint32_t c = 0; c < d; c++) dk += (y_i[c] - y_k[c]) * (y_i[c] - y_k[c]);
                    const double dik = (double)dk;
                    const double uk = dik / s2;
                    const double ck = common_coeff(uk, s2, b);
                    if (dik > 0.0) {
                        const double rep = 1.0 / fmax(uk * uk, 1.0 / 16.0);
                        const double cik = fmin(grad_step * ck * rep, 2.0);
                        const float cf = (float)cik;
                        float *t = (float *)malloc(d * sizeof(float));
                        for (uint32_t c = 0; c < d; c++) t[c] = y_k[c] - y_i[c];
                        for (uint32_t c = 0; c < d; c++) t[c] = t[c] * cf;
                        free(gradient); gradient = t;
                    }
                    for (uint32_t c = 0; c < d; c++) y_i[c] -= gradient[c];          /* :1297 */
                    free(y_k);
                }
                rl_write_move(rows, node_i, y_i);                                    /* :1301 */
                free(gradient);
            }
        }
        t_loop += wall_seconds() - t0;
        done += (int64_t)nb_sample;
    }
    if (loop_seconds) *loop_seconds = t_loop;
    for (uint64_t i = 0; i < n; i++) {                                               /* :888-899 copy-out */
        memcpy(y + i * d, rows[i]->ptr, d * sizeof(float));
        free(rows[i]->data); free(rows[i]); free(node_params[i].edges);
    }
    alias_free(&pos);
    if (have_neg) alias_free(&negt);
    free(rows); free(node_params); free(edges);
    return done;
}

/* fromhnsw/hubness.rs:39-76: in-degree counts; embedder.rs:826-833 clamps to [1, N] as f32 */
void oracle_hubness_weights(uint64_t n, const uint64_t *row_ptr, const uint32_t *col, float *w)
{
    uint32_t *cnt = (uint32_t *)calloc(n, sizeof(uint32_t));
    for (uint64_t m = 0; m < row_ptr[n]; m++) cnt[col[m]]++;
    for (uint64_t i = 0; i < n; i++) w[i] = fminf(fmaxf((float)cnt[i], 1.0f), (float)n);
    free(cnt);
}

/* ------------------------------------------------------------------------------------------
 * Quality pieces.  embedder.rs:478-522 get_transformed_kgraph: per node, embedded L2 distance to
 * each original neighbour stored as a RUNNING MINIMUM in graph order (:500-509), then sorted.
 * out: E floats (row_ptr layout), each row ascending.
 * ------------------------------------------------------------------------------------------ */
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }
void oracle_transformed_kgraph(uint64_t n, uint32_t d, const uint64_t *row_ptr, const uint32_t *col,
                               const float *y, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        float run = FLT_MAX;
        for (uint64_t m = row_ptr[i]; m < row_ptr[i + 1]; m++) {
            const float *yi = y + (uint64_t)i * d, *yj = y + (uint64_t)col[m] * d;
            float ds = 0.0f;
            for (uint32_t c = 0; c < d; c++) ds += (yi[c] - yj[c]) * (yi[c] - yj[c]);
            float dd = sqrtf(ds);
            run = fminf(dd, run);
            out[m] = run;
        }
        qsort(out + row_ptr[i], row_ptr[i + 1] - row_ptr[i], sizeof(float), cmp_float);
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
