// C ABI of libsegp.so (include/segp.h): handle management, chunked rollout driver, host-buffer entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <chrono>
#include <vector>

#include "segp_internal.cuh"

namespace segp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
static int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return SEGP_OK;
    SEGP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return SEGP_OK;
}
template <typename T>
static void dev_free(T*& p) {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
}

// The handle-free entry points (scoring, arg-best, stand-alone ellipsoid step, safety distance) take their few hundred
// bytes of device-side parameters from the stream-ordered allocator.  With the default release threshold of 0 the
// pool gives everything back to the driver at the next synchronisation and the following cudaMallocAsync pays for a
// fresh reservation: measured 2-6 ms per call inside a sampling-MPC iteration whose kernels take 0.5 ms.  Let the
// device's default pool keep up to 64 MB (once per device).
static void retain_small_async_allocations() {
    static bool seen[64] = {};
    if (!first_call_on_device(seen)) return;
    int dev = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    uint64_t cur = 0, want = 64ull << 20;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < want)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
    cudaGetLastError();
}

}  // namespace segp

using namespace segp;


namespace segp {
constexpr int META_N = 32;      // doubles at the head of the factor arena: decisions of the factorising rank
constexpr int PROBE_N = 1024;   // probe inputs of the factorize-time calibration of the int8 error model
// probe statistics (read through segp_get_param "probe_*"); worst case over the output dimensions
enum {
    PS_RAN = 0,   // 1 if the probe ran
    PS_FRAC4,     // fraction of probe inputs the guard flags on the 10-product set
    PS_FRAC5,     // ... on the 15-product set
    PS_ERR4,      // max |error of |v|^2| of the 10-product set against float64
    PS_ERR5,
    PS_REL4,      // max error relative to sigma^2
    PS_REL5,
    PS_RATIO4,    // max |error| / (predicted 1-sigma): how well the error model holds (before calibration)
    PS_RATIO5,
    PS_RHO4,      // calibration factor applied to the model's standard deviation (>= 1)
    PS_RHO5,
    PS_MINVAR,    // smallest sigma^2 / k** over the probe inputs
    PS_MARGIN4,   // largest (kappa x estimated error) / (rtol sigma^2) of the 10-product set over the probes (1 = at the guard)
    PROBE_STATS
};
}  // namespace segp

namespace segp {
struct GraphEntry;
}

struct segp_model {
    int device = 0;
    int n_s = 0, n_in = 0, n_u = 0, dim = 0;
    int kern[SEGP_MAX_NS] = {0};
    // model data (host copies kept for re-factorisation / introspection)
    int n_train = 0, n_pad = 0, nblk = 0;
    long ntri = 0;
    std::vector<double> h_x, h_y, h_ls, h_var, h_noise;
    bool has_data = false, factorized = false;
    // device model
    double* xs = nullptr;      // [n_s][n_pad][dim] inputs scaled by 1/lengthscale
    double* yp = nullptr;      // [n_s][n_pad] targets, zero padded
    double* invls = nullptr;   // [n_s][dim]
    double* var = nullptr;     // [n_s]
    // factorised state: ONE device allocation (`arena`), so a multi-GPU setup broadcasts it with one collective
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0;
    double* meta = nullptr;    // [META_N] decisions of the factorising rank (digit set, float64 fallback, ...)
    double* beta = nullptr;    // [n_s][n_pad]
    double* logdet = nullptr;  // [n_s]
    int8_t* wi8 = nullptr;     // [n_s][nblk (nblk+1)][I8_S][I8_A_TILE] digit planes of W, classic set (15 products)
    double* rowfac = nullptr;  // [n_s][n_pad] per-row factors of the classic set
    int8_t* wi8s = nullptr;    // [n_s][nblk (nblk+1)][I8_SS][I8_A_TILE] diagonal-split set (10 products)
    int8_t* wm1 = nullptr;     // [n_s][nblk][2][I8_A_TILE] leading digit of the diagonal (split set)
    double* rowfac_s = nullptr;
    float* werr5 = nullptr;    // [n_s][n_pad] error-model variance weights of the two sets (calibrated by the probe)
    float* werr4 = nullptr;
    // float64 DMMA operand: only kept when the float64 contraction is (or may be) used
    double* wt = nullptr;      // [n_s][ntri][128*128]
    // composite (linear x stationary + linear) kernels: segp_set_linear_terms
    bool has_composite = false, has_linear_terms = false;
    std::vector<double> h_plin, h_lin;
    double* xraw = nullptr;    // [n_pad][dim] unscaled inputs
    double* plin = nullptr;    // [n_s][dim]
    double* lin = nullptr;     // [n_s][dim]
    double* xtb = nullptr;     // [n_s][dim] X^T beta_d
    double* xmax = nullptr;    // [dim] max_i |x_ij| (scale bound of the composite kernels on the digit-plane path)
    double* colfac2 = nullptr; // workspace: [n_s][b_cap] squared per-trajectory scales (composite models, int8 path)
    double* jac2_part = nullptr;   // workspace: additive Jacobian partials
    double* kss = nullptr;         // workspace: [n_s][b_cap] prior variances
    double* wdense = nullptr;  // [n_s][n_pad][n_pad] W = L^-1 kept dense for segp_append (only if opt_keep_w)
    long opt_keep_w = 0;       // keep wdense after factorising (set by the first segp_append)
    long opt_fact_i8 = -1;     // factorisation GEMMs on tcgen05 digit planes: -1 automatic (n_pad >= 1024), 0 off, 1 on
    bool last_fact_i8 = false; // the last segp_factorize ran them there
    // scratch of segp_factorize kept between calls (an MPC loop refits the same-size model every step; allocating and
    // freeing ~3 GB costs as much as the factorisation itself): up to 4 slots; option "scratch_cache" -1 = keep when
    // it is <= 4 GB, 0 = never, 1 = always
    struct FactScratch {
        double *kbuf = nullptr, *wbuf = nullptr, *tmp = nullptr, *diag_inv = nullptr, *u_tmp = nullptr;
        FdScratch fd{nullptr, nullptr, nullptr, nullptr};
    };
    FactScratch fact_cache[4];
    int fact_cache_npad = 0, fact_cache_slots = 0;
    bool fact_cache_fd = false, fact_cache_w = false;
    long opt_scratch_cache = -1;
    bool last_append_incremental = false;
    // precision management of the int8 path
    int i8_primary = 5;        // digit set of the first contraction pass: 4 (10 products) or 5 (15 products)
    bool auto_fp64 = false;    // the probe found even the 15-product set short of guard_rtol: automatic mode runs float64
    long opt_i8_digits = 0;    // 0 = automatic (probe), 4 / 5 = forced
    long opt_guard = 1;        // run the a-posteriori precision guard (flag + 15-product recomputation of flagged panels)
    long opt_probe = 1;        // calibrate the error model at factorize time (PROBE_N random inputs against float64)
    long opt_keep_fp64 = 0;    // keep the float64 operand after factorising even when automatic mode does not need it
    double guard_rtol = 1e-4, guard_kappa = 5.0;
    double probe_stat[PROBE_STATS] = {0};
    int force_mode = -2, force_digits = 0;   // probe only: overrides of the kernel selection
    unsigned int* fallback_counter = nullptr;   // device: panels recomputed on the 15-product set so far
    unsigned int* h_fallback = nullptr;         // pinned host mirror, refreshed (asynchronously) after every rollout call
    unsigned int fallback_seen = 0;             // value of the mirror when the previous call was issued
    long panels_prev_call = 0;                  // panel contractions the previous guarded call issued
    bool demoted = false;                       // automatic mode fell back to the 15-product first pass at run time
    bool unguarded = false;                     // the probe found the 10-product set at least 4x inside the tolerance
                                                // everywhere: no guard / recomputation launches (status flags stay)
    // workspace
    long b_cap = 0;
    int nsplit = 1, blocks_per_split = 1;
    double* ks = nullptr;      // fp64 K* block (DMMA path)
    int8_t* ki8 = nullptr;     // digit planes of the K* block (tcgen05 path)
    float* epart = nullptr;    // [n_s][nblk][b_cap] error-model partials (tcgen05 path)
    int32_t* pflag = nullptr;  // [npanel_cap] panels to recompute on the 15-product set
    long npanel_cap = 0;
    int ws_mode = 0;           // layout of the current workspace: 0 = fp64 ks, 1 = ki8
    double* mu_part = nullptr;
    double* jac_part = nullptr;
    double* qpart = nullptr;
    StepParams* d_sp = nullptr;
    size_t workspace_bytes = 0;
    // host-entry staging (grown on demand)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t s_host = nullptr;   // stream of the host entry points
    cudaStream_t s_cap = nullptr;    // capture stream of the CUDA-graph path
    cudaStream_t s_copy = nullptr;   // device-to-host result copies of the host entry points
    cudaStream_t s_sub[3] = {nullptr, nullptr, nullptr};   // sub-batch chains of small models (segp_multistep)
    cudaEvent_t ev_sub[3] = {nullptr, nullptr, nullptr}, ev_sub_fork = nullptr;
    long opt_substreams = -1;        // -1 automatic (2 for models of <= 8 block rows), 0 off, n = at most n (<= 4)
    cudaEvent_t ev_chunk = nullptr;
    // options
    long opt_chunk = 8192;
    long opt_panel_group = 16;
    long opt_ksplit = 0;   // 0 = automatic
    long opt_i8_panel_group = 0;   // tri_i8m / tri_i8mp: panels per L2 group (even), 0 = automatic
    long opt_i8_cluster = 2;       // tri_i8m: CTAs per cluster sharing one W stage by multicast (2 or 4)
    long opt_tri_mode = -1;   // -1 = automatic (int8 tcgen05 when n_pad <= I8_MAX_NPAD, the kernels are not composite and
                              // the probe did not ask for float64; else 0), 0 = fp64 DMMA, 1 = int8 reference kernel (one
                              // CTA per tile, classic set; test cross-check), 4 = tri_i8m (single-CTA MMAs over two K*
                              // planes at once, W multicast over a CTA pair), 5 = tri_i8mp (the same, persistent)
    long long* i8_prof = nullptr;   // profiling only: [128][8] counters of the persistent kernel's MMA threads
    bool last_tri_persistent = false;   // which tcgen05 kernel the last contraction launch used (automatic mode)
    int last_tri_digits = 0;
    long launches = 0;
    // optional per-launch timing of the contraction (bench.py roofline): event pairs recorded on the launching stream
    bool time_tri = false;
    std::vector<cudaEvent_t> tri_events;
    size_t tri_events_used = 0;
    // CUDA-graph cache of the rollout schedule (segp_multistep, option "graph")
    long opt_graph = 1;
    segp::GraphEntry* graphs = nullptr;
};

namespace segp {

struct GraphEntry {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec = nullptr;
    long launches = 0, events = 0;
    GraphEntry* next = nullptr;
};

static void free_graphs(segp_model* m) {
    while (m->graphs != nullptr) {
        GraphEntry* g = m->graphs;
        m->graphs = g->next;
        if (g->exec != nullptr) cudaGraphExecDestroy(g->exec);
        delete g;
    }
}

static void free_model_buffers(segp_model* m) {
    free_graphs(m);
    dev_free(m->xs);
    dev_free(m->yp);
    dev_free(m->invls);
    dev_free(m->var);
    dev_free(m->arena);
    m->arena_bytes = 0;
    m->meta = m->beta = m->logdet = m->rowfac = m->rowfac_s = nullptr;
    m->wi8 = m->wi8s = m->wm1 = nullptr;
    m->werr5 = m->werr4 = nullptr;
    dev_free(m->wt);
    dev_free(m->wdense);
    dev_free(m->xraw);
    dev_free(m->plin);
    dev_free(m->lin);
    dev_free(m->xtb);
    dev_free(m->xmax);
    m->has_linear_terms = false;
    m->factorized = false;
}

static void free_fact_cache(segp_model* m) {
    for (auto& c : m->fact_cache) {
        dev_free(c.kbuf);
        dev_free(c.wbuf);
        dev_free(c.tmp);
        dev_free(c.diag_inv);
        dev_free(c.u_tmp);
        dev_free(c.fd.ap);
        dev_free(c.fd.bp);
        dev_free(c.fd.as);
        dev_free(c.fd.bs);
    }
    m->fact_cache_npad = m->fact_cache_slots = 0;
}

static void free_workspace(segp_model* m) {
    free_graphs(m);
    dev_free(m->ks);
    dev_free(m->ki8);
    dev_free(m->epart);
    dev_free(m->pflag);
    m->npanel_cap = 0;
    dev_free(m->mu_part);
    dev_free(m->jac_part);
    dev_free(m->qpart);
    dev_free(m->jac2_part);
    dev_free(m->kss);
    dev_free(m->colfac2);
    m->b_cap = 0;
    m->workspace_bytes = 0;
}

// the int32 accumulators of the digit-plane products bound the (padded) training size; composite kernels run on the
// digit planes too, with a per-trajectory scale (KstarI8Args::colfac2)
static bool i8_capable(const segp_model* m) { return m->n_pad <= I8_MAX_NPAD; }
// which kernel runs the variance contraction: 0 fp64 DMMA, 1 int8 reference kernel, 4 tri_i8m, 5 tri_i8mp
static int tri_mode(const segp_model* m) {
    if (m->force_mode > -2) return m->force_mode == -1 ? 4 : m->force_mode;
    if (m->opt_tri_mode >= 0) return (int)m->opt_tri_mode;
    return (i8_capable(m) && !m->auto_fp64) ? 4 : 0;
}
// digit set of the first contraction pass (int8 modes 4 / 5; the reference kernel knows the classic set only)
static int tri_digits(const segp_model* m) {
    if (m->force_digits != 0) return m->force_digits;
    if (tri_mode(m) == 1) return 5;
    if (m->opt_i8_digits != 0) return (int)m->opt_i8_digits;
    return m->demoted ? 5 : m->i8_primary;
}
static double guard_gs(const segp_model* m) {
    const double r = 2.0 * m->guard_kappa / m->guard_rtol;
    return r * r;
}

static int ensure_workspace(segp_model* m, long n_batch) {
    constexpr long ALIGN = 384;   // lcm of the DMMA tile (128 columns) and the tcgen05 panel (96 columns)
    long want = std::min<long>(m->opt_chunk, n_batch);
    want = (want + ALIGN - 1) / ALIGN * ALIGN;
    const int mode = tri_mode(m);
    const bool i8 = mode != 0;
    if (i8 && m->wi8 == nullptr) {
        set_error("tri_mode=%d (int8 tcgen05) needs n_train_padded <= %ld; this model has %d", mode, I8_MAX_NPAD,
                  m->n_pad);
        return SEGP_ERR_UNSUPPORTED;
    }
    if (!i8 && m->wt == nullptr) {
        set_error("internal: the float64 operand is not resident");
        return SEGP_ERR_INVALID;
    }
    if ((i8 ? 1 : 0) != m->ws_mode && m->b_cap > 0) free_workspace(m);
    // split of the N-length reductions of kstar_mean_jac over blockIdx.z so small batches still fill 148 SMs
    const long col_blocks = want / TILE;
    int nsplit;
    if (m->opt_ksplit > 0) {
        nsplit = (int)std::min<long>(m->opt_ksplit, m->nblk);
    } else {
        const long target = 16 * 148;   // ~48 warps per SM for the (latency-bound, fp64) K* kernels
        nsplit = (int)std::max<long>(1, std::min<long>(m->nblk, (target + col_blocks * m->n_s - 1) / (col_blocks * m->n_s)));
    }
    const int bps = (m->nblk + nsplit - 1) / nsplit;
    nsplit = (m->nblk + bps - 1) / bps;
    if (want <= m->b_cap && nsplit == m->nsplit && bps == m->blocks_per_split) return SEGP_OK;
    if (want <= m->b_cap) {
        // same capacity, only the split changed: partial buffers are sized for nblk splits, nothing to do
        free_graphs(m);
        m->nsplit = nsplit;
        m->blocks_per_split = bps;
        return SEGP_OK;
    }
    free_workspace(m);
    const size_t n_ks = i8 ? 0 : (size_t)m->n_s * m->n_pad * want;
    const long npanel_cap = want / I8_N;
    const size_t n_ki8 = i8 ? (size_t)m->n_s * npanel_cap * (m->n_pad / I8_KB) * (I8_S * I8_B_TILE) : 0;
    const size_t n_mu = (size_t)m->nblk * m->n_s * want;
    const size_t n_jac = n_mu * m->dim;
    const size_t n_q = (size_t)m->n_s * m->nblk * want;
    SEGP_CHECK(dev_alloc(&m->ks, n_ks));
    SEGP_CHECK(dev_alloc(&m->ki8, n_ki8));
    SEGP_CHECK(dev_alloc(&m->mu_part, n_mu));
    SEGP_CHECK(dev_alloc(&m->jac_part, n_jac));
    SEGP_CHECK(dev_alloc(&m->qpart, n_q));
    if (i8) {
        SEGP_CHECK(dev_alloc(&m->epart, n_q));
        SEGP_CHECK(dev_alloc(&m->pflag, (size_t)npanel_cap));
        SEGP_CUDA_CHECK(cudaMemset(m->epart, 0, n_q * sizeof(float)));
        SEGP_CUDA_CHECK(cudaMemset(m->pflag, 0, npanel_cap * sizeof(int32_t)));
    }
    if (m->has_composite) {
        SEGP_CHECK(dev_alloc(&m->jac2_part, n_jac));
        SEGP_CHECK(dev_alloc(&m->kss, (size_t)m->n_s * want));
        if (i8) SEGP_CHECK(dev_alloc(&m->colfac2, (size_t)m->n_s * want));
    }
    if (n_ks > 0) SEGP_CUDA_CHECK(cudaMemset(m->ks, 0, n_ks * sizeof(double)));
    if (n_ki8 > 0) SEGP_CUDA_CHECK(cudaMemset(m->ki8, 0, n_ki8));
    m->workspace_bytes = (n_ks + n_mu + n_jac + n_q) * sizeof(double) + n_ki8 + (i8 ? n_q * sizeof(float) : 0);
    m->b_cap = want;
    m->npanel_cap = npanel_cap;
    m->ws_mode = i8 ? 1 : 0;
    m->nsplit = nsplit;
    m->blocks_per_split = bps;
    return SEGP_OK;
}

static int fill_step_params(StepParams* sp, const segp_reach_params* prm, int n_s, int n_in, int n_u) {
    memset(sp, 0, sizeof(*sp));
    if (prm == nullptr || prm->h_l_mu == nullptr || prm->h_l_sigma == nullptr) {
        set_error("reach params: l_mu and l_sigma are required");
        return SEGP_ERR_INVALID;
    }
    for (int i = 0; i < n_s; ++i) {
        sp->l_mu[i] = prm->h_l_mu[i];
        sp->l_sigma[i] = prm->h_l_sigma[i];
    }
    sp->c_safety = prm->c_safety;
    for (int i = 0; i < n_s; ++i)
        for (int j = 0; j < n_s; ++j) sp->a[i * n_s + j] = prm->h_a ? prm->h_a[i * n_s + j] : (i == j ? 1.0 : 0.0);
    for (int i = 0; i < n_s; ++i)
        for (int j = 0; j < n_u; ++j) sp->b[i * n_u + j] = prm->h_b ? prm->h_b[i * n_u + j] : 0.0;
    if (prm->propagation < SEGP_PROP_ELLIPSOID || prm->propagation > SEGP_PROP_MEAN_EQUIVALENT) {
        set_error("reach params: unknown propagation mode %d", prm->propagation);
        return SEGP_ERR_INVALID;
    }
    sp->prop_mode = prm->propagation;
    sp->has_t = prm->h_t_z_gp != nullptr;
    if (sp->has_t) {
        for (int i = 0; i < n_in * n_s; ++i) sp->t[i] = prm->h_t_z_gp[i];
    } else if (n_in != n_s) {
        set_error("GP input state dimension %d differs from the state dimension %d: t_z_gp is required", n_in, n_s);
        return SEGP_ERR_INVALID;
    }
    return SEGP_OK;
}

static int check_ready(const segp_model* m) {
    if (m == nullptr) {
        set_error("null model handle");
        return SEGP_ERR_INVALID;
    }
    if (!m->factorized) {
        set_error("model is not factorised (call segp_set_model + segp_factorize first)");
        return SEGP_ERR_NOT_TRAINED;
    }
    return SEGP_OK;
}

static KstarArgs base_kstar_args(const segp_model* m) {
    KstarArgs k{};
    k.xs = m->xs;
    k.invls = m->invls;
    k.var = m->var;
    k.beta = m->beta;
    for (int d = 0; d < m->n_s; ++d) k.kern[d] = m->kern[d];
    k.n_train = m->n_train;
    k.n_pad = m->n_pad;
    k.dim = m->dim;
    k.n_in = m->n_in;
    k.n_u = m->n_u;
    k.n_s_state = m->n_s;
    k.b_cap = m->b_cap;
    k.groups_per_split = m->blocks_per_split * (TILE / 4);
    k.ks = m->ks;
    k.mu_part = m->mu_part;
    k.jac_part = m->jac_part;
    if (m->has_composite) {
        k.xraw = m->xraw;
        k.plin = m->plin;
        k.lin = m->lin;
        k.xtb = m->xtb;
        k.jac2_part = m->jac2_part;
        k.kss = m->kss;
    }
    return k;
}

// K* block (+ mean / Jacobian partials) in the operand format of the active contraction kernel
static int run_kstar(segp_model* m, const KstarArgs& k, cudaStream_t st, int panel0 = 0) {
    if (m->ws_mode != 0) {
        KstarI8Args k8{};
        k8.k = k;
        k8.ki8 = m->ki8;
        k8.npanel_cap = m->npanel_cap;
        k8.colfac2 = m->colfac2;
        k8.xmax = m->xmax;
        k8.panel0 = panel0;
        return launch_kstar_i8(k8, m->n_s, m->nsplit, st);
    }
    return launch_kstar(k, m->n_s, m->nsplit, st);
}

// argument block of an int8 contraction launch on digit set `digits`
static TriI8Args tri_i8_args(const segp_model* m, long nb, int panel0, int digits) {
    TriI8Args t{};
    const bool split = digits == 4;
    t.wi8 = split ? m->wi8s : m->wi8;
    t.rowfac = split ? m->rowfac_s : m->rowfac;
    t.wm1 = m->wm1;
    t.werr = split ? m->werr4 : m->werr5;
    t.colfac2 = m->colfac2;
    t.digits = digits;
    t.ki8 = m->ki8;
    t.qpart = m->qpart;
    t.epart = m->epart;
    t.nblk = m->nblk;
    t.npanels = (int)((nb + I8_N - 1) / I8_N);
    t.panel0 = panel0;   // != 0 for the sub-batch chains of small models
    // panels per L2 group of tri_i8m / tri_i8mp (0 = 24): a sweep over 8..24 at C4 and C5 moved the launch time by
    // less than 0.5 %, so there is no automatic choice
    t.pgroup = (int)m->opt_i8_panel_group;
    t.cluster = (int)m->opt_i8_cluster;
    t.npanel_cap = m->npanel_cap;
    t.b_cap = m->b_cap;
    t.fix_bi = -1;
    t.prof = m->i8_prof;
    return t;
}

// Variance contraction of the trajectories [panel0 * 96, nb) of the current chunk: tri_sumsq on the DMMA pipe, or the
// tcgen05 kernels on the selected digit set followed -- for the 10-product set -- by the precision guard and the
// recomputation of the flagged panels on the 15-product set.  Optionally bracketed by a CUDA-event pair on the
// launching stream (bench.py: the pair spans the first pass only, the kernel the roofline is quoted for).
static int run_tri(segp_model* m, long nb, cudaStream_t st, int panel0 = 0) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (m->time_tri) {
        while (m->tri_events.size() < m->tri_events_used + 2) {
            cudaEvent_t e;
            SEGP_CUDA_CHECK(cudaEventCreate(&e));
            m->tri_events.push_back(e);
        }
        e0 = m->tri_events[m->tri_events_used];
        e1 = m->tri_events[m->tri_events_used + 1];
        m->tri_events_used += 2;
        SEGP_CUDA_CHECK(cudaEventRecord(e0, st));
    }
    if (m->ws_mode != 0) {
        const int mode = tri_mode(m);
        const int digits = tri_digits(m);
        TriI8Args t = tri_i8_args(m, nb, panel0, digits);
        // Automatic mode: the persistent folded kernel where per-tile overhead matters (short tiles, enough of them to
        // balance a static schedule: +5 % at C3), the one-cluster-per-tile kernel otherwise (C4: power-capped, +1 %
        // at best; C5: the persistent order runs 13 % slower; C2: too few tiles) -- profiles/round1/persistent_tri_i8mp.txt
        bool persistent = mode == 5;
        if (mode == 4 && m->opt_tri_mode < 0 && m->force_mode == -2 && t.panel0 == 0) {
            const long ntiles = (long)m->n_s * ((m->nblk + 1) / 2) * ((t.npanels + 1) / 2);
            persistent = m->nblk <= 32 && ntiles >= 4 * 74;
        }
        m->last_tri_persistent = persistent;
        m->last_tri_digits = digits;
        if (mode == 1) {
            t.epart = nullptr;   // the reference kernel has no error-model output: clear it so no stale estimate is read
            SEGP_CUDA_CHECK(cudaMemsetAsync(m->epart, 0, (size_t)m->n_s * m->nblk * m->b_cap * sizeof(float), st));
            SEGP_CHECK(launch_tri_i8(t, m->n_s, st));
        } else {
            SEGP_CHECK(persistent ? launch_tri_i8mp(t, m->n_s, st) : launch_tri_i8m(t, m->n_s, st));
        }
        if (e1 != nullptr) SEGP_CUDA_CHECK(cudaEventRecord(e1, st));
        if (digits == 4 && m->opt_guard != 0 && m->force_digits == 0 && !(m->unguarded && m->opt_i8_digits == 0)) {
            GuardArgs g{};
            g.qpart = m->qpart;
            g.epart = m->epart;
            g.gp_var = m->var;
            g.kss = m->has_composite ? m->kss : nullptr;
            g.nblk = m->nblk;
            g.n_s = m->n_s;
            g.panel0 = panel0;
            g.b_cap = m->b_cap;
            g.n_batch = nb;
            g.gs = guard_gs(m);
            g.pflag = m->pflag;
            g.counter = m->fallback_counter;
            SEGP_CHECK(launch_i8_guard(g, st));
            // recomputation on the 15-product set as a RESIDENT grid (one cluster per TPC walking the tile list and
            // skipping unflagged panel pairs): when nothing is flagged it costs one wave of CTAs that exit at once,
            // not n_s x nblk x npanels launches of empty 220 KB CTAs
            TriI8Args t5 = tri_i8_args(m, nb, panel0, 5);
            t5.pflag = m->pflag;
            SEGP_CHECK(launch_tri_i8mp(t5, m->n_s, st));
            m->launches += 2;
        }
        return SEGP_OK;
    }
    TriArgs t{};
    t.wt = m->wt;
    t.ks = m->ks;
    t.qpart = m->qpart;
    t.nblk = m->nblk;
    t.npanels = (int)((nb + TILE - 1) / TILE);
    t.group = (int)std::max<long>(1, std::min<long>(m->opt_panel_group, t.npanels));
    t.b_cap = m->b_cap;
    t.ntri = m->ntri;
    SEGP_CHECK(launch_tri_sumsq(t, m->n_s, st));
    if (e1 != nullptr) SEGP_CUDA_CHECK(cudaEventRecord(e1, st));
    return SEGP_OK;
}

}  // namespace segp


// ============================================================================================== C ABI
extern "C" {

int segp_abi_version(void) { return SEGP_ABI_VERSION; }

const char* segp_last_error(void) { return g_err; }

int segp_create(segp_model** out, int device, int n_s_out, int n_s_in, int n_u, const int* kern_type) {
    if (out == nullptr || kern_type == nullptr) {
        set_error("segp_create: null argument");
        return SEGP_ERR_INVALID;
    }
    *out = nullptr;
    if (n_s_out < 1 || n_s_out > SEGP_MAX_NS || n_s_in < 1 || n_s_in > SEGP_MAX_NS || n_u < 0 || n_u > SEGP_MAX_NU) {
        set_error("segp_create: dimensions out of range (n_s_out=%d n_s_in=%d n_u=%d; limits %d/%d)", n_s_out, n_s_in,
                  n_u, SEGP_MAX_NS, SEGP_MAX_NU);
        return SEGP_ERR_INVALID;
    }
    for (int d = 0; d < n_s_out; ++d)
        if (kern_type[d] < SEGP_KERN_RBF || kern_type[d] > SEGP_KERN_LIN_MAT52) {
            set_error("segp_create: unsupported kernel type %d for output %d", kern_type[d], d);
            return SEGP_ERR_UNSUPPORTED;
        }
    int ndev = 0;
    SEGP_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("segp_create: no CUDA device %d (%d visible); this library has no CPU path", device, ndev);
        return SEGP_ERR_CUDA;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_create: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    cudaDeviceProp prop;
    SEGP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        set_error("segp_create: device %d is sm_%d%d; this build targets sm_100a", device, prop.major, prop.minor);
        return SEGP_ERR_UNSUPPORTED;
    }
    SEGP_CHECK(tri_sumsq_init());
    SEGP_CHECK(tri_i8_init());
    segp_model* m = new (std::nothrow) segp_model();
    if (m == nullptr) {
        set_error("out of host memory");
        return SEGP_ERR_INVALID;
    }
    m->device = device;
    m->n_s = n_s_out;
    m->n_in = n_s_in;
    m->n_u = n_u;
    m->dim = n_s_in + n_u;
    for (int d = 0; d < n_s_out; ++d) {
        m->kern[d] = kern_type[d];
        if (kern_is_composite(kern_type[d])) m->has_composite = true;
    }
    if (dev_alloc(&m->d_sp, 1) != SEGP_OK || dev_alloc(&m->fallback_counter, 1) != SEGP_OK ||
        cudaMemset(m->fallback_counter, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&m->h_fallback), sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) {
        dev_free(m->d_sp);
        dev_free(m->fallback_counter);
        delete m;
        set_error("segp_create: allocation failed");
        return SEGP_ERR_CUDA;
    }
    *m->h_fallback = 0;
    *out = m;
    return SEGP_OK;
}

int segp_destroy(segp_model* m) {
    if (m == nullptr) return SEGP_OK;
    DeviceGuard guard(m->device);
    cudaDeviceSynchronize();
    free_model_buffers(m);
    free_workspace(m);
    free_fact_cache(m);
    dev_free(m->d_sp);
    dev_free(m->i8_prof);
    dev_free(m->fallback_counter);
    if (m->h_fallback != nullptr) cudaFreeHost(m->h_fallback);
    if (m->s_host != nullptr) cudaStreamDestroy(m->s_host);
    if (m->s_cap != nullptr) cudaStreamDestroy(m->s_cap);
    if (m->s_copy != nullptr) cudaStreamDestroy(m->s_copy);
    for (int i = 0; i < 3; ++i) {
        if (m->s_sub[i] != nullptr) cudaStreamDestroy(m->s_sub[i]);
        if (m->ev_sub[i] != nullptr) cudaEventDestroy(m->ev_sub[i]);
    }
    if (m->ev_sub_fork != nullptr) cudaEventDestroy(m->ev_sub_fork);
    if (m->ev_chunk != nullptr) cudaEventDestroy(m->ev_chunk);
    for (cudaEvent_t e : m->tri_events) cudaEventDestroy(e);
    if (m->stage != nullptr) cudaFree(m->stage);
    delete m;
    return SEGP_OK;
}

int segp_set_model(segp_model* m, int n_train, const double* h_x, const double* h_y, const double* h_lengthscale,
                   const double* h_variance, const double* h_noise) {
    if (m == nullptr || h_x == nullptr || h_y == nullptr || h_lengthscale == nullptr || h_variance == nullptr ||
        h_noise == nullptr || n_train < 1) {
        set_error("segp_set_model: null argument or n_train < 1");
        return SEGP_ERR_INVALID;
    }
    for (int d = 0; d < m->n_s; ++d) {
        if (!(h_variance[d] > 0.0) || !(h_noise[d] >= 0.0)) {
            set_error("segp_set_model: variance must be > 0 and noise >= 0 (output %d)", d);
            return SEGP_ERR_INVALID;
        }
        for (int j = 0; j < m->dim; ++j) {
            const double l = h_lengthscale[d * m->dim + j];
            if (!(l > 0.0) || (std::isinf(l) && !kern_is_composite(m->kern[d]))) {
                set_error("segp_set_model: lengthscale[%d][%d] must be > 0 (and finite unless the kernel is composite)",
                          d, j);
                return SEGP_ERR_INVALID;
            }
        }
    }
    DeviceGuard guard(m->device);
    cudaDeviceSynchronize();
    free_model_buffers(m);
    free_workspace(m);
    const int dim = m->dim, n_s = m->n_s;
    m->n_train = n_train;
    m->n_pad = (n_train + TILE - 1) / TILE * TILE;
    m->nblk = m->n_pad / TILE;
    m->ntri = (long)m->nblk * (m->nblk + 1) / 2;
    m->h_x.assign(h_x, h_x + (size_t)n_train * dim);
    m->h_y.assign(h_y, h_y + (size_t)n_train * n_s);
    m->h_ls.assign(h_lengthscale, h_lengthscale + (size_t)n_s * dim);
    m->h_var.assign(h_variance, h_variance + n_s);
    m->h_noise.assign(h_noise, h_noise + n_s);

    std::vector<double> xs((size_t)n_s * m->n_pad * dim, 0.0), yp((size_t)n_s * m->n_pad, 0.0), invls((size_t)n_s * dim);
    for (int d = 0; d < n_s; ++d) {
        for (int j = 0; j < dim; ++j) invls[d * dim + j] = 1.0 / h_lengthscale[d * dim + j];
        for (int i = 0; i < n_train; ++i) {
            for (int j = 0; j < dim; ++j)
                xs[((size_t)d * m->n_pad + i) * dim + j] = h_x[(size_t)i * dim + j] * invls[d * dim + j];
            yp[(size_t)d * m->n_pad + i] = h_y[(size_t)i * n_s + d];
        }
    }
    SEGP_CHECK(dev_alloc(&m->xs, xs.size()));
    SEGP_CHECK(dev_alloc(&m->yp, yp.size()));
    SEGP_CHECK(dev_alloc(&m->invls, invls.size()));
    SEGP_CHECK(dev_alloc(&m->var, (size_t)n_s));
    SEGP_CUDA_CHECK(cudaMemcpy(m->xs, xs.data(), xs.size() * sizeof(double), cudaMemcpyHostToDevice));
    SEGP_CUDA_CHECK(cudaMemcpy(m->yp, yp.data(), yp.size() * sizeof(double), cudaMemcpyHostToDevice));
    SEGP_CUDA_CHECK(cudaMemcpy(m->invls, invls.data(), invls.size() * sizeof(double), cudaMemcpyHostToDevice));
    SEGP_CUDA_CHECK(cudaMemcpy(m->var, h_variance, n_s * sizeof(double), cudaMemcpyHostToDevice));
    if (m->has_composite) {
        std::vector<double> xraw((size_t)m->n_pad * dim, 0.0);
        std::copy(h_x, h_x + (size_t)n_train * dim, xraw.begin());
        SEGP_CHECK(dev_alloc(&m->xraw, xraw.size()));
        SEGP_CUDA_CHECK(cudaMemcpy(m->xraw, xraw.data(), xraw.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> xmax(dim, 0.0);
        for (int i = 0; i < n_train; ++i)
            for (int j = 0; j < dim; ++j) xmax[j] = std::max(xmax[j], std::fabs(h_x[(size_t)i * dim + j]));
        SEGP_CHECK(dev_alloc(&m->xmax, (size_t)dim));
        SEGP_CUDA_CHECK(cudaMemcpy(m->xmax, xmax.data(), dim * sizeof(double), cudaMemcpyHostToDevice));
    }
    m->has_data = true;
    return SEGP_OK;
}

int segp_set_linear_terms(segp_model* m, const double* h_prod_linear, const double* h_linear) {
    if (m == nullptr || h_prod_linear == nullptr || h_linear == nullptr) {
        set_error("segp_set_linear_terms: null argument");
        return SEGP_ERR_INVALID;
    }
    if (!m->has_data) {
        set_error("segp_set_linear_terms: call segp_set_model first");
        return SEGP_ERR_NOT_TRAINED;
    }
    if (!m->has_composite) return SEGP_OK;   // nothing uses them
    const size_t n = (size_t)m->n_s * m->dim;
    for (int d = 0; d < m->n_s; ++d)
        for (int j = 0; j < m->dim && kern_is_composite(m->kern[d]); ++j)
            if (!(h_prod_linear[d * m->dim + j] >= 0.0) || !(h_linear[d * m->dim + j] >= 0.0)) {
                set_error("segp_set_linear_terms: weights of output %d must be >= 0", d);
                return SEGP_ERR_INVALID;
            }
    DeviceGuard guard(m->device);
    m->h_plin.assign(h_prod_linear, h_prod_linear + n);
    m->h_lin.assign(h_linear, h_linear + n);
    for (int d = 0; d < m->n_s; ++d)
        if (!kern_is_composite(m->kern[d]))
            for (int j = 0; j < m->dim; ++j) m->h_plin[d * m->dim + j] = m->h_lin[d * m->dim + j] = 0.0;
    if (m->plin == nullptr) SEGP_CHECK(dev_alloc(&m->plin, n));
    if (m->lin == nullptr) SEGP_CHECK(dev_alloc(&m->lin, n));
    if (m->xtb == nullptr) SEGP_CHECK(dev_alloc(&m->xtb, n));
    SEGP_CUDA_CHECK(cudaMemcpy(m->plin, m->h_plin.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    SEGP_CUDA_CHECK(cudaMemcpy(m->lin, m->h_lin.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    SEGP_CUDA_CHECK(cudaMemset(m->xtb, 0, n * sizeof(double)));
    m->has_linear_terms = true;
    m->factorized = false;
    return SEGP_OK;
}

// X^T beta_d for the composite outputs (after beta is known: factorisation, or the broadcast of the factor buffers)
static int compute_xtb(segp_model* m, cudaStream_t st) {
    for (int d = 0; d < m->n_s; ++d)
        if (kern_is_composite(m->kern[d])) {
            SEGP_CHECK(launch_xtb(m->xraw, m->beta + (size_t)d * m->n_pad, m->xtb + (size_t)d * m->dim, m->n_pad, m->dim, st));
            ++m->launches;
        }
    return SEGP_OK;
}

// layout of the factor arena: [meta | beta | logdet | rowfac | rowfac_s | werr5 | werr4 | wm1 | wi8s | wi8], every part
// 256-byte aligned; the int8 parts only when the model can run on the tcgen05 path
static int alloc_arena(segp_model* m) {
    if (m->arena != nullptr) return SEGP_OK;
    const bool i8 = i8_capable(m);
    const size_t n_rows = (size_t)m->n_s * m->n_pad;
    const size_t n_kb = (size_t)m->n_s * m->nblk * (m->nblk + 1);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_meta = take(META_N * sizeof(double)), o_beta = take(n_rows * sizeof(double)),
                 o_logdet = take((size_t)m->n_s * sizeof(double));
    size_t o_rowfac = 0, o_rowfac_s = 0, o_werr5 = 0, o_werr4 = 0, o_wm1 = 0, o_wi8s = 0, o_wi8 = 0;
    if (i8) {
        o_rowfac = take(n_rows * sizeof(double));
        o_rowfac_s = take(n_rows * sizeof(double));
        o_werr5 = take(n_rows * sizeof(float));
        o_werr4 = take(n_rows * sizeof(float));
        o_wm1 = take((size_t)m->n_s * m->nblk * 2 * I8_A_TILE);
        o_wi8s = take(n_kb * (I8_SS * I8_A_TILE));
        o_wi8 = take(n_kb * (I8_S * I8_A_TILE));
    }
    SEGP_CHECK(dev_alloc(&m->arena, off));
    SEGP_CUDA_CHECK(cudaMemset(m->arena, 0, o_beta));   // meta
    m->arena_bytes = off;
    m->meta = reinterpret_cast<double*>(m->arena + o_meta);
    m->beta = reinterpret_cast<double*>(m->arena + o_beta);
    m->logdet = reinterpret_cast<double*>(m->arena + o_logdet);
    if (i8) {
        m->rowfac = reinterpret_cast<double*>(m->arena + o_rowfac);
        m->rowfac_s = reinterpret_cast<double*>(m->arena + o_rowfac_s);
        m->werr5 = reinterpret_cast<float*>(m->arena + o_werr5);
        m->werr4 = reinterpret_cast<float*>(m->arena + o_werr4);
        m->wm1 = reinterpret_cast<int8_t*>(m->arena + o_wm1);
        m->wi8s = reinterpret_cast<int8_t*>(m->arena + o_wi8s);
        m->wi8 = reinterpret_cast<int8_t*>(m->arena + o_wi8);
    }
    return SEGP_OK;
}

// does this model (as configured now) ever launch the float64 contraction?
static bool fp64_operand_needed(const segp_model* m) {
    return !i8_capable(m) || m->auto_fp64 || m->opt_tri_mode == 0 || m->opt_keep_fp64 != 0;
}

static int write_meta(segp_model* m) {
    double h[META_N] = {0};
    h[0] = (double)m->i8_primary;
    h[1] = m->auto_fp64 ? 1.0 : 0.0;
    h[2] = m->wt != nullptr ? 1.0 : 0.0;
    h[3] = m->unguarded ? 1.0 : 0.0;
    for (int i = 0; i < PROBE_STATS; ++i) h[4 + i] = m->probe_stat[i];
    SEGP_CUDA_CHECK(cudaMemcpy(m->meta, h, sizeof(h), cudaMemcpyHostToDevice));
    return SEGP_OK;
}

static int read_meta(segp_model* m, bool* root_has_fp64) {
    double h[META_N];
    SEGP_CUDA_CHECK(cudaMemcpy(h, m->meta, sizeof(h), cudaMemcpyDeviceToHost));
    m->i8_primary = h[0] == 4.0 ? 4 : 5;
    m->auto_fp64 = h[1] != 0.0;
    if (root_has_fp64 != nullptr) *root_has_fp64 = h[2] != 0.0;
    m->unguarded = h[3] != 0.0;
    for (int i = 0; i < PROBE_STATS; ++i) m->probe_stat[i] = h[4 + i];
    return SEGP_OK;
}

int segp_alloc_factor_buffers(segp_model* m) {
    if (m == nullptr || !m->has_data) {
        set_error("segp_alloc_factor_buffers: segp_set_model has not been called");
        return SEGP_ERR_NOT_TRAINED;
    }
    DeviceGuard guard(m->device);
    return alloc_arena(m);
}

int segp_alloc_fp64_operand(segp_model* m) {
    if (m == nullptr || !m->has_data) {
        set_error("segp_alloc_fp64_operand: segp_set_model has not been called");
        return SEGP_ERR_NOT_TRAINED;
    }
    DeviceGuard guard(m->device);
    if (m->wt == nullptr) SEGP_CHECK(dev_alloc(&m->wt, (size_t)m->n_s * m->ntri * TILE * TILE));
    return SEGP_OK;
}

int segp_num_factor_buffers(segp_model* m) {
    if (m == nullptr || m->arena == nullptr) return 0;
    return m->wt != nullptr ? 2 : 1;
}

int segp_factor_buffer(segp_model* m, int index, void** d_ptr, size_t* bytes) {
    if (m == nullptr || d_ptr == nullptr || bytes == nullptr || m->arena == nullptr) {
        set_error("segp_factor_buffer: buffers are not allocated");
        return SEGP_ERR_NOT_TRAINED;
    }
    if (index == 0) {
        *d_ptr = m->arena;
        *bytes = m->arena_bytes;
        return SEGP_OK;
    }
    if (index == 1 && m->wt != nullptr) {
        *d_ptr = m->wt;
        *bytes = (size_t)m->n_s * m->ntri * TILE * TILE * sizeof(double);
        return SEGP_OK;
    }
    set_error("segp_factor_buffer: index %d out of range", index);
    return SEGP_ERR_INVALID;
}

int segp_mark_factorized(segp_model* m) {
    if (m == nullptr || m->arena == nullptr) {
        set_error("segp_mark_factorized: buffers are not allocated");
        return SEGP_ERR_NOT_TRAINED;
    }
    DeviceGuard guard(m->device);
    bool root_has_fp64 = false;
    SEGP_CHECK(read_meta(m, &root_has_fp64));
    if (fp64_operand_needed(m) && m->wt == nullptr) {
        set_error("segp_mark_factorized: this model runs the float64 contraction; allocate (segp_alloc_fp64_operand) "
                  "and receive factor buffer 1 as well%s", root_has_fp64 ? "" : " (the factorising rank did not keep it: "
                  "set tri_mode / keep_fp64 there before segp_factorize)");
        return SEGP_ERR_INVALID;
    }
    if (m->has_composite) {
        if (!m->has_linear_terms) {
            set_error("segp_mark_factorized: composite kernel without linear terms (call segp_set_linear_terms)");
            return SEGP_ERR_INVALID;
        }
        SEGP_CHECK(compute_xtb(m, nullptr));
        SEGP_CUDA_CHECK(cudaStreamSynchronize(nullptr));
    }
    m->factorized = true;
    return SEGP_OK;
}

// destination pointers of output dimension d inside the factor arena
static PackI8Out pack_out(const segp_model* m, int d) {
    PackI8Out o{};
    const size_t kbs = (size_t)m->nblk * (m->nblk + 1);
    o.wi8 = m->wi8 + (size_t)d * kbs * (I8_S * I8_A_TILE);
    o.rowfac = m->rowfac + (size_t)d * m->n_pad;
    o.wi8s = m->wi8s + (size_t)d * kbs * (I8_SS * I8_A_TILE);
    o.wm1 = m->wm1 + (size_t)d * m->nblk * 2 * I8_A_TILE;
    o.rowfac_s = m->rowfac_s + (size_t)d * m->n_pad;
    o.werr5 = m->werr5 + (size_t)d * m->n_pad;
    o.werr4 = m->werr4 + (size_t)d * m->n_pad;
    return o;
}

// scale the error-model weights by the probe's calibration factors (standard deviation x rho -> variance x rho^2)
static int apply_calibration(segp_model* m, cudaStream_t st) {
    const double r4 = std::max(1.0, m->probe_stat[PS_RHO4]), r5 = std::max(1.0, m->probe_stat[PS_RHO5]);
    const long n = (long)m->n_s * m->n_pad;
    if (r4 > 1.0) SEGP_CHECK(launch_scale_f32(m->werr4, n, (float)(r4 * r4), st));
    if (r5 > 1.0) SEGP_CHECK(launch_scale_f32(m->werr5, n, (float)(r5 * r5), st));
    return SEGP_OK;
}

// ---------------------------------------------------------------------------------------------- factorize-time probe
// Calibrates the statistical error model of the int8 contraction and chooses the digit set: PROBE_N inputs drawn
// uniformly from the bounding box of the training inputs run through the float64 contraction (reference), the
// 15-product and the 10-product kernels.  Per output dimension the measured error of |v|^2 is compared with the
// model's a-posteriori estimate 2 sqrt(sum_i w_i v_i^2): if any probe exceeds 4.5 predicted standard deviations the
// weights of that digit set are inflated (rho); the digit set of the first pass is the 10-product one when the guard
// flags at most a quarter of the probes on it (and the model is large enough for the extra launches to pay); when the
// guard would flag more than a quarter of the probes even on the 15-product set, automatic mode runs float64.
static int probe_pass(segp_model* m, const double* d_z, long np, int mode, int digits, cudaStream_t st,
                      std::vector<double>& q, std::vector<double>& e, std::vector<double>* prior = nullptr) {
    m->force_mode = mode;
    m->force_digits = digits;
    int rc = ensure_workspace(m, np);
    if (rc == SEGP_OK) {
        KstarArgs k = base_kstar_args(m);
        k.z = d_z;
        k.n_batch = np;
        rc = run_kstar(m, k, st);
    }
    if (rc == SEGP_OK) rc = run_tri(m, np, st);
    m->force_mode = -2;
    m->force_digits = 0;
    SEGP_CHECK(rc);
    const size_t n = (size_t)m->n_s * m->nblk * m->b_cap;
    std::vector<double> hq(n);
    std::vector<float> he(mode != 0 ? n : 0);
    SEGP_CUDA_CHECK(cudaStreamSynchronize(st));
    SEGP_CUDA_CHECK(cudaMemcpy(hq.data(), m->qpart, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (mode != 0) SEGP_CUDA_CHECK(cudaMemcpy(he.data(), m->epart, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (prior != nullptr) {   // prior variance k(z, z) per probe: s_f^2, or per input for the composite kernels
        prior->assign((size_t)m->n_s * np, 0.0);
        std::vector<double> hk(m->has_composite ? (size_t)m->n_s * m->b_cap : 0);
        if (m->has_composite)
            SEGP_CUDA_CHECK(cudaMemcpy(hk.data(), m->kss, hk.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int d = 0; d < m->n_s; ++d)
            for (long p = 0; p < np; ++p)
                (*prior)[(size_t)d * np + p] = m->has_composite ? hk[(size_t)d * m->b_cap + p] : m->h_var[d];
    }
    q.assign((size_t)m->n_s * np, 0.0);
    e.assign((size_t)m->n_s * np, 0.0);
    for (int d = 0; d < m->n_s; ++d)
        for (int i = 0; i < m->nblk; ++i) {
            const size_t row = ((size_t)d * m->nblk + i) * m->b_cap;
            for (long p = 0; p < np; ++p) {
                q[(size_t)d * np + p] += hq[row + p];
                if (mode != 0) e[(size_t)d * np + p] += (double)he[row + p];
            }
        }
    return SEGP_OK;
}

static int run_probe(segp_model* m, cudaStream_t st) {
    const long np = PROBE_N;
    const int dim = m->dim;
    std::vector<double> lo(dim, 1e300), hi(dim, -1e300), z((size_t)np * dim);
    for (int i = 0; i < m->n_train; ++i)
        for (int j = 0; j < dim; ++j) {
            lo[j] = std::min(lo[j], m->h_x[(size_t)i * dim + j]);
            hi[j] = std::max(hi[j], m->h_x[(size_t)i * dim + j]);
        }
    uint64_t state = 0x9E3779B97F4A7C15ull;
    auto next_unit = [&state]() {   // splitmix64 -> [0, 1)
        uint64_t x = (state += 0x9E3779B97F4A7C15ull);
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        return (double)(x >> 11) * (1.0 / 9007199254740992.0);
    };
    for (long p = 0; p < np; ++p)
        for (int j = 0; j < dim; ++j) z[(size_t)p * dim + j] = lo[j] + (hi[j] - lo[j]) * next_unit();
    double* d_z = nullptr;
    SEGP_CHECK(dev_alloc(&d_z, z.size()));
    const bool timed = m->time_tri;
    m->time_tri = false;
    std::vector<double> q0, e0, q5, e5, q4, e4, prior;
    int rc = SEGP_OK;
    do {
        if (cudaMemcpy(d_z, z.data(), z.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("probe: upload failed");
            rc = SEGP_ERR_CUDA;
            break;
        }
        const bool timing = getenv("SEGP_FACT_TIMING") != nullptr;
        auto now_ms = []() {
            return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
        };
        double t0 = now_ms();
        auto lap = [&](const char* what) {
            if (!timing) return;
            const double t = now_ms();
            fprintf(stderr, "[probe] %-24s %8.2f ms\n", what, t - t0);
            t0 = t;
        };
        if ((rc = probe_pass(m, d_z, np, 0, 0, st, q0, e0, &prior)) != SEGP_OK) break;
        lap("float64 pass");
        if ((rc = probe_pass(m, d_z, np, -1, 5, st, q5, e5)) != SEGP_OK) break;
        lap("15-product pass");
        if ((rc = probe_pass(m, d_z, np, -1, 4, st, q4, e4)) != SEGP_OK) break;
        lap("10-product pass");
    } while (0);
    m->time_tri = timed;
    dev_free(d_z);
    cudaDeviceSynchronize();
    free_workspace(m);
    SEGP_CHECK(rc);

    constexpr double KAPPA_PROBE = 4.5;
    double* ps = m->probe_stat;
    for (int i = 0; i < PROBE_STATS; ++i) ps[i] = 0.0;
    ps[PS_RAN] = 1.0;
    ps[PS_RHO4] = ps[PS_RHO5] = 1.0;
    ps[PS_MINVAR] = 1e300;
    for (int d = 0; d < m->n_s; ++d)
        for (long p = 0; p < np; ++p) {
            const size_t i = (size_t)d * np + p;
            const double s2 = prior[i] - q0[i];
            const double err4 = std::fabs(q4[i] - q0[i]), err5 = std::fabs(q5[i] - q0[i]);
            const double sd4 = 2.0 * std::sqrt(e4[i]), sd5 = 2.0 * std::sqrt(e5[i]);
            ps[PS_ERR4] = std::max(ps[PS_ERR4], err4);
            ps[PS_ERR5] = std::max(ps[PS_ERR5], err5);
            if (s2 > 0.0) {
                ps[PS_REL4] = std::max(ps[PS_REL4], err4 / s2);
                ps[PS_REL5] = std::max(ps[PS_REL5], err5 / s2);
            }
            // the float64 reference itself carries ~1e-13 of rounding relative to k**: do not calibrate against that
            const double floor = 1e-12 * prior[i];
            if (sd4 > 0.0 && err4 > floor) ps[PS_RATIO4] = std::max(ps[PS_RATIO4], err4 / sd4);
            if (sd5 > 0.0 && err5 > floor) ps[PS_RATIO5] = std::max(ps[PS_RATIO5], err5 / sd5);
            if (prior[i] > 0.0) ps[PS_MINVAR] = std::min(ps[PS_MINVAR], s2 / prior[i]);
        }
    ps[PS_RHO4] = std::max(1.0, ps[PS_RATIO4] / KAPPA_PROBE);
    ps[PS_RHO5] = std::max(1.0, ps[PS_RATIO5] / KAPPA_PROBE);
    // guard decisions on the probes with the calibrated model
    const double gs = guard_gs(m);
    long n4 = 0, n5 = 0;
    for (long p = 0; p < np; ++p) {
        bool f4 = false, f5 = false;
        for (int d = 0; d < m->n_s; ++d) {
            const size_t i = (size_t)d * np + p;
            const double s2 = prior[i] - q0[i];
            if (!(s2 > 0.0)) {
                f4 = f5 = true;
                continue;
            }
            if (gs * ps[PS_RHO4] * ps[PS_RHO4] * e4[i] > s2 * s2) f4 = true;
            ps[PS_MARGIN4] = std::max(ps[PS_MARGIN4], std::sqrt(gs * e4[i]) * ps[PS_RHO4] / s2);
            if (gs * ps[PS_RHO5] * ps[PS_RHO5] * e5[i] > s2 * s2) f5 = true;
        }
        n4 += f4;
        n5 += f5;
    }
    ps[PS_FRAC4] = (double)n4 / (double)np;
    ps[PS_FRAC5] = (double)n5 / (double)np;
    // 10-product first pass + recomputation of a flagged fraction g of the panels costs (10 + 15 g) / 15 of the
    // 15-product kernel: it pays while g < 1/3.  The probes are spread over the whole training box, a rollout batch
    // sits in one place, so the probe fraction only says whether typical inputs pass; the run-time demotion in
    // segp_multistep (fallback_rate) catches a batch that lives where they do not.
    // Where the estimate stays 4x inside the tolerance on every probe the first pass runs unguarded: no guard and no
    // recomputation launch per step (the ellipsoid step still raises SEGP_STATUS_LOW_PRECISION on whatever exceeds the
    // tolerance).  Models below 1024 padded points stay on the 15-product set: their steps are launch-latency-bound
    // (the contraction of C2 is 33 us whatever the digit set), so there is nothing to buy with the coarser digits.
    m->unguarded = ps[PS_MARGIN4] <= 0.25;
    m->i8_primary = (ps[PS_FRAC4] <= 0.25 && m->n_pad >= 1024) ? 4 : 5;
    m->auto_fp64 = ps[PS_FRAC5] > 0.25;
    m->demoted = false;
    SEGP_CHECK(apply_calibration(m, st));
    SEGP_CUDA_CHECK(cudaStreamSynchronize(st));
    return SEGP_OK;
}


int segp_factorize(segp_model* m, void* stream) {
    if (m == nullptr || !m->has_data) {
        set_error("segp_factorize: segp_set_model has not been called");
        return SEGP_ERR_NOT_TRAINED;
    }
    if (m->has_composite && !m->has_linear_terms) {
        set_error("segp_factorize: composite kernel without linear terms (call segp_set_linear_terms)");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    m->factorized = false;
    free_graphs(m);
    // host-side phase timer (SEGP_FACT_TIMING=1 prints it): where a factorisation's wall time goes
    const bool timing = getenv("SEGP_FACT_TIMING") != nullptr;
    auto now_ms = []() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    double t_phase = now_ms();
    auto phase = [&](const char* what) {
        if (!timing) return;
        const double t = now_ms();
        fprintf(stderr, "[segp_factorize] %-28s %8.2f ms\n", what, t - t_phase);
        t_phase = t;
    };
    SEGP_CHECK(alloc_arena(m));
    // the float64 operand: the probe's reference, and the contraction itself where the int8 path cannot run
    SEGP_CHECK(segp_alloc_fp64_operand(m));
    m->i8_primary = 5;
    m->auto_fp64 = false;
    for (double& v : m->probe_stat) v = 0.0;
    m->unguarded = false;
    m->demoted = false;
    const size_t nn = (size_t)m->n_pad * m->n_pad;
    const int nb64 = m->n_pad / NBLK;
    // The output dimensions are independent factorisations and each is a chain of ~270 launches, many of them one
    // block wide (the 64 x 64 diagonal blocks): up to FACT_SLOTS of them run concurrently on their own streams, each
    // with its own scratch, so the latency-bound launches of one overlap the GEMMs of the others.
    constexpr int FACT_SLOTS = 4;
    struct Slot : segp_model::FactScratch {
        cudaStream_t s = nullptr;
        cudaEvent_t done = nullptr;
    };
    int nslots = std::min(m->n_s, FACT_SLOTS);
    if (const char* e = getenv("SEGP_FACT_SLOTS")) nslots = std::max(1, std::min(nslots, atoi(e)));   // tuning experiments
    Slot slots[FACT_SLOTS];
    // dense GEMMs of potrf / trtri on the tensor cores (fact_i8.cu) where the model is large enough for the splits to
    // pay (below ~1000 points the chain of 64 x 64 diagonal blocks sets the time, not the GEMMs)
    const bool use_fd = m->n_pad >= 512 && m->n_pad <= I8_MAX_NPAD &&
                        (m->opt_fact_i8 == 1 || (m->opt_fact_i8 < 0 && m->n_pad >= 1024));
    m->last_fact_i8 = use_fd;
    cudaEvent_t fork = nullptr;
    int* d_fail = nullptr;
    int rc = SEGP_OK;
    std::vector<int> fails(m->n_s, 0);
    // scratch kept from the previous factorisation of a same-size model
    const bool need_w = !m->opt_keep_w;
    const bool cache_hit = m->fact_cache_npad == m->n_pad && m->fact_cache_slots >= nslots && m->fact_cache_fd == use_fd &&
                           m->fact_cache_w == need_w;
    if (cache_hit) {
        for (int i = 0; i < nslots; ++i) {
            static_cast<segp_model::FactScratch&>(slots[i]) = m->fact_cache[i];
            m->fact_cache[i] = segp_model::FactScratch();
        }
    }
    if (!cache_hit || m->fact_cache_slots > nslots) free_fact_cache(m);
    m->fact_cache_npad = m->fact_cache_slots = 0;
    const size_t slot_bytes = (3 * nn + (size_t)nb64 * NBLK * NBLK + 33 * (size_t)m->n_pad) * sizeof(double) +
                              (use_fd ? 2 * fd_scratch_plane_bytes(m->n_pad) + 2 * m->n_pad * sizeof(double) : 0);
    do {
        if (m->opt_keep_w && m->wdense == nullptr && (rc = dev_alloc(&m->wdense, (size_t)m->n_s * nn)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_fail, (size_t)m->n_s)) != SEGP_OK) break;
        if (cudaMemsetAsync(d_fail, 0, m->n_s * sizeof(int), st) != cudaSuccess ||
            cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(fork, st) != cudaSuccess) {
            set_error("segp_factorize: stream set-up failed");
            rc = SEGP_ERR_CUDA;
            break;
        }
        for (int i = 0; i < nslots && rc == SEGP_OK; ++i) {
            Slot& sl = slots[i];
            if (!cache_hit) {
                if ((rc = dev_alloc(&sl.kbuf, nn)) != SEGP_OK) break;
                if (need_w && (rc = dev_alloc(&sl.wbuf, nn)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.tmp, nn)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.diag_inv, (size_t)nb64 * NBLK * NBLK)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.u_tmp, (size_t)33 * m->n_pad)) != SEGP_OK) break;
            }
            if (use_fd && !cache_hit) {
                const size_t pb = fd_scratch_plane_bytes(m->n_pad);
                if ((rc = dev_alloc(&sl.fd.ap, pb)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.fd.bp, pb)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.fd.as, (size_t)m->n_pad)) != SEGP_OK) break;
                if ((rc = dev_alloc(&sl.fd.bs, (size_t)m->n_pad)) != SEGP_OK) break;
            }
            if (cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess ||
                cudaStreamWaitEvent(sl.s, fork, 0) != cudaSuccess) {
                set_error("segp_factorize: stream set-up failed");
                rc = SEGP_ERR_CUDA;
            }
        }
        if (rc != SEGP_OK) break;
        phase("arena + scratch allocation");
        SetupDims sd{m->n_train, m->n_pad, m->dim};
        for (int d = 0; d < m->n_s && rc == SEGP_OK; ++d) {
            Slot& sl = slots[d % nslots];   // a slot's buffers are reused in stream order
            cudaStream_t ss = sl.s;
            double* wbuf = m->opt_keep_w ? m->wdense + (size_t)d * nn : sl.wbuf;
            const double* xs_d = m->xs + (size_t)d * m->n_pad * m->dim;
            const bool comp = kern_is_composite(m->kern[d]);
            if ((rc = launch_kmat(sl.kbuf, xs_d, m->kern[d], m->h_var[d], m->h_noise[d], sd, comp ? m->xraw : nullptr,
                                  comp ? m->plin + (size_t)d * m->dim : nullptr,
                                  comp ? m->lin + (size_t)d * m->dim : nullptr, ss)) != SEGP_OK)
                break;
            ++m->launches;
            if ((rc = potrf_lower(sl.kbuf, m->n_pad, sl.diag_inv, d_fail + d, ss, &m->launches, use_fd ? &sl.fd : nullptr)) != SEGP_OK)
                break;
            if ((rc = logdet_from_chol(sl.kbuf, m->n_train, m->n_pad, m->logdet + d, ss)) != SEGP_OK) break;
            ++m->launches;
            if (cudaMemsetAsync(wbuf, 0, nn * sizeof(double), ss) != cudaSuccess) {
                set_error("cudaMemsetAsync failed");
                rc = SEGP_ERR_CUDA;
                break;
            }
            if ((rc = trtri_lower(sl.kbuf, wbuf, m->n_pad, sl.diag_inv, sl.tmp, ss, &m->launches, use_fd ? &sl.fd : nullptr)) != SEGP_OK)
                break;
            if ((rc = solve_beta(wbuf, m->yp + (size_t)d * m->n_pad, sl.u_tmp, m->beta + (size_t)d * m->n_pad, m->n_pad,
                                 ss)) != SEGP_OK)
                break;
            m->launches += 3;
            if ((rc = pack_w(wbuf, m->wt + (size_t)d * m->ntri * TILE * TILE, m->n_pad, ss)) != SEGP_OK) break;
            ++m->launches;
            if (m->wi8 != nullptr) {
                if ((rc = pack_w_i8(wbuf, pack_out(m, d), comp ? 1.0 : m->h_var[d], m->n_pad, m->n_train, ss)) != SEGP_OK)
                    break;
                m->launches += 2;
            }
        }
        if (rc != SEGP_OK) break;
        phase("enqueue (host)");
        for (int i = 0; i < nslots; ++i) {   // join
            if (cudaEventRecord(slots[i].done, slots[i].s) != cudaSuccess ||
                cudaStreamWaitEvent(st, slots[i].done, 0) != cudaSuccess) {
                set_error("segp_factorize: stream join failed");
                rc = SEGP_ERR_CUDA;
                break;
            }
        }
        if (rc != SEGP_OK) break;
        if (m->has_composite && (rc = compute_xtb(m, st)) != SEGP_OK) break;
        cudaError_t e = cudaStreamSynchronize(st);
        phase("wait for the GPU");
        if (e == cudaSuccess)
            e = cudaMemcpy(fails.data(), d_fail, m->n_s * sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("segp_factorize: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
            break;
        }
        for (int d = 0; d < m->n_s; ++d)
            if (fails[d] != 0) {
                set_error("segp_factorize: K + noise I is not positive definite for output %d (pivot %d)", d,
                          fails[d] - 1);
                rc = SEGP_ERR_NOT_POSDEF;
                break;
            }
    } while (0);
    cudaDeviceSynchronize();
    const bool keep_scratch = rc == SEGP_OK && (m->opt_scratch_cache == 1 ||
                                                (m->opt_scratch_cache < 0 && (size_t)nslots * slot_bytes <= ((size_t)4 << 30)));
    if (keep_scratch) {
        for (int i = 0; i < nslots; ++i) {
            m->fact_cache[i] = static_cast<segp_model::FactScratch&>(slots[i]);
            static_cast<segp_model::FactScratch&>(slots[i]) = segp_model::FactScratch();
        }
        m->fact_cache_npad = m->n_pad;
        m->fact_cache_slots = nslots;
        m->fact_cache_fd = use_fd;
        m->fact_cache_w = need_w;
    }
    for (int i = 0; i < FACT_SLOTS; ++i) {
        Slot& sl = slots[i];
        dev_free(sl.kbuf);
        dev_free(sl.wbuf);
        dev_free(sl.tmp);
        dev_free(sl.diag_inv);
        dev_free(sl.u_tmp);
        dev_free(sl.fd.ap);
        dev_free(sl.fd.bp);
        dev_free(sl.fd.as);
        dev_free(sl.fd.bs);
        if (sl.done != nullptr) cudaEventDestroy(sl.done);
        if (sl.s != nullptr) cudaStreamDestroy(sl.s);
    }
    if (fork != nullptr) cudaEventDestroy(fork);
    dev_free(d_fail);
    phase("free scratch");
    if (rc != SEGP_OK) return rc;
    m->factorized = true;
    // calibrate the int8 error model and choose the digit set (needs the float64 operand as the reference)
    if (i8_capable(m) && m->opt_probe != 0) {
        rc = run_probe(m, st);
        if (rc != SEGP_OK) {
            m->factorized = false;
            return rc;
        }
    }
    phase("probe");
    if (!fp64_operand_needed(m)) dev_free(m->wt);
    const int rc_meta = write_meta(m);
    phase("free fp64 operand + meta");
    return rc_meta;
}

int segp_append(segp_model* m, int n_new, const double* h_x, const double* h_y, void* stream) {
    SEGP_CHECK(check_ready(m));
    if (n_new < 0 || (n_new > 0 && (h_x == nullptr || h_y == nullptr))) {
        set_error("segp_append: null argument");
        return SEGP_ERR_INVALID;
    }
    if (n_new == 0) return SEGP_OK;
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int dim = m->dim, n_s = m->n_s;
    const int n_old = m->n_train, n_tot = n_old + n_new;
    // host copies first (copies of the arguments: h_x may alias nothing of ours, but set_model re-assigns from them)
    std::vector<double> hx(m->h_x), hy(m->h_y);
    hx.insert(hx.end(), h_x, h_x + (size_t)n_new * dim);
    hy.insert(hy.end(), h_y, h_y + (size_t)n_new * n_s);
    m->last_append_incremental = false;
    if (m->wdense == nullptr || n_tot > m->n_pad) {
        // no dense factor kept yet, or the padded size grows: factorise from scratch, keeping W dense from now on
        const std::vector<double> ls(m->h_ls), var(m->h_var), noise(m->h_noise), plin(m->h_plin), lin(m->h_lin);
        const bool had_lin = m->has_linear_terms;
        m->opt_keep_w = 1;
        SEGP_CHECK(segp_set_model(m, n_tot, hx.data(), hy.data(), ls.data(), var.data(), noise.data()));
        if (had_lin) SEGP_CHECK(segp_set_linear_terms(m, plin.data(), lin.data()));
        return segp_factorize(m, stream);
    }
    // ---- incremental: the padded size and every earlier row of W stay
    SEGP_CUDA_CHECK(cudaStreamSynchronize(st));
    m->factorized = false;
    const int n_pad = m->n_pad;
    const size_t nn = (size_t)n_pad * n_pad;
    std::vector<double> rows((size_t)n_new * dim);
    for (int d = 0; d < n_s; ++d) {
        for (int i = 0; i < n_new; ++i)
            for (int j = 0; j < dim; ++j) rows[(size_t)i * dim + j] = h_x[(size_t)i * dim + j] / m->h_ls[d * dim + j];
        SEGP_CUDA_CHECK(cudaMemcpy(m->xs + ((size_t)d * n_pad + n_old) * dim, rows.data(), rows.size() * sizeof(double),
                                   cudaMemcpyHostToDevice));
        for (int i = 0; i < n_new; ++i) rows[i] = h_y[(size_t)i * n_s + d];
        SEGP_CUDA_CHECK(cudaMemcpy(m->yp + (size_t)d * n_pad + n_old, rows.data(), (size_t)n_new * sizeof(double),
                                   cudaMemcpyHostToDevice));
    }
    if (m->xraw != nullptr) {
        SEGP_CUDA_CHECK(cudaMemcpy(m->xraw + (size_t)n_old * dim, h_x, (size_t)n_new * dim * sizeof(double),
                                   cudaMemcpyHostToDevice));
        std::vector<double> xmax(dim, 0.0);
        for (size_t i = 0; i < hx.size() / dim; ++i)
            for (int j = 0; j < dim; ++j) xmax[j] = std::max(xmax[j], std::fabs(hx[i * dim + j]));
        SEGP_CUDA_CHECK(cudaMemcpy(m->xmax, xmax.data(), dim * sizeof(double), cudaMemcpyHostToDevice));
    }
    m->h_x.swap(hx);
    m->h_y.swap(hy);
    m->n_train = n_tot;
    const int r0 = n_old / NBLK * NBLK;
    const int r1 = (n_tot + NBLK - 1) / NBLK * NBLK;
    const int sblk = r1 - r0;
    double *krows = nullptr, *l21 = nullptr, *tbuf = nullptr, *sbuf = nullptr, *w22 = nullptr, *tmp = nullptr,
           *diag_inv = nullptr, *u_tmp = nullptr;
    int* d_fail = nullptr;
    int rc = SEGP_OK;
    std::vector<int> fails(n_s, 0);
    do {
        if ((rc = dev_alloc(&krows, (size_t)sblk * n_pad)) != SEGP_OK) break;
        if ((rc = dev_alloc(&l21, (size_t)sblk * n_pad)) != SEGP_OK) break;
        if ((rc = dev_alloc(&tbuf, (size_t)sblk * n_pad)) != SEGP_OK) break;
        if ((rc = dev_alloc(&sbuf, (size_t)sblk * sblk)) != SEGP_OK) break;
        if ((rc = dev_alloc(&w22, (size_t)sblk * sblk)) != SEGP_OK) break;
        if ((rc = dev_alloc(&tmp, (size_t)sblk * sblk)) != SEGP_OK) break;
        if ((rc = dev_alloc(&diag_inv, (size_t)sblk * NBLK)) != SEGP_OK) break;
        if ((rc = dev_alloc(&u_tmp, (size_t)33 * n_pad)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_fail, (size_t)n_s)) != SEGP_OK) break;
        SetupDims sd{m->n_train, n_pad, dim};
        for (int d = 0; d < n_s && rc == SEGP_OK; ++d) {
            const bool comp = kern_is_composite(m->kern[d]);
            double* w = m->wdense + (size_t)d * nn;
            if ((rc = launch_kmat_rows(krows, m->xs + (size_t)d * n_pad * dim, m->kern[d], m->h_var[d], m->h_noise[d], sd,
                                       comp ? m->xraw : nullptr, comp ? m->plin + (size_t)d * dim : nullptr,
                                       comp ? m->lin + (size_t)d * dim : nullptr, r0, sblk, st)) != SEGP_OK)
                break;
            ++m->launches;
            if ((rc = append_rows(w, n_pad, r0, sblk, krows, l21, tbuf, sbuf, w22, tmp, diag_inv, d_fail + d, st,
                                  &m->launches)) != SEGP_OK)
                break;
            if ((rc = logdet_from_winv(w, m->n_train, n_pad, m->logdet + d, st)) != SEGP_OK) break;
            if ((rc = solve_beta(w, m->yp + (size_t)d * n_pad, u_tmp, m->beta + (size_t)d * n_pad, n_pad, st)) != SEGP_OK)
                break;
            if (m->wt != nullptr && (rc = pack_w(w, m->wt + (size_t)d * m->ntri * TILE * TILE, n_pad, st)) != SEGP_OK)
                break;
            m->launches += 5;
            if (m->wi8 != nullptr) {
                if ((rc = pack_w_i8(w, pack_out(m, d), comp ? 1.0 : m->h_var[d], n_pad, m->n_train, st)) != SEGP_OK) break;
                m->launches += 2;
            }
        }
        if (rc != SEGP_OK) break;
        if (m->has_composite && (rc = compute_xtb(m, st)) != SEGP_OK) break;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaMemcpy(fails.data(), d_fail, n_s * sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("segp_append: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
            break;
        }
        for (int d = 0; d < n_s; ++d)
            if (fails[d] != 0) {
                set_error("segp_append: K + noise I is not positive definite for output %d (pivot %d)", d,
                          r0 + fails[d] - 1);
                rc = SEGP_ERR_NOT_POSDEF;
                break;
            }
    } while (0);
    cudaStreamSynchronize(st);
    dev_free(krows);
    dev_free(l21);
    dev_free(tbuf);
    dev_free(sbuf);
    dev_free(w22);
    dev_free(tmp);
    dev_free(diag_inv);
    dev_free(u_tmp);
    dev_free(d_fail);
    if (rc == SEGP_OK && m->wi8 != nullptr) rc = apply_calibration(m, st);   // the probe's factors carry over
    if (rc == SEGP_OK) {
        m->factorized = true;
        m->last_append_incremental = true;
        free_graphs(m);
        return write_meta(m);
    }
    // The update failed half way (rows of W overwritten, data already appended): restore the model the caller had
    // -- the old data, factorised from scratch -- so that the handle stays usable and a retry does not append twice.
    {
        char msg[sizeof(g_err)];
        snprintf(msg, sizeof(msg), "%s", g_err);
        const std::vector<double> ls(m->h_ls), var(m->h_var), noise(m->h_noise), plin(m->h_plin), lin(m->h_lin);
        const bool had_lin = m->has_linear_terms;
        // after the swaps above hx / hy hold the previous training set
        if (segp_set_model(m, n_old, hx.data(), hy.data(), ls.data(), var.data(), noise.data()) == SEGP_OK &&
            (!had_lin || segp_set_linear_terms(m, plin.data(), lin.data()) == SEGP_OK))
            segp_factorize(m, stream);
        set_error("%s (the previous model was restored)", msg);
    }
    return rc;
}

int segp_beta(segp_model* m, double* h_out) {
    SEGP_CHECK(check_ready(m));
    if (h_out == nullptr) {
        set_error("segp_beta: null output");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(m->device);
    for (int d = 0; d < m->n_s; ++d)
        SEGP_CUDA_CHECK(cudaMemcpy(h_out + (size_t)d * m->n_train, m->beta + (size_t)d * m->n_pad,
                                   (size_t)m->n_train * sizeof(double), cudaMemcpyDeviceToHost));
    return SEGP_OK;
}

int segp_logdet(segp_model* m, double* h_out) {
    SEGP_CHECK(check_ready(m));
    DeviceGuard guard(m->device);
    SEGP_CUDA_CHECK(cudaMemcpy(h_out, m->logdet, m->n_s * sizeof(double), cudaMemcpyDeviceToHost));
    return SEGP_OK;
}

// make the float64 operand resident again (it is dropped after factorising when automatic mode does not need it):
// the packed tiles come from the dense W, which is gone, so the model is factorised once more
static int ensure_fp64_operand(segp_model* m, void* stream) {
    if (m->wt != nullptr || tri_mode(m) != 0) return SEGP_OK;
    const long keep = m->opt_keep_fp64;
    m->opt_keep_fp64 = 1;
    const int rc = segp_factorize(m, stream);
    m->opt_keep_fp64 = keep;
    return rc;
}

int segp_predict_ex(segp_model* m, long n_batch, const double* d_z, double* d_mu, double* d_var, double* d_jac,
                    int32_t* d_status, void* stream) {
    SEGP_CHECK(check_ready(m));
    if (n_batch < 0 || (n_batch > 0 && (d_z == nullptr || d_mu == nullptr || d_var == nullptr))) {
        set_error("segp_predict: null buffer");
        return SEGP_ERR_INVALID;
    }
    if (n_batch == 0) return SEGP_OK;
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SEGP_CHECK(ensure_fp64_operand(m, stream));
    SEGP_CHECK(ensure_workspace(m, n_batch));
    for (long c0 = 0; c0 < n_batch; c0 += m->b_cap) {
        const long nb = std::min<long>(m->b_cap, n_batch - c0);
        KstarArgs k = base_kstar_args(m);
        k.z = d_z + c0 * m->dim;
        k.n_batch = nb;
        SEGP_CHECK(run_kstar(m, k, st));
        SEGP_CHECK(run_tri(m, nb, st));
        FinalizeArgs f{};
        f.mu_part = m->mu_part;
        f.jac_part = m->jac_part;
        f.qpart = m->qpart;
        f.gp_var = m->var;
        f.invls = m->invls;
        f.jac2_part = m->jac2_part;
        f.kss = m->kss;
        f.epart = (m->ws_mode != 0 && m->opt_guard != 0) ? m->epart : nullptr;
        f.guard_gs = guard_gs(m);
        f.status = d_status ? d_status + c0 : nullptr;
        f.nsplit = m->nsplit;
        f.nblk = m->nblk;
        f.n_s = m->n_s;
        f.dim = m->dim;
        f.b_cap = m->b_cap;
        f.n_batch = nb;
        f.mu = d_mu + c0 * m->n_s;
        f.var = d_var + c0 * m->n_s;
        f.jac = d_jac ? d_jac + c0 * m->n_s * m->dim : nullptr;
        SEGP_CHECK(launch_finalize_predict(f, st));
        m->launches += 3;
    }
    return SEGP_OK;
}

int segp_predict(segp_model* m, long n_batch, const double* d_z, double* d_mu, double* d_var, double* d_jac,
                 void* stream) {
    return segp_predict_ex(m, n_batch, d_z, d_mu, d_var, d_jac, nullptr, stream);
}

int segp_multistep(segp_model* m, long n_batch, int horizon, const double* d_p0, long p0_stride, const double* d_q0,
                   long q0_stride, const double* d_k_ff, const double* d_k_fb, long kfb_stride,
                   const double* d_k_fb_init, long kfb_init_stride, const segp_reach_params* params, double* d_p_all,
                   double* d_q_all, double* d_var_all, int32_t* d_status, void* stream) {
    SEGP_CHECK(check_ready(m));
    if (n_batch < 0 || horizon < 1) {
        set_error("segp_multistep: n_batch >= 0 and horizon >= 1 required");
        return SEGP_ERR_INVALID;
    }
    if (n_batch == 0) return SEGP_OK;
    if (d_p0 == nullptr || d_k_ff == nullptr || d_p_all == nullptr || d_q_all == nullptr) {
        set_error("segp_multistep: null buffer");
        return SEGP_ERR_INVALID;
    }
    if (horizon > 1 && d_k_fb == nullptr) {
        set_error("segp_multistep: k_fb is required for horizon > 1");
        return SEGP_ERR_INVALID;
    }
    if (d_q0 != nullptr && d_k_fb_init == nullptr) {
        set_error("segp_multistep: k_fb_init is required when an initial shape matrix q_0 is given");
        return SEGP_ERR_INVALID;
    }
    StepParams sp;
    SEGP_CHECK(fill_step_params(&sp, params, m->n_s, m->n_in, m->n_u));
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Run-time demotion of automatic mode: the previous guarded call mirrored the recomputation counter to the host
    // (asynchronously; it has landed by the time a caller that consumed that call's results is back here).  When more
    // than a third of its panel contractions were redone on the 15-product set, the 10-product first pass costs more
    // than it saves for this batch: start on the 15-product set from now on (until the next factorisation).
    if (m->panels_prev_call > 0 && !m->demoted && m->opt_i8_digits == 0) {
        const unsigned int now = *reinterpret_cast<volatile unsigned int*>(m->h_fallback);
        if ((double)(now - m->fallback_seen) > (double)m->panels_prev_call / 3.0) m->demoted = true;
        m->fallback_seen = now;
    }
    m->panels_prev_call = 0;
    SEGP_CHECK(ensure_fp64_operand(m, stream));
    SEGP_CHECK(ensure_workspace(m, n_batch));
    // pageable source: staged by the runtime before the call returns, and outside any captured graph
    SEGP_CUDA_CHECK(cudaMemcpyAsync(m->d_sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, st));

    const int n_s = m->n_s, n_u = m->n_u;
    const long hs = (long)horizon * n_s, hss = (long)horizon * n_s * n_s;
    // argument blocks of step t for the trajectories [b0, b1) of the chunk starting at c0 (indices inside the
    // kernels are relative to the chunk; b1 / the panel range end bound the launch)
    auto kstar_args = [&](long c0, int t, long b1) {
        KstarArgs k = base_kstar_args(m);
        k.z = nullptr;
        k.p = (t == 0) ? d_p0 + c0 * p0_stride : d_p_all + c0 * hs + (long)(t - 1) * n_s;
        k.p_stride = (t == 0) ? p0_stride : hs;
        k.kff = d_k_ff + (c0 * horizon + t) * n_u;
        k.kff_stride = (long)horizon * n_u;
        k.sp = m->d_sp;
        k.n_batch = b1;
        return k;
    };
    auto step_args = [&](long c0, int t, long b0, long b1) {
        StepArgs s{};
        s.mu_part = m->mu_part;
        s.jac_part = m->jac_part;
        s.qpart = m->qpart;
        s.gp_var = m->var;
        s.invls = m->invls;
        s.jac2_part = m->jac2_part;
        s.kss = m->kss;
        s.epart = (m->ws_mode != 0 && m->opt_guard != 0) ? m->epart : nullptr;
        s.guard_gs = guard_gs(m);
        s.nsplit = m->nsplit;
        s.nblk = m->nblk;
        s.b_cap = m->b_cap;
        s.p = (t == 0) ? d_p0 + c0 * p0_stride : d_p_all + c0 * hs + (long)(t - 1) * n_s;
        s.p_stride = (t == 0) ? p0_stride : hs;
        if (t == 0) {
            s.q = d_q0 ? d_q0 + c0 * q0_stride : nullptr;
            s.q_stride = q0_stride;
            s.kfb = d_k_fb_init ? d_k_fb_init + c0 * kfb_init_stride : nullptr;
            s.kfb_stride = kfb_init_stride;
        } else {
            s.q = d_q_all + c0 * hss + (long)(t - 1) * n_s * n_s;
            s.q_stride = hss;
            s.kfb = d_k_fb + c0 * kfb_stride + (long)(t - 1) * n_u * n_s;
            s.kfb_stride = kfb_stride;
        }
        s.kff = d_k_ff + (c0 * horizon + t) * n_u;
        s.kff_stride = (long)horizon * n_u;
        s.sp = m->d_sp;
        s.p_out = d_p_all + c0 * hs + (long)t * n_s;
        s.p_out_stride = hs;
        s.q_out = d_q_all + c0 * hss + (long)t * n_s * n_s;
        s.q_out_stride = hss;
        s.var_out = d_var_all ? d_var_all + c0 * hs + (long)t * n_s : nullptr;
        s.var_out_stride = hs;
        s.status = d_status ? d_status + c0 : nullptr;
        s.b0 = b0;
        s.n_batch = b1;
        s.n_s = n_s;
        s.n_in = m->n_in;
        s.n_u = n_u;
        return s;
    };

    const bool guarded = m->ws_mode != 0 && tri_mode(m) >= 4 && tri_digits(m) == 4 && m->opt_guard != 0 &&
                         !(m->unguarded && m->opt_i8_digits == 0);
    // The serial schedule: per chunk and step kstar -> contraction (+ guard + recomputation) -> ellipsoid step, every
    // launch asynchronous on one stream.
    // Small models (few block rows): every kernel of a step is latency-bound and leaves most SMs idle (C2: ~30 us each
    // for 4096 trajectories).  The candidates are independent, so the chunk is cut into sub-batches of whole panel
    // pairs whose H-step chains run on their own internal streams and overlap on the device: same kernels, same
    // arithmetic per trajectory, disjoint parts of the workspace -- bit-identical results.  Measured at C2: 2 chains
    // 0.75 ms per call against 0.83 (one) and 0.82 (four: the contraction CTAs take a whole SM's shared memory each, so
    // more chains mostly queue) -- automatic mode uses two.  (Large models fill the GPU with every launch; there
    // co-resident kernels only get in each other's way: round 1's two-stream half-chunk pipeline ran the K* kernel 7x
    // slower next to the contraction, profiles/round1/overlap_pipeline_c3_c4_c5.txt, and was retired.)
    auto sub_batches = [&](long nb) -> int {
        if (m->opt_substreams == 0 || m->ws_mode == 0 || m->nblk > 8) return 1;
        const long np = (nb + I8_N - 1) / I8_N;
        const long want = m->opt_substreams > 0 ? m->opt_substreams : 2;
        return (int)std::max<long>(1, std::min<long>(want, np / 8));
    };
    auto issue_serial = [&](cudaStream_t s1) -> int {
        if (d_status != nullptr) SEGP_CUDA_CHECK(cudaMemsetAsync(d_status, 0, n_batch * sizeof(int32_t), s1));
        for (long c0 = 0; c0 < n_batch; c0 += m->b_cap) {
            const long nb = std::min<long>(m->b_cap, n_batch - c0);
            const int nsub = sub_batches(nb);
            if (nsub == 1) {
                for (int t = 0; t < horizon; ++t) {
                    SEGP_CHECK(run_kstar(m, kstar_args(c0, t, nb), s1));
                    SEGP_CHECK(run_tri(m, nb, s1));
                    SEGP_CHECK(launch_ellipsoid_step(step_args(c0, t, 0, nb), s1));
                    m->launches += 3;
                }
                continue;
            }
            if (m->s_sub[0] == nullptr) {
                for (int i = 0; i < 3; ++i) {
                    SEGP_CUDA_CHECK(cudaStreamCreateWithFlags(&m->s_sub[i], cudaStreamNonBlocking));
                    SEGP_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_sub[i], cudaEventDisableTiming));
                }
                SEGP_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_sub_fork, cudaEventDisableTiming));
            }
            const long np = (nb + I8_N - 1) / I8_N;
            const long q = 2 * ((np + 2 * nsub - 1) / (2 * nsub));   // panels per sub-batch: even (cluster pairs stay whole)
            SEGP_CUDA_CHECK(cudaEventRecord(m->ev_sub_fork, s1));
            for (int i = 0; i < nsub; ++i) {
                const long p0 = (long)i * q, p1 = std::min<long>(p0 + q, np);
                if (p0 >= p1) continue;
                cudaStream_t ss = i == 0 ? s1 : m->s_sub[i - 1];
                if (i > 0) SEGP_CUDA_CHECK(cudaStreamWaitEvent(ss, m->ev_sub_fork, 0));
                const long b0 = p0 * I8_N, b1 = std::min<long>(p1 * I8_N, nb);
                for (int t = 0; t < horizon; ++t) {
                    SEGP_CHECK(run_kstar(m, kstar_args(c0, t, b1), ss, (int)p0));
                    SEGP_CHECK(run_tri(m, b1, ss, (int)p0));
                    SEGP_CHECK(launch_ellipsoid_step(step_args(c0, t, b0, b1), ss));
                    m->launches += 3;
                }
                if (i > 0) {
                    SEGP_CUDA_CHECK(cudaEventRecord(m->ev_sub[i - 1], ss));
                    SEGP_CUDA_CHECK(cudaStreamWaitEvent(s1, m->ev_sub[i - 1], 0));
                }
            }
        }
        if (guarded)
            SEGP_CUDA_CHECK(cudaMemcpyAsync(m->h_fallback, m->fallback_counter, sizeof(unsigned int),
                                            cudaMemcpyDeviceToHost, s1));
        return SEGP_OK;
    };
    if (guarded) m->panels_prev_call = (long)horizon * ((n_batch + I8_N - 1) / I8_N);
    {
        // K5: the 3 H launches of a call are replayed as ONE CUDA graph from the second call with the same
        // arguments on (a sampling-MPC loop re-uses its buffers); sizes where a launch costs as much as a kernel
        // (C2: 30 launches of ~10 us of work) are otherwise launch-bound.  First sight of an argument set: direct
        // launches; second: stream capture on an internal stream + instantiate; then cudaGraphLaunch on the caller's.
        if (m->opt_graph == 0) return issue_serial(st);
        struct Key {
            long n_batch, p0_stride, q0_stride, kfb_stride, kfb_init_stride, b_cap, events_at;
            const void *p0, *q0, *kff, *kfb, *kfbi, *p_all, *q_all, *var_all, *status;
            int horizon, mode, digits, guard, nsplit, timed, pgroup, cluster, substreams;
        } key;
        memset(&key, 0, sizeof(key));
        key.n_batch = n_batch;
        key.p0_stride = p0_stride;
        key.q0_stride = q0_stride;
        key.kfb_stride = kfb_stride;
        key.kfb_init_stride = kfb_init_stride;
        key.b_cap = m->b_cap;
        key.events_at = m->time_tri ? (long)m->tri_events_used : -1;
        key.p0 = d_p0;
        key.q0 = d_q0;
        key.kff = d_k_ff;
        key.kfb = d_k_fb;
        key.kfbi = d_k_fb_init;
        key.p_all = d_p_all;
        key.q_all = d_q_all;
        key.var_all = d_var_all;
        key.status = d_status;
        key.horizon = horizon;
        key.mode = tri_mode(m);
        key.digits = tri_digits(m);
        key.guard = (int)m->opt_guard;
        key.nsplit = m->nsplit;
        key.timed = m->time_tri ? 1 : 0;
        key.pgroup = (int)m->opt_i8_panel_group;
        key.cluster = (int)m->opt_i8_cluster;
        key.substreams = (int)m->opt_substreams;
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
        GraphEntry* g = nullptr;
        int n_entries = 0;
        for (GraphEntry* e = m->graphs; e != nullptr; e = e->next, ++n_entries)
            if (e->key.size() == sizeof(key) && memcmp(e->key.data(), kb, sizeof(key)) == 0) g = e;
        if (g != nullptr && g->exec != nullptr) {
            SEGP_CUDA_CHECK(cudaGraphLaunch(g->exec, st));
            m->launches += g->launches;
            if (m->time_tri) m->tri_events_used = (size_t)key.events_at + (size_t)g->events;
            return SEGP_OK;
        }
        if (g == nullptr) {   // first sight: remember the arguments, launch directly
            if (n_entries >= 16) free_graphs(m);
            g = new (std::nothrow) GraphEntry();
            if (g != nullptr) {
                g->key.assign(kb, kb + sizeof(key));
                g->next = m->graphs;
                m->graphs = g;
            }
            return issue_serial(st);
        }
        // second sight: capture
        if (m->s_cap == nullptr) SEGP_CUDA_CHECK(cudaStreamCreateWithFlags(&m->s_cap, cudaStreamNonBlocking));
        const long launches0 = m->launches;
        const size_t events0 = m->tri_events_used;
        if (m->time_tri)   // events are created outside the capture
            while (m->tri_events.size() < events0 + 2 * (size_t)horizon * ((n_batch + m->b_cap - 1) / m->b_cap)) {
                cudaEvent_t e;
                SEGP_CUDA_CHECK(cudaEventCreate(&e));
                m->tri_events.push_back(e);
            }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(m->s_cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            return issue_serial(st);
        }
        const int rc_issue = issue_serial(m->s_cap);
        const cudaError_t ec = cudaStreamEndCapture(m->s_cap, &graph);
        const long captured = m->launches - launches0;
        const size_t captured_events = m->tri_events_used - events0;
        m->launches = launches0;
        m->tri_events_used = events0;
        cudaGraphExec_t exec = nullptr;
        if (rc_issue != SEGP_OK || ec != cudaSuccess || graph == nullptr ||
            cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph != nullptr) cudaGraphDestroy(graph);
            m->opt_graph = 0;   // this configuration cannot be captured: stay on direct launches
            return issue_serial(st);
        }
        cudaGraphDestroy(graph);
        g->exec = exec;
        g->launches = captured;
        g->events = (long)captured_events;
        SEGP_CUDA_CHECK(cudaGraphLaunch(g->exec, st));
        m->launches += g->launches;
        m->tri_events_used = events0 + captured_events;
        return SEGP_OK;
    }

}

// ---------------------------------------------------------------------------------------------- host entry
static int ensure_stage(segp_model* m, size_t bytes) {
    if (bytes <= m->stage_bytes) return SEGP_OK;
    if (m->stage != nullptr) cudaFree(m->stage);
    m->stage = nullptr;
    m->stage_bytes = 0;
    SEGP_CUDA_CHECK(cudaMalloc(&m->stage, bytes));
    m->stage_bytes = bytes;
    return SEGP_OK;
}

int segp_multistep_host(segp_model* m, long n_batch, int horizon, const double* h_p0, long p0_stride,
                        const double* h_q0, long q0_stride, const double* h_k_ff, const double* h_k_fb,
                        long kfb_stride, const double* h_k_fb_init, long kfb_init_stride,
                        const segp_reach_params* params, double* h_p_all, double* h_q_all, double* h_var_all,
                        int32_t* h_status) {
    SEGP_CHECK(check_ready(m));
    if (n_batch < 0 || horizon < 1) {
        set_error("segp_multistep_host: n_batch >= 0 and horizon >= 1 required");
        return SEGP_ERR_INVALID;
    }
    if (n_batch == 0) return SEGP_OK;
    if (h_p0 == nullptr || h_k_ff == nullptr || h_p_all == nullptr || h_q_all == nullptr) {
        set_error("segp_multistep_host: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(m->device);
    const int n_s = m->n_s, n_u = m->n_u;
    auto al = [](size_t v) { return (v + 31) / 32 * 32; };   // in doubles, keeps 256-byte alignment
    const size_t n_p0 = p0_stride ? (size_t)n_batch * n_s : n_s;
    const size_t n_q0 = h_q0 ? (q0_stride ? (size_t)n_batch * n_s * n_s : (size_t)n_s * n_s) : 0;
    const size_t n_kff = (size_t)n_batch * horizon * n_u;
    const size_t n_kfb1 = (size_t)std::max(horizon - 1, 0) * n_u * n_s;
    const size_t n_kfb = h_k_fb ? (kfb_stride ? (size_t)n_batch * n_kfb1 : n_kfb1) : 0;
    const size_t n_kfbi = h_k_fb_init ? (kfb_init_stride ? (size_t)n_batch * n_u * n_s : (size_t)n_u * n_s) : 0;
    const size_t n_pall = (size_t)n_batch * horizon * n_s;
    const size_t n_qall = n_pall * n_s;
    const size_t n_stat = (size_t)(n_batch + 1) / 2;   // int32 pairs in double units
    size_t off = 0;
    auto take = [&](size_t n) {
        const size_t o = off;
        off += al(std::max<size_t>(n, 1));
        return o;
    };
    const size_t o_p0 = take(n_p0), o_q0 = take(n_q0), o_kff = take(n_kff), o_kfb = take(n_kfb), o_kfbi = take(n_kfbi),
                 o_pall = take(n_pall), o_qall = take(n_qall), o_var = take(n_pall), o_stat = take(n_stat);
    SEGP_CHECK(ensure_stage(m, off * sizeof(double)));
    double* base = static_cast<double*>(m->stage);
    // Own streams (not the legacy default stream): compute on s_host, result copies on s_copy, so the device-to-host
    // copy of chunk i runs under the rollout of chunk i + 1 when the batch exceeds one workspace chunk.
    if (m->s_host == nullptr) {
        SEGP_CUDA_CHECK(cudaStreamCreateWithFlags(&m->s_host, cudaStreamNonBlocking));
        SEGP_CUDA_CHECK(cudaStreamCreateWithFlags(&m->s_copy, cudaStreamNonBlocking));
        SEGP_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_chunk, cudaEventDisableTiming));
    }
    cudaStream_t st = m->s_host;
    auto up = [&](size_t o, const double* src, size_t n) -> cudaError_t {
        if (n == 0 || src == nullptr) return cudaSuccess;
        return cudaMemcpyAsync(base + o, src, n * sizeof(double), cudaMemcpyHostToDevice, st);
    };
    SEGP_CUDA_CHECK(up(o_p0, h_p0, n_p0));
    SEGP_CUDA_CHECK(up(o_q0, h_q0, n_q0));
    SEGP_CUDA_CHECK(up(o_kff, h_k_ff, n_kff));
    SEGP_CUDA_CHECK(up(o_kfb, h_k_fb, n_kfb));
    SEGP_CUDA_CHECK(up(o_kfbi, h_k_fb_init, n_kfbi));
    int32_t* d_stat = reinterpret_cast<int32_t*>(base + o_stat);
    const long chunk = std::max<long>(1, std::min<long>(m->opt_chunk, n_batch));
    for (long c0 = 0; c0 < n_batch; c0 += chunk) {
        const long nb = std::min<long>(chunk, n_batch - c0);
        const size_t hs = (size_t)horizon * n_s;
        SEGP_CHECK(segp_multistep(m, nb, horizon, base + o_p0 + c0 * p0_stride, p0_stride,
                                  h_q0 ? base + o_q0 + c0 * q0_stride : nullptr, q0_stride,
                                  base + o_kff + (size_t)c0 * horizon * n_u, h_k_fb ? base + o_kfb + c0 * kfb_stride : nullptr,
                                  kfb_stride, h_k_fb_init ? base + o_kfbi + c0 * kfb_init_stride : nullptr, kfb_init_stride,
                                  params, base + o_pall + c0 * hs, base + o_qall + c0 * hs * n_s,
                                  h_var_all ? base + o_var + c0 * hs : nullptr, h_status ? d_stat + c0 : nullptr, st));
        SEGP_CUDA_CHECK(cudaEventRecord(m->ev_chunk, st));
        SEGP_CUDA_CHECK(cudaStreamWaitEvent(m->s_copy, m->ev_chunk, 0));
        cudaStream_t sc = m->s_copy;
        SEGP_CUDA_CHECK(cudaMemcpyAsync(h_p_all + c0 * hs, base + o_pall + c0 * hs, nb * hs * sizeof(double),
                                        cudaMemcpyDeviceToHost, sc));
        SEGP_CUDA_CHECK(cudaMemcpyAsync(h_q_all + c0 * hs * n_s, base + o_qall + c0 * hs * n_s, nb * hs * n_s * sizeof(double),
                                        cudaMemcpyDeviceToHost, sc));
        if (h_var_all != nullptr)
            SEGP_CUDA_CHECK(cudaMemcpyAsync(h_var_all + c0 * hs, base + o_var + c0 * hs, nb * hs * sizeof(double),
                                            cudaMemcpyDeviceToHost, sc));
        if (h_status != nullptr)
            SEGP_CUDA_CHECK(cudaMemcpyAsync(h_status + c0, d_stat + c0, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, sc));
    }
    SEGP_CUDA_CHECK(cudaStreamSynchronize(m->s_copy));
    SEGP_CUDA_CHECK(cudaStreamSynchronize(st));
    return SEGP_OK;
}

// ---------------------------------------------------------------------------------------------- handle-free leaves
namespace {
struct TempParams {
    StepParams* d = nullptr;
    cudaStream_t st;
    explicit TempParams(cudaStream_t s) : st(s) {}
    int upload(const StepParams& sp) {
        retain_small_async_allocations();
        SEGP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d), sizeof(StepParams), st));
        SEGP_CUDA_CHECK(cudaMemcpyAsync(d, &sp, sizeof(sp), cudaMemcpyHostToDevice, st));
        return SEGP_OK;
    }
    ~TempParams() {
        if (d != nullptr) cudaFreeAsync(d, st);
    }
};
int check_dims(int n_s, int n_in, int n_u) {
    if (n_s < 1 || n_s > SEGP_MAX_NS || n_in < 1 || n_in > SEGP_MAX_NS || n_u < 0 || n_u > SEGP_MAX_NU) {
        set_error("dimensions out of range (n_s=%d n_in=%d n_u=%d; limits %d/%d)", n_s, n_in, n_u, SEGP_MAX_NS,
                  SEGP_MAX_NU);
        return SEGP_ERR_INVALID;
    }
    return SEGP_OK;
}
}  // namespace

int segp_ellipsoid_step(int device, long n_batch, int n_s, int n_s_in, int n_u, const double* d_mu,
                        const double* d_var, const double* d_jac, const double* d_p, const double* d_q,
                        const double* d_k_ff, const double* d_k_fb, long kfb_stride, const segp_reach_params* params,
                        double* d_p_out, double* d_q_out, int32_t* d_status, void* stream) {
    SEGP_CHECK(check_dims(n_s, n_s_in, n_u));
    if (n_batch <= 0) return n_batch == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (d_mu == nullptr || d_var == nullptr || d_p == nullptr || d_k_ff == nullptr || d_p_out == nullptr ||
        d_q_out == nullptr || (d_q != nullptr && (d_jac == nullptr || d_k_fb == nullptr))) {
        set_error("segp_ellipsoid_step: null buffer");
        return SEGP_ERR_INVALID;
    }
    StepParams sp;
    SEGP_CHECK(fill_step_params(&sp, params, n_s, n_s_in, n_u));
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TempParams tp(st);
    SEGP_CHECK(tp.upload(sp));
    if (d_status != nullptr) SEGP_CUDA_CHECK(cudaMemsetAsync(d_status, 0, n_batch * sizeof(int32_t), st));
    StepArgs s{};
    s.mu_d = d_mu;
    s.var_d = d_var;
    s.jac_d = d_jac;
    s.p = d_p;
    s.p_stride = n_s;
    s.q = d_q;
    s.q_stride = (long)n_s * n_s;
    s.kff = d_k_ff;
    s.kff_stride = n_u;
    s.kfb = d_k_fb;
    s.kfb_stride = kfb_stride;
    s.sp = tp.d;
    s.p_out = d_p_out;
    s.p_out_stride = n_s;
    s.q_out = d_q_out;
    s.q_out_stride = (long)n_s * n_s;
    s.status = d_status;
    s.n_batch = n_batch;
    s.n_s = n_s;
    s.n_in = n_s_in;
    s.n_u = n_u;
    return launch_ellipsoid_step(s, st);
}

int segp_remainder_overapproximations(int device, long n_batch, int n_s, int n_u, const double* d_q,
                                      const double* d_k_fb, long kfb_stride, const double* h_l_mu,
                                      const double* h_l_sigma, double* d_u_mu, double* d_u_sigma, void* stream) {
    SEGP_CHECK(check_dims(n_s, n_s, n_u));
    if (n_batch <= 0) return n_batch == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (d_q == nullptr || d_k_fb == nullptr || h_l_mu == nullptr || h_l_sigma == nullptr || d_u_mu == nullptr ||
        d_u_sigma == nullptr) {
        set_error("segp_remainder_overapproximations: null buffer");
        return SEGP_ERR_INVALID;
    }
    segp_reach_params prm{h_l_mu, h_l_sigma, 1.0, nullptr, nullptr, nullptr, SEGP_PROP_ELLIPSOID};
    StepParams sp;
    SEGP_CHECK(fill_step_params(&sp, &prm, n_s, n_s, n_u));
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TempParams tp(st);
    SEGP_CHECK(tp.upload(sp));
    return launch_remainder(n_batch, n_s, n_u, d_q, d_k_fb, kfb_stride, tp.d, d_u_mu, d_u_sigma, st);
}

int segp_sum_two_ellipsoids(int device, long n_batch, int n, const double* d_p1, const double* d_q1,
                            const double* d_p2, const double* d_q2, double* d_p, double* d_q, void* stream) {
    if (n_batch <= 0) return n_batch == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (n < 1 || d_p1 == nullptr || d_q1 == nullptr || d_p2 == nullptr || d_q2 == nullptr || d_p == nullptr ||
        d_q == nullptr) {
        set_error("segp_sum_two_ellipsoids: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    return launch_sum_two(n_batch, n, d_p1, d_q1, d_p2, d_q2, d_p, d_q, static_cast<cudaStream_t>(stream));
}

int segp_ellipsoid_from_rectangle(int device, long n_batch, int n, const double* d_ub, double* d_q, int32_t* d_status,
                                  void* stream) {
    if (n_batch <= 0) return n_batch == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (n < 1 || d_ub == nullptr || d_q == nullptr) {
        set_error("segp_ellipsoid_from_rectangle: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    return launch_from_rectangle(n_batch, n, d_ub, d_q, d_status, static_cast<cudaStream_t>(stream));
}

int segp_safety_distance(int device, long n_items, int n_s, int m, const double* d_p, const double* d_q,
                         const double* h_h_mat, const double* h_h_vec, double c_safety, double* d_dist, void* stream) {
    if (n_items <= 0) return n_items == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (n_s < 1 || m < 1 || d_p == nullptr || d_q == nullptr || h_h_mat == nullptr || h_h_vec == nullptr ||
        d_dist == nullptr) {
        set_error("segp_safety_distance: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* d_h = nullptr;
    const size_t n_h = (size_t)m * n_s + m;
    retain_small_async_allocations();
    SEGP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_h), n_h * sizeof(double), st));
    cudaError_t e = cudaMemcpyAsync(d_h, h_h_mat, (size_t)m * n_s * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_h + (size_t)m * n_s, h_h_vec, m * sizeof(double), cudaMemcpyHostToDevice, st);
    int rc = SEGP_OK;
    if (e != cudaSuccess) {
        set_error("segp_safety_distance: %s", cudaGetErrorString(e));
        rc = SEGP_ERR_CUDA;
    } else {
        rc = launch_safety_distance(n_items, n_s, m, d_p, d_q, d_h, d_h + (size_t)m * n_s, c_safety, d_dist, st);
    }
    cudaFreeAsync(d_h, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------- scoring
static int fill_score_params(ScoreParams* sp, const segp_score_params* prm, int n_s, int n_u) {
    memset(sp, 0, sizeof(*sp));
    if (prm == nullptr) {
        set_error("score params: null");
        return SEGP_ERR_INVALID;
    }
    if (prm->m_obs < 0 || prm->m_obs > SEGP_MAX_CONSTR || prm->m_safe < 0 || prm->m_safe > SEGP_MAX_CONSTR) {
        set_error("score params: polytopes are limited to %d rows (m_obs=%d, m_safe=%d)", SEGP_MAX_CONSTR, prm->m_obs,
                  prm->m_safe);
        return SEGP_ERR_INVALID;
    }
    if ((prm->m_obs > 0 && (prm->h_mat_obs == nullptr || prm->h_obs == nullptr)) ||
        (prm->m_safe > 0 && (prm->h_mat_safe == nullptr || prm->h_safe == nullptr)) ||
        ((prm->h_u_min == nullptr) != (prm->h_u_max == nullptr))) {
        set_error("score params: inconsistent null pointers");
        return SEGP_ERR_INVALID;
    }
    sp->has_ctrl = prm->h_u_min != nullptr;
    for (int j = 0; j < n_u && sp->has_ctrl; ++j) {
        sp->u_min[j] = prm->h_u_min[j];
        sp->u_max[j] = prm->h_u_max[j];
    }
    sp->m_obs = prm->m_obs;
    sp->m_safe = prm->m_safe;
    for (int i = 0; i < prm->m_obs * n_s; ++i) sp->h_mat_obs[i] = prm->h_mat_obs[i];
    for (int i = 0; i < prm->m_obs; ++i) sp->h_obs[i] = prm->h_obs[i];
    for (int i = 0; i < prm->m_safe * n_s; ++i) sp->h_mat_safe[i] = prm->h_mat_safe[i];
    for (int i = 0; i < prm->m_safe; ++i) sp->h_safe[i] = prm->h_safe[i];
    sp->c_safety = prm->c_safety;
    sp->eps_constraints = prm->eps_constraints;
    sp->eps_noise = prm->eps_noise;
    sp->cost_type = prm->cost_type;
    if (prm->layout != SEGP_SCORE_SAFEMPC && prm->layout != SEGP_SCORE_CAUTIOUS) {
        set_error("score params: unknown constraint layout %d", prm->layout);
        return SEGP_ERR_INVALID;
    }
    sp->layout = prm->layout;
    if (sp->layout == SEGP_SCORE_CAUTIOUS) sp->m_safe = 0;
    if ((prm->h_q0 == nullptr) != (prm->h_k_fb_0 == nullptr)) {
        set_error("score params: h_q0 and h_k_fb_0 go together");
        return SEGP_ERR_INVALID;
    }
    sp->has_q0 = prm->h_q0 != nullptr;
    if (sp->has_q0) {
        for (int i = 0; i < n_s * n_s; ++i) sp->q0[i] = prm->h_q0[i];
        for (int i = 0; i < n_u * n_s; ++i) sp->kfb0[i] = prm->h_k_fb_0[i];
    }
    if (prm->cost_type == SEGP_COST_QUADRATIC) {
        if (prm->h_wx == nullptr || prm->h_wu == nullptr) {
            set_error("score params: the quadratic cost needs h_wx and h_wu");
            return SEGP_ERR_INVALID;
        }
        for (int i = 0; i < n_s * n_s; ++i) sp->wx[i] = prm->h_wx[i];
        for (int i = 0; i < n_u * n_u; ++i) sp->wu[i] = prm->h_wu[i];
        for (int i = 0; i < n_s; ++i) sp->x_ref[i] = prm->h_x_ref ? prm->h_x_ref[i] : 0.0;
    } else if (prm->cost_type != SEGP_COST_EXPLORATION) {
        set_error("score params: unknown cost type %d", prm->cost_type);
        return SEGP_ERR_INVALID;
    }
    return SEGP_OK;
}

int segp_score_num_constraints(int horizon, int n_u, const segp_score_params* params) {
    if (params == nullptr || horizon < 1) return -1;
    const int ctrl = params->h_u_min != nullptr ? 2 * n_u * horizon : 0;
    if (params->layout == SEGP_SCORE_CAUTIOUS) return ctrl + horizon * params->m_obs;
    return ctrl + (horizon - 1) * params->m_obs + params->m_safe;
}

int segp_score_rollouts(int device, long n_batch, int horizon, int n_s, int n_u, const double* d_p_all,
                        const double* d_q_all, const double* d_var_all, const double* d_k_ff, const double* d_k_fb,
                        long kfb_stride, const int32_t* d_status, const segp_score_params* params, double* d_cost,
                        int32_t* d_feasible, double* d_violation, double* d_g, void* stream) {
    SEGP_CHECK(check_dims(n_s, n_s, n_u));
    if (n_batch <= 0 || horizon < 1) return (n_batch == 0 && horizon >= 1) ? SEGP_OK : SEGP_ERR_INVALID;
    ScoreParams sp;
    SEGP_CHECK(fill_score_params(&sp, params, n_s, n_u));
    if (d_p_all == nullptr || d_q_all == nullptr || d_k_ff == nullptr || d_cost == nullptr || d_feasible == nullptr ||
        d_violation == nullptr || (sp.cost_type == SEGP_COST_EXPLORATION && d_var_all == nullptr) ||
        (sp.has_ctrl && horizon > 1 && d_k_fb == nullptr)) {
        set_error("segp_score_rollouts: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScoreParams* d_sp = nullptr;
    retain_small_async_allocations();
    SEGP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_sp), sizeof(ScoreParams), st));
    cudaError_t e = cudaMemcpyAsync(d_sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, st);
    int rc = SEGP_OK;
    if (e != cudaSuccess) {
        set_error("segp_score_rollouts: %s", cudaGetErrorString(e));
        rc = SEGP_ERR_CUDA;
    } else {
        // the host copy `sp` must outlive the asynchronous upload: pageable memory is staged by the runtime before
        // cudaMemcpyAsync returns, so this is safe
        ScoreArgs a{};
        a.p_all = d_p_all;
        a.q_all = d_q_all;
        a.var_all = d_var_all;
        a.kff = d_k_ff;
        a.kfb = d_k_fb;
        a.kfb_stride = kfb_stride;
        a.status = d_status;
        a.sp = d_sp;
        a.cost = d_cost;
        a.feasible = d_feasible;
        a.violation = d_violation;
        a.g = d_g;
        a.n_batch = n_batch;
        a.horizon = horizon;
        a.n_s = n_s;
        a.n_u = n_u;
        a.n_g = segp_score_num_constraints(horizon, n_u, params);
        rc = launch_score(a, st);
    }
    cudaFreeAsync(d_sp, st);
    return rc;
}

int segp_argbest(int device, long n_batch, const double* d_cost, const int32_t* d_feasible, const double* d_violation,
                 long* h_index, double* h_cost, double* h_violation, int* h_feasible, void* stream) {
    if (h_index == nullptr) {
        set_error("segp_argbest: null output");
        return SEGP_ERR_INVALID;
    }
    *h_index = -1;
    if (h_feasible) *h_feasible = 0;
    if (n_batch <= 0) return n_batch == 0 ? SEGP_OK : SEGP_ERR_INVALID;
    if (d_cost == nullptr || d_feasible == nullptr || d_violation == nullptr) {
        set_error("segp_argbest: null buffer");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BestCandidate* d_out = nullptr;
    retain_small_async_allocations();
    SEGP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_out), sizeof(BestCandidate), st));
    int rc = launch_argbest(n_batch, d_cost, d_feasible, d_violation, d_out, st);
    BestCandidate out{};
    if (rc == SEGP_OK) {
        cudaError_t e = cudaMemcpyAsync(&out, d_out, sizeof(out), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("segp_argbest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
        }
    }
    cudaFreeAsync(d_out, st);
    if (rc == SEGP_OK) {
        *h_index = out.index;
        if (h_cost) *h_cost = out.cost;
        if (h_violation) *h_violation = out.violation;
        if (h_feasible) *h_feasible = out.feasible;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------- tcgen05 diagnostics
int segp_i8_peak(int device, int umma_n, int iters, double* tops) {
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_i8_peak: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    return i8_peak(umma_n, iters, 0, tops);
}

int segp_i8_peak_pattern(int device, int umma_n, int pattern, int iters, double* tops) {
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_i8_peak_pattern: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    return i8_peak(umma_n, iters, pattern, tops);
}

int segp_i8_gemm_selftest(int device, int m, int n, int k, const double* h_a, const double* h_b, double* h_c, double alpha,
                          double beta, int trans_b, int flags) {
    if (m < 1 || n < 1 || k < 1 || h_a == nullptr || h_b == nullptr || h_c == nullptr) {
        set_error("segp_i8_gemm_selftest: bad argument");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_i8_gemm_selftest: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    return gemm_i8d_selftest(m, n, k, h_a, h_b, h_c, alpha, beta, trans_b, flags);
}

int segp_i8_selftest(int device, int variant, int k_blocks, const int8_t* h_a, const int8_t* h_b, int32_t* h_acc,
                     double* h_colsum) {
    // variant 1: reference kernel tri_i8; 4: tri_i8m on the classic set; 6: tri_i8m on the diagonal-split set -- all on
    // the last block row of a (128 k_blocks)-point model, one panel.  h_a holds the planes the kernel multiplies: 5
    // (variants 1, 4) or 5 with plane 0 = the diagonal's leading digit (variant 6: non-zero on the diagonal only).
    if ((variant != 1 && variant != 4 && variant != 6) || k_blocks < 1 || k_blocks > 64 || h_a == nullptr ||
        h_b == nullptr || h_acc == nullptr || h_colsum == nullptr) {
        set_error("segp_i8_selftest: bad argument (variant 1|4|6, k_blocks in [1,64])");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_i8_selftest: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    SEGP_CHECK(tri_i8_init());
    const int nblk = k_blocks, kdim = TILE * k_blocks, nkb = 2 * k_blocks;
    const int bi = k_blocks - 1;
    const bool split = variant == 6;
    const int na = split ? I8_SS : I8_S;
    const size_t a_bytes = (size_t)nblk * (nblk + 1) * (na * I8_A_TILE);
    const size_t m1_bytes = (size_t)nblk * 2 * I8_A_TILE;
    const size_t b_bytes = (size_t)2 * nkb * (I8_S * I8_B_TILE);   // two panels (the cluster kernels work on pairs)
    std::vector<int8_t> a_img(a_bytes, 0), m1_img(m1_bytes, 0), b_img(b_bytes, 0);
    auto sw = [](int r, int k) { return r * I8_KB + ((((k >> 4) ^ ((r >> 1) & 3)) << 4) | (k & 15)); };
    // same image formats as pack_w_i8_kernel / kstar_i8_kernel
    for (int kb = 0; kb < 2 * (bi + 1); ++kb)
        for (int pl = 0; pl < I8_S; ++pl) {
            int8_t* at;
            if (!split)
                at = a_img.data() + ((size_t)bi * (bi + 1) + kb) * (I8_S * I8_A_TILE) + (size_t)pl * I8_A_TILE;
            else if (pl == 0) {
                if (kb < 2 * bi) continue;
                at = m1_img.data() + ((size_t)bi * 2 + (kb - 2 * bi)) * I8_A_TILE;
            } else
                at = a_img.data() + ((size_t)bi * (bi + 1) + kb) * (I8_SS * I8_A_TILE) + (size_t)(pl - 1) * I8_A_TILE;
            for (int k = 0; k < I8_KB; ++k)
                for (int r = 0; r < TILE; ++r) at[sw(r, k)] = h_a[((size_t)pl * TILE + r) * kdim + (size_t)kb * I8_KB + k];
        }
    for (int kb = 0; kb < nkb; ++kb)
        for (int pl = 0; pl < I8_S; ++pl)
            for (int k = 0; k < I8_KB; ++k)
                for (int r = 0; r < I8_N; ++r)
                    b_img[(size_t)kb * (I8_S * I8_B_TILE) + (size_t)pl * I8_B_TILE + sw(r, k)] =
                        h_b[((size_t)pl * I8_N + r) * kdim + (size_t)kb * I8_KB + k];
    int8_t *d_a = nullptr, *d_m1 = nullptr, *d_b = nullptr;
    double *d_rf = nullptr, *d_q = nullptr;
    int32_t* d_dbg = nullptr;
    const size_t n_dbg = (size_t)I8_S * TILE * I8_N;
    int rc = SEGP_OK;
    do {
        if ((rc = dev_alloc(&d_a, a_bytes)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_m1, m1_bytes)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_b, b_bytes)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_rf, (size_t)kdim)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_q, (size_t)nblk * 2 * I8_N)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_dbg, n_dbg)) != SEGP_OK) break;
        std::vector<double> ones((size_t)kdim, 1.0);
        cudaError_t e = cudaMemcpy(d_a, a_img.data(), a_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_m1, m1_img.data(), m1_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_b, b_img.data(), b_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_rf, ones.data(), kdim * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(d_q, 0, (size_t)nblk * 2 * I8_N * sizeof(double));
        if (e == cudaSuccess) e = cudaMemset(d_dbg, 0, n_dbg * sizeof(int32_t));
        if (e != cudaSuccess) {
            set_error("segp_i8_selftest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
            break;
        }
        TriI8Args t{};
        t.wi8 = d_a;
        t.wm1 = d_m1;
        t.rowfac = d_rf;
        t.ki8 = d_b;
        t.qpart = d_q;
        t.nblk = nblk;
        t.npanels = 1;
        t.npanel_cap = 2;
        t.b_cap = 2 * I8_N;
        t.digits = split ? 4 : 5;
        t.cluster = 2;
        if (variant == 1) {
            t.dbg = d_dbg;
            t.fix_bi = bi;
            rc = launch_tri_i8(t, 1, nullptr);
        } else {
            t.fix_bi = -1;
            rc = launch_tri_i8m(t, 1, nullptr);   // all block rows of the (zero elsewhere) model; row bi is checked
        }
        if (rc != SEGP_OK) break;
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess && variant == 1) e = cudaMemcpy(h_acc, d_dbg, n_dbg * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)   // column sums of the block row: [96]
            e = cudaMemcpy(h_colsum, d_q + (size_t)bi * 2 * I8_N, (size_t)I8_N * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("segp_i8_selftest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
        }
    } while (0);
    dev_free(d_a);
    dev_free(d_m1);
    dev_free(d_b);
    dev_free(d_rf);
    dev_free(d_q);
    dev_free(d_dbg);
    return rc;
}

// ---------------------------------------------------------------------------------------------- options
int segp_set_option(segp_model* m, const char* name, long value) {
    if (m == nullptr || name == nullptr) {
        set_error("segp_set_option: null argument");
        return SEGP_ERR_INVALID;
    }
    if (strcmp(name, "chunk") == 0 && value >= TILE) {
        DeviceGuard guard(m->device);
        cudaDeviceSynchronize();
        free_workspace(m);
        m->opt_chunk = (value + TILE - 1) / TILE * TILE;
        return SEGP_OK;
    }
    if (strcmp(name, "panel_group") == 0 && value >= 1) {
        m->opt_panel_group = value;
        return SEGP_OK;
    }
    if (strcmp(name, "i8_panel_group") == 0 && value >= 0 && value <= 256 && value % 2 == 0) {
        m->opt_i8_panel_group = value;
        return SEGP_OK;
    }
    if (strcmp(name, "i8_cluster") == 0 && (value == 2 || value == 4)) {
        m->opt_i8_cluster = value;
        return SEGP_OK;
    }
    if (strcmp(name, "ksplit") == 0 && value >= 0) {
        m->opt_ksplit = value;
        return SEGP_OK;
    }
    if (strcmp(name, "tri_mode") == 0 && (value == -1 || value == 0 || value == 1 || value == 4 || value == 5)) {
        if (value >= 1 && m->has_data && !i8_capable(m)) {
            set_error("tri_mode=%ld (int8 tcgen05) needs n_train_padded <= %ld; this model has %d", value, I8_MAX_NPAD,
                      m->n_pad);
            return SEGP_ERR_UNSUPPORTED;
        }
        if (value != m->opt_tri_mode) {
            DeviceGuard guard(m->device);
            cudaDeviceSynchronize();
            free_graphs(m);
        }
        m->opt_tri_mode = value;
        return SEGP_OK;
    }
    if (strcmp(name, "i8_digits") == 0 && (value == 0 || value == 4 || value == 5)) {
        m->opt_i8_digits = value;
        return SEGP_OK;
    }
    if (strcmp(name, "guard") == 0 && (value == 0 || value == 1)) {
        m->opt_guard = value;
        return SEGP_OK;
    }
    if (strcmp(name, "probe") == 0 && (value == 0 || value == 1)) {   // takes effect at the next segp_factorize
        m->opt_probe = value;
        return SEGP_OK;
    }
    if (strcmp(name, "keep_fp64") == 0 && (value == 0 || value == 1)) {   // takes effect at the next segp_factorize
        m->opt_keep_fp64 = value;
        return SEGP_OK;
    }
    if (strcmp(name, "graph") == 0 && (value == 0 || value == 1)) {
        m->opt_graph = value;
        return SEGP_OK;
    }
    if (strcmp(name, "substreams") == 0 && value >= -1 && value <= 4) {
        DeviceGuard guard(m->device);
        cudaDeviceSynchronize();
        free_graphs(m);
        m->opt_substreams = value;
        return SEGP_OK;
    }
    if (strcmp(name, "scratch_cache") == 0 && value >= -1 && value <= 1) {
        m->opt_scratch_cache = value;
        if (value == 0) {
            DeviceGuard guard(m->device);
            cudaDeviceSynchronize();
            free_fact_cache(m);
        }
        return SEGP_OK;
    }
    if (strcmp(name, "fact_i8") == 0 && value >= -1 && value <= 1) {   // takes effect at the next segp_factorize
        m->opt_fact_i8 = value;
        return SEGP_OK;
    }
    if (strcmp(name, "keep_w") == 0 && (value == 0 || value == 1)) {   // takes effect at the next segp_factorize
        m->opt_keep_w = value;
        if (value == 0) {
            DeviceGuard guard(m->device);
            cudaDeviceSynchronize();
            dev_free(m->wdense);
        }
        return SEGP_OK;
    }
    if (strcmp(name, "i8_prof") == 0) {   // 1: allocate the counter buffer; 0: release it
        DeviceGuard guard(m->device);
        if (value != 0 && m->i8_prof == nullptr) {
            SEGP_CHECK(dev_alloc(&m->i8_prof, (size_t)128 * 8));
            SEGP_CUDA_CHECK(cudaMemset(m->i8_prof, 0, 128 * 8 * sizeof(long long)));
        } else if (value == 0) {
            cudaDeviceSynchronize();
            dev_free(m->i8_prof);
        }
        return SEGP_OK;
    }
    if (strcmp(name, "time_tri") == 0) {
        m->time_tri = value != 0;
        if (m->time_tri) m->tri_events_used = 0;
        return SEGP_OK;
    }
    set_error("segp_set_option: unknown option or bad value: %s=%ld", name, value);
    return SEGP_ERR_INVALID;
}

int segp_get_option(segp_model* m, const char* name, long* value) {
    if (m == nullptr || name == nullptr || value == nullptr) {
        set_error("segp_get_option: null argument");
        return SEGP_ERR_INVALID;
    }
    if (strcmp(name, "chunk") == 0) *value = m->opt_chunk;
    else if (strcmp(name, "panel_group") == 0) *value = m->opt_panel_group;
    else if (strcmp(name, "ksplit") == 0) *value = m->opt_ksplit;
    else if (strcmp(name, "i8_panel_group") == 0) *value = m->opt_i8_panel_group;
    else if (strcmp(name, "i8_cluster") == 0) *value = m->opt_i8_cluster;
    else if (strcmp(name, "tri_mode") == 0) *value = m->opt_tri_mode;
    else if (strcmp(name, "i8_prof_ptr") == 0) *value = (long)(uintptr_t)m->i8_prof;
    else if (strcmp(name, "tri_mode_effective") == 0) *value = tri_mode(m);
    else if (strcmp(name, "i8_digits") == 0) *value = m->opt_i8_digits;
    else if (strcmp(name, "i8_digits_effective") == 0) *value = tri_mode(m) == 0 ? 0 : tri_digits(m);
    else if (strcmp(name, "guard") == 0) *value = m->opt_guard;
    else if (strcmp(name, "probe") == 0) *value = m->opt_probe;
    else if (strcmp(name, "keep_fp64") == 0) *value = m->opt_keep_fp64;
    else if (strcmp(name, "graph") == 0) *value = m->opt_graph;
    else if (strcmp(name, "substreams") == 0) *value = m->opt_substreams;
    else if (strcmp(name, "graphs_cached") == 0) {
        long n = 0;
        for (GraphEntry* e = m->graphs; e != nullptr; e = e->next) n += e->exec != nullptr;
        *value = n;
    }
    else if (strcmp(name, "demoted") == 0) *value = m->demoted ? 1 : 0;
    else if (strcmp(name, "unguarded") == 0) *value = m->unguarded ? 1 : 0;
    else if (strcmp(name, "fp64_operand_resident") == 0) *value = m->wt != nullptr ? 1 : 0;
    else if (strcmp(name, "fp64_operand_needed") == 0) *value = fp64_operand_needed(m) ? 1 : 0;
    else if (strcmp(name, "factor_bytes") == 0) *value = (long)m->arena_bytes;
    else if (strcmp(name, "fallback_panels") == 0) {
        // panels recomputed on the 15-product digit set since the handle was created
        DeviceGuard guard(m->device);
        unsigned int v = 0;
        SEGP_CUDA_CHECK(cudaDeviceSynchronize());
        SEGP_CUDA_CHECK(cudaMemcpy(&v, m->fallback_counter, sizeof(v), cudaMemcpyDeviceToHost));
        *value = (long)v;
    }
    else if (strcmp(name, "tri_persistent") == 0) *value = m->last_tri_persistent ? 1 : 0;
    else if (strcmp(name, "append_incremental") == 0) *value = m->last_append_incremental ? 1 : 0;
    else if (strcmp(name, "keep_w") == 0) *value = m->opt_keep_w;
    else if (strcmp(name, "fact_i8") == 0) *value = m->opt_fact_i8;
    else if (strcmp(name, "scratch_cache") == 0) *value = m->opt_scratch_cache;
    else if (strcmp(name, "scratch_cached_bytes") == 0) {
        *value = 0;
        if (m->fact_cache_slots > 0) {
            const size_t nn = (size_t)m->fact_cache_npad * m->fact_cache_npad;
            size_t b = ((m->fact_cache_w ? 3 : 2) * nn + nn / NBLK * NBLK + 33 * (size_t)m->fact_cache_npad) * sizeof(double);
            if (m->fact_cache_fd) b += 2 * fd_scratch_plane_bytes(m->fact_cache_npad) + 2 * m->fact_cache_npad * sizeof(double);
            *value = (long)(b * m->fact_cache_slots);
        }
    }
    else if (strcmp(name, "fact_i8_effective") == 0) *value = m->last_fact_i8 ? 1 : 0;
    else if (strcmp(name, "n_train") == 0) *value = m->n_train;
    else if (strcmp(name, "factorized") == 0) *value = m->factorized ? 1 : 0;
    else if (strcmp(name, "launches") == 0) *value = m->launches;
    else if (strcmp(name, "n_train_padded") == 0) *value = m->n_pad;
    else if (strcmp(name, "workspace_bytes") == 0) *value = (long)m->workspace_bytes;
    else if (strcmp(name, "tri_launches") == 0) *value = (long)(m->tri_events_used / 2);
    else if (strcmp(name, "tri_ns") == 0) {
        // total device time of the tri_sumsq launches recorded since time_tri was switched on
        DeviceGuard guard(m->device);
        SEGP_CUDA_CHECK(cudaDeviceSynchronize());
        double total_ms = 0.0;
        for (size_t i = 0; i + 1 < m->tri_events_used; i += 2) {
            float ms = 0.f;
            SEGP_CUDA_CHECK(cudaEventElapsedTime(&ms, m->tri_events[i], m->tri_events[i + 1]));
            total_ms += ms;
        }
        *value = (long)(total_ms * 1e6);
    }
    else {
        set_error("segp_get_option: unknown option %s", name);
        return SEGP_ERR_INVALID;
    }
    return SEGP_OK;
}

int segp_set_param(segp_model* m, const char* name, double value) {
    if (m == nullptr || name == nullptr) {
        set_error("segp_set_param: null argument");
        return SEGP_ERR_INVALID;
    }
    if (strcmp(name, "guard_rtol") == 0 && value > 0.0 && value < 1.0) {
        m->guard_rtol = value;
        DeviceGuard guard(m->device);
        cudaDeviceSynchronize();
        free_graphs(m);
        return SEGP_OK;
    }
    if (strcmp(name, "guard_kappa") == 0 && value >= 1.0 && value <= 100.0) {
        m->guard_kappa = value;
        DeviceGuard guard(m->device);
        cudaDeviceSynchronize();
        free_graphs(m);
        return SEGP_OK;
    }
    set_error("segp_set_param: unknown parameter or bad value: %s=%g", name, value);
    return SEGP_ERR_INVALID;
}

int segp_get_param(segp_model* m, const char* name, double* value) {
    if (m == nullptr || name == nullptr || value == nullptr) {
        set_error("segp_get_param: null argument");
        return SEGP_ERR_INVALID;
    }
    static const struct {
        const char* name;
        int index;
    } stats[] = {{"probe_ran", PS_RAN},         {"probe_frac4", PS_FRAC4},   {"probe_frac5", PS_FRAC5},
                 {"probe_err4", PS_ERR4},       {"probe_err5", PS_ERR5},     {"probe_rel4", PS_REL4},
                 {"probe_rel5", PS_REL5},       {"probe_ratio4", PS_RATIO4}, {"probe_ratio5", PS_RATIO5},
                 {"probe_rho4", PS_RHO4},       {"probe_rho5", PS_RHO5},     {"probe_min_var_ratio", PS_MINVAR},
                 {"probe_margin4", PS_MARGIN4}};
    if (strcmp(name, "guard_rtol") == 0) {
        *value = m->guard_rtol;
        return SEGP_OK;
    }
    if (strcmp(name, "guard_kappa") == 0) {
        *value = m->guard_kappa;
        return SEGP_OK;
    }
    for (const auto& s : stats)
        if (strcmp(name, s.name) == 0) {
            *value = m->probe_stat[s.index];
            return SEGP_OK;
        }
    set_error("segp_get_param: unknown parameter %s", name);
    return SEGP_ERR_INVALID;
}

}  // extern "C"
