// C ABI of libsegp.so (include/segp.h): handle management, chunked rollout driver, host-buffer entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <new>
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

}  // namespace segp

using namespace segp;

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
    double* beta = nullptr;    // [n_s][n_pad]
    double* wt = nullptr;      // [n_s][ntri][128*128]
    double* logdet = nullptr;  // [n_s]
    // composite (linear x stationary + linear) kernels: segp_set_linear_terms
    bool has_composite = false, has_linear_terms = false;
    std::vector<double> h_plin, h_lin;
    double* xraw = nullptr;    // [n_pad][dim] unscaled inputs
    double* plin = nullptr;    // [n_s][dim]
    double* lin = nullptr;     // [n_s][dim]
    double* xtb = nullptr;     // [n_s][dim] X^T beta_d
    double* jac2_part = nullptr;   // workspace: additive Jacobian partials
    double* kss = nullptr;         // workspace: [n_s][b_cap] prior variances
    double* wdense = nullptr;  // [n_s][n_pad][n_pad] W = L^-1 kept dense for segp_append (only if opt_keep_w)
    long opt_keep_w = 0;       // keep wdense after factorising (set by the first segp_append)
    bool last_append_incremental = false;
    int8_t* wi8 = nullptr;     // [n_s][nblk (nblk+1)][I8_S][I8_A_TILE] digit planes of W (tcgen05 path)
    double* rowfac = nullptr;  // [n_s][n_pad] per-row factors of the digit planes
    // workspace
    long b_cap = 0;
    int nsplit = 1, blocks_per_split = 1;
    double* ks = nullptr;      // fp64 K* block (DMMA path)
    int8_t* ki8 = nullptr;     // digit planes of the K* block (tcgen05 path)
    long npanel_cap = 0;
    int ws_mode = 0;           // tri mode the current workspace is laid out for (0 fp64 ks, 1 ki8, 2 ki8 split halves)
    int8_t* i8zero = nullptr;  // I8_S * I8_A_TILE zero bytes (pair kernel)
    double* mu_part = nullptr;
    double* jac_part = nullptr;
    double* qpart = nullptr;
    StepParams* d_sp = nullptr;
    size_t workspace_bytes = 0;
    // host-entry staging (grown on demand)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // options
    long opt_chunk = 8192;
    long opt_panel_group = 16;
    long opt_ksplit = 0;   // 0 = automatic
    long opt_i8_panel_group = 0;   // tri_i8m / tri_i8mp: panels per L2 group (even), 0 = automatic
    long opt_i8_cluster = 2;       // tri_i8m: CTAs per cluster sharing one W stage by multicast (2 or 4)
    long opt_tri_mode = -1;   // -1 = automatic (4 when n_pad <= I8_MAX_NPAD, else 0), 0 = fp64 DMMA,
                              // 1 = int8 tcgen05 single CTA, 2 = CTA pair (cta_group::2), 3 = persistent CTA pair,
                              // 4 = single-CTA MMAs over two K* planes at once, W multicast over a CTA pair
    long opt_overlap = 0;     // 1 = software-pipeline two half-chunks over two internal streams (tri_mode 4 / 5): the
                              // FP64-bound K* kernel of one half runs as a resident grid of small CTAs NEXT TO the
                              // persistent contraction (tri_i8mp<4>) of the other.  Bit-identical.  The kernels do
                              // share the SMs, but the co-resident K* kernel runs ~7x slower than alone (most likely
                              // queued behind the contraction's shared-memory operand traffic) and becomes the
                              // critical path: 6 % slower than the serial schedule at C4, off by default
                              // (profiles/round1/overlap_pipeline_c3_c4_c5.txt)
    cudaStream_t s_hi = nullptr, s_lo = nullptr;   // internal streams of the pipelined driver (created on first use)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ks[2] = {nullptr, nullptr}, ev_tri[2] = {nullptr, nullptr};
    long opt_i8_ablate = 0;   // profiling only, see TriI8Args::ablate
    long long* i8_prof = nullptr;   // profiling only: [128][8] counters of the persistent kernel's MMA threads
    bool last_tri_persistent = false;   // which tcgen05 kernel the last contraction launch used (automatic mode)
    long launches = 0;
    // optional per-launch timing of tri_sumsq (bench.py roofline): event pairs recorded on the launching stream
    bool time_tri = false;
    std::vector<cudaEvent_t> tri_events;
    size_t tri_events_used = 0;
};

namespace segp {

static void free_model_buffers(segp_model* m) {
    dev_free(m->xs);
    dev_free(m->yp);
    dev_free(m->invls);
    dev_free(m->var);
    dev_free(m->beta);
    dev_free(m->wt);
    dev_free(m->logdet);
    dev_free(m->wi8);
    dev_free(m->rowfac);
    dev_free(m->i8zero);
    dev_free(m->wdense);
    dev_free(m->xraw);
    dev_free(m->plin);
    dev_free(m->lin);
    dev_free(m->xtb);
    m->has_linear_terms = false;
    m->factorized = false;
}

static void free_workspace(segp_model* m) {
    dev_free(m->ks);
    dev_free(m->ki8);
    m->npanel_cap = 0;
    dev_free(m->mu_part);
    dev_free(m->jac_part);
    dev_free(m->qpart);
    dev_free(m->jac2_part);
    dev_free(m->kss);
    m->b_cap = 0;
    m->workspace_bytes = 0;
}

// the int8 digit planes assume kernel values in [0, s_f^2]: composite kernels (unbounded linear terms) run in float64
static bool i8_capable(const segp_model* m) { return m->n_pad <= I8_MAX_NPAD && !m->has_composite; }
static int tri_mode(const segp_model* m) {
    if (m->has_composite) return 0;
    if (m->opt_tri_mode >= 0) return (int)m->opt_tri_mode;
    return i8_capable(m) ? 4 : 0;
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
    if (mode != m->ws_mode && m->b_cap > 0) free_workspace(m);
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
    if (m->has_composite) {
        SEGP_CHECK(dev_alloc(&m->jac2_part, n_jac));
        SEGP_CHECK(dev_alloc(&m->kss, (size_t)m->n_s * want));
    }
    if (n_ks > 0) SEGP_CUDA_CHECK(cudaMemset(m->ks, 0, n_ks * sizeof(double)));
    if (n_ki8 > 0) SEGP_CUDA_CHECK(cudaMemset(m->ki8, 0, n_ki8));
    m->workspace_bytes = (n_ks + n_mu + n_jac + n_q) * sizeof(double) + n_ki8;
    m->b_cap = want;
    m->npanel_cap = npanel_cap;
    m->ws_mode = mode;
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
static int run_kstar(segp_model* m, const KstarArgs& k, cudaStream_t st, int panel0 = 0, bool coresident = false) {
    if (m->ws_mode != 0) {
        KstarI8Args k8{};
        k8.k = k;
        k8.ki8 = m->ki8;
        k8.npanel_cap = m->npanel_cap;
        k8.split_halves = m->ws_mode == 2 || m->ws_mode == 3;
        k8.panel0 = panel0;
        k8.resident_ctas = coresident ? 3 * 148 : 0;   // three small CTAs per SM next to the persistent contraction
        return launch_kstar_i8(k8, m->n_s, m->nsplit, st);
    }
    return launch_kstar(k, m->n_s, m->nsplit, st);
}

// variance contraction launch (tri_i8 on tcgen05 or tri_sumsq on the DMMA pipe), optionally bracketed by a
// CUDA-event pair on the launching stream
static int run_tri(segp_model* m, long nb, cudaStream_t st, int panel0 = 0, bool coresident = false) {
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
        TriI8Args t{};
        t.wi8 = m->wi8;
        t.rowfac = m->rowfac;
        t.ki8 = m->ki8;
        t.qpart = m->qpart;
        t.nblk = m->nblk;
        t.npanels = (int)((nb + I8_N - 1) / I8_N);
        t.panel0 = panel0;   // != 0 only from the pipelined driver (ws_mode 4)
        // panels per L2 group of tri_i8m / tri_i8mp (0 = 24): a sweep over 8..24 at C4 and C5 moved the launch time by
        // less than 0.5 % (scripts/gpu_pgroup.sh), so there is no automatic choice
        t.pgroup = (int)m->opt_i8_panel_group;
        t.cluster = (int)m->opt_i8_cluster;
        t.npanel_cap = m->npanel_cap;
        t.b_cap = m->b_cap;
        t.dbg = nullptr;
        t.fix_bi = -1;
        t.zero_a = m->i8zero;
        t.ablate = (int)m->opt_i8_ablate;
        t.prof = m->i8_prof;
        // Automatic mode: the persistent folded kernel where per-tile overhead matters (short tiles, enough of them to
        // balance a static schedule: +5 % at C3), the one-cluster-per-tile kernel otherwise (C4: power-capped, +1 %
        // at best; C5: the persistent order runs 13 % slower; C2: too few tiles) -- profiles/round1/persistent_tri_i8mp.txt
        bool persistent = m->ws_mode == 5;
        if (m->ws_mode == 4 && m->opt_tri_mode < 0 && t.panel0 == 0) {
            const long ntiles = (long)m->n_s * ((m->nblk + 1) / 2) * ((t.npanels + 1) / 2);
            persistent = m->nblk <= 32 && ntiles >= 4 * 74;
        }
        if (coresident) persistent = true;   // the pipelined driver needs a resident contraction grid
        m->last_tri_persistent = persistent;
        SEGP_CHECK(persistent        ? launch_tri_i8mp(t, m->n_s, st, coresident)
                   : m->ws_mode == 4 ? launch_tri_i8m(t, m->n_s, st)
                   : m->ws_mode == 3 ? launch_tri_i8x2p(t, m->n_s, st)
                   : m->ws_mode == 2 ? launch_tri_i8x2(t, m->n_s, st)
                                     : launch_tri_i8(t, m->n_s, st));
    } else {
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
    }
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
    if (dev_alloc(&m->d_sp, 1) != SEGP_OK) {
        delete m;
        return SEGP_ERR_CUDA;
    }
    *out = m;
    return SEGP_OK;
}

int segp_destroy(segp_model* m) {
    if (m == nullptr) return SEGP_OK;
    DeviceGuard guard(m->device);
    cudaDeviceSynchronize();
    free_model_buffers(m);
    free_workspace(m);
    dev_free(m->d_sp);
    dev_free(m->i8_prof);
    for (cudaEvent_t e : m->tri_events) cudaEventDestroy(e);
    for (cudaEvent_t e : {m->ev_fork, m->ev_join, m->ev_ks[0], m->ev_ks[1], m->ev_tri[0], m->ev_tri[1]})
        if (e != nullptr) cudaEventDestroy(e);
    if (m->s_hi != nullptr) cudaStreamDestroy(m->s_hi);
    if (m->s_lo != nullptr) cudaStreamDestroy(m->s_lo);
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

int segp_alloc_factor_buffers(segp_model* m) {
    if (m == nullptr || !m->has_data) {
        set_error("segp_alloc_factor_buffers: segp_set_model has not been called");
        return SEGP_ERR_NOT_TRAINED;
    }
    DeviceGuard guard(m->device);
    if (m->wt == nullptr) SEGP_CHECK(dev_alloc(&m->wt, (size_t)m->n_s * m->ntri * TILE * TILE));
    if (m->beta == nullptr) SEGP_CHECK(dev_alloc(&m->beta, (size_t)m->n_s * m->n_pad));
    if (m->logdet == nullptr) SEGP_CHECK(dev_alloc(&m->logdet, (size_t)m->n_s));
    if (i8_capable(m)) {
        if (m->wi8 == nullptr)
            SEGP_CHECK(dev_alloc(&m->wi8, (size_t)m->n_s * m->nblk * (m->nblk + 1) * (I8_S * I8_A_TILE)));
        if (m->rowfac == nullptr) SEGP_CHECK(dev_alloc(&m->rowfac, (size_t)m->n_s * m->n_pad));
        if (m->i8zero == nullptr) {
            SEGP_CHECK(dev_alloc(&m->i8zero, (size_t)I8_S * I8_A_TILE));
            SEGP_CUDA_CHECK(cudaMemset(m->i8zero, 0, (size_t)I8_S * I8_A_TILE));
        }
    }
    return SEGP_OK;
}

int segp_num_factor_buffers(segp_model* m) {
    if (m == nullptr) return 0;
    return i8_capable(m) ? 5 : 3;
}

int segp_factor_buffer(segp_model* m, int index, void** d_ptr, size_t* bytes) {
    if (m == nullptr || d_ptr == nullptr || bytes == nullptr || m->wt == nullptr) {
        set_error("segp_factor_buffer: buffers are not allocated");
        return SEGP_ERR_NOT_TRAINED;
    }
    switch (index) {
        case 0:
            *d_ptr = m->wt;
            *bytes = (size_t)m->n_s * m->ntri * TILE * TILE * sizeof(double);
            return SEGP_OK;
        case 1:
            *d_ptr = m->beta;
            *bytes = (size_t)m->n_s * m->n_pad * sizeof(double);
            return SEGP_OK;
        case 2:
            *d_ptr = m->logdet;
            *bytes = (size_t)m->n_s * sizeof(double);
            return SEGP_OK;
        case 3:
            if (m->wi8 != nullptr) {
                *d_ptr = m->wi8;
                *bytes = (size_t)m->n_s * m->nblk * (m->nblk + 1) * (I8_S * I8_A_TILE);
                return SEGP_OK;
            }
            [[fallthrough]];
        case 4:
            if (index == 4 && m->rowfac != nullptr) {
                *d_ptr = m->rowfac;
                *bytes = (size_t)m->n_s * m->n_pad * sizeof(double);
                return SEGP_OK;
            }
            [[fallthrough]];
        default:
            set_error("segp_factor_buffer: index %d out of range", index);
            return SEGP_ERR_INVALID;
    }
}

int segp_mark_factorized(segp_model* m) {
    if (m == nullptr || m->wt == nullptr) {
        set_error("segp_mark_factorized: buffers are not allocated");
        return SEGP_ERR_NOT_TRAINED;
    }
    if (m->has_composite) {
        if (!m->has_linear_terms) {
            set_error("segp_mark_factorized: composite kernel without linear terms (call segp_set_linear_terms)");
            return SEGP_ERR_INVALID;
        }
        DeviceGuard guard(m->device);
        SEGP_CHECK(compute_xtb(m, nullptr));
        SEGP_CUDA_CHECK(cudaStreamSynchronize(nullptr));
    }
    m->factorized = true;
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
    SEGP_CHECK(segp_alloc_factor_buffers(m));
    const size_t nn = (size_t)m->n_pad * m->n_pad;
    const int nb64 = m->n_pad / NBLK;
    // The output dimensions are independent factorisations and each is a chain of ~270 launches, many of them one
    // block wide (the 64 x 64 diagonal blocks): up to FACT_SLOTS of them run concurrently on their own streams, each
    // with its own scratch, so the latency-bound launches of one overlap the GEMMs of the others.
    constexpr int FACT_SLOTS = 4;
    struct Slot {
        double *kbuf = nullptr, *wbuf = nullptr, *tmp = nullptr, *diag_inv = nullptr, *u_tmp = nullptr;
        cudaStream_t s = nullptr;
        cudaEvent_t done = nullptr;
    };
    const int nslots = std::min(m->n_s, FACT_SLOTS);
    Slot slots[FACT_SLOTS];
    cudaEvent_t fork = nullptr;
    int* d_fail = nullptr;
    int rc = SEGP_OK;
    std::vector<int> fails(m->n_s, 0);
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
            if ((rc = dev_alloc(&sl.kbuf, nn)) != SEGP_OK) break;
            if (!m->opt_keep_w && (rc = dev_alloc(&sl.wbuf, nn)) != SEGP_OK) break;
            if ((rc = dev_alloc(&sl.tmp, nn)) != SEGP_OK) break;
            if ((rc = dev_alloc(&sl.diag_inv, (size_t)nb64 * NBLK * NBLK)) != SEGP_OK) break;
            if ((rc = dev_alloc(&sl.u_tmp, (size_t)33 * m->n_pad)) != SEGP_OK) break;
            if (cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess ||
                cudaStreamWaitEvent(sl.s, fork, 0) != cudaSuccess) {
                set_error("segp_factorize: stream set-up failed");
                rc = SEGP_ERR_CUDA;
            }
        }
        if (rc != SEGP_OK) break;
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
            if ((rc = potrf_lower(sl.kbuf, m->n_pad, sl.diag_inv, d_fail + d, ss, &m->launches)) != SEGP_OK) break;
            if ((rc = logdet_from_chol(sl.kbuf, m->n_train, m->n_pad, m->logdet + d, ss)) != SEGP_OK) break;
            ++m->launches;
            if (cudaMemsetAsync(wbuf, 0, nn * sizeof(double), ss) != cudaSuccess) {
                set_error("cudaMemsetAsync failed");
                rc = SEGP_ERR_CUDA;
                break;
            }
            if ((rc = trtri_lower(sl.kbuf, wbuf, m->n_pad, sl.diag_inv, sl.tmp, ss, &m->launches)) != SEGP_OK) break;
            if ((rc = solve_beta(wbuf, m->yp + (size_t)d * m->n_pad, sl.u_tmp, m->beta + (size_t)d * m->n_pad, m->n_pad,
                                 ss)) != SEGP_OK)
                break;
            m->launches += 3;
            if ((rc = pack_w(wbuf, m->wt + (size_t)d * m->ntri * TILE * TILE, m->n_pad, ss)) != SEGP_OK) break;
            ++m->launches;
            if (m->wi8 != nullptr) {
                if ((rc = pack_w_i8(wbuf, m->wi8 + (size_t)d * m->nblk * (m->nblk + 1) * (I8_S * I8_A_TILE),
                                    m->rowfac + (size_t)d * m->n_pad, m->h_var[d], m->n_pad, ss)) != SEGP_OK)
                    break;
                m->launches += 2;
            }
        }
        if (rc != SEGP_OK) break;
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
    for (int i = 0; i < FACT_SLOTS; ++i) {
        Slot& sl = slots[i];
        dev_free(sl.kbuf);
        dev_free(sl.wbuf);
        dev_free(sl.tmp);
        dev_free(sl.diag_inv);
        dev_free(sl.u_tmp);
        if (sl.done != nullptr) cudaEventDestroy(sl.done);
        if (sl.s != nullptr) cudaStreamDestroy(sl.s);
    }
    if (fork != nullptr) cudaEventDestroy(fork);
    dev_free(d_fail);
    if (rc == SEGP_OK) m->factorized = true;
    return rc;
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
    if (m->xraw != nullptr)
        SEGP_CUDA_CHECK(cudaMemcpy(m->xraw + (size_t)n_old * dim, h_x, (size_t)n_new * dim * sizeof(double),
                                   cudaMemcpyHostToDevice));
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
            if ((rc = pack_w(w, m->wt + (size_t)d * m->ntri * TILE * TILE, n_pad, st)) != SEGP_OK) break;
            m->launches += 5;
            if (m->wi8 != nullptr) {
                if ((rc = pack_w_i8(w, m->wi8 + (size_t)d * m->nblk * (m->nblk + 1) * (I8_S * I8_A_TILE),
                                    m->rowfac + (size_t)d * n_pad, m->h_var[d], n_pad, st)) != SEGP_OK)
                    break;
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
    if (rc == SEGP_OK) {
        m->factorized = true;
        m->last_append_incremental = true;
    }
    return rc;
}

int segp_logdet(segp_model* m, double* h_out) {
    SEGP_CHECK(check_ready(m));
    DeviceGuard guard(m->device);
    SEGP_CUDA_CHECK(cudaMemcpy(h_out, m->logdet, m->n_s * sizeof(double), cudaMemcpyDeviceToHost));
    return SEGP_OK;
}

int segp_predict(segp_model* m, long n_batch, const double* d_z, double* d_mu, double* d_var, double* d_jac,
                 void* stream) {
    SEGP_CHECK(check_ready(m));
    if (n_batch < 0 || (n_batch > 0 && (d_z == nullptr || d_mu == nullptr || d_var == nullptr))) {
        set_error("segp_predict: null buffer");
        return SEGP_ERR_INVALID;
    }
    if (n_batch == 0) return SEGP_OK;
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
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
    SEGP_CHECK(ensure_workspace(m, n_batch));
    SEGP_CUDA_CHECK(cudaMemcpyAsync(m->d_sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, st));
    if (d_status != nullptr) SEGP_CUDA_CHECK(cudaMemsetAsync(d_status, 0, n_batch * sizeof(int32_t), st));

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

    bool forked = false;
    for (long c0 = 0; c0 < n_batch; c0 += m->b_cap) {
        const long nb = std::min<long>(m->b_cap, n_batch - c0);
        const int npanels = (int)((nb + I8_N - 1) / I8_N);
        // Two half-chunks, software-pipelined over two internal streams: all contractions on the high-priority
        // stream, back to back (tensor pipe), the K* / ellipsoid kernels of the OTHER half under them on the
        // low-priority stream (FP64 pipe).  Within a half the order kstar -> tri -> ellipsoid -> kstar(t+1) is kept by
        // events; the halves touch disjoint panel ranges of the workspace.  Same kernels, same arithmetic, same
        // fixed-order reductions: results are bit-identical to the serial schedule.
        const bool pipelined = (m->ws_mode == 4 || m->ws_mode == 5) && m->opt_overlap != 0 && npanels >= 48 && m->n_pad >= 1024;
        if (!pipelined) {
            cudaStream_t s1 = forked ? m->s_lo : st;
            for (int t = 0; t < horizon; ++t) {
                SEGP_CHECK(run_kstar(m, kstar_args(c0, t, nb), s1));
                SEGP_CHECK(run_tri(m, nb, s1));
                SEGP_CHECK(launch_ellipsoid_step(step_args(c0, t, 0, nb), s1));
                m->launches += 3;
            }
            continue;
        }
        if (m->s_hi == nullptr) {
            int least = 0, greatest = 0;
            SEGP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            SEGP_CUDA_CHECK(cudaStreamCreateWithPriority(&m->s_hi, cudaStreamNonBlocking, greatest));
            SEGP_CUDA_CHECK(cudaStreamCreateWithPriority(&m->s_lo, cudaStreamNonBlocking, least));
            for (cudaEvent_t* e : {&m->ev_fork, &m->ev_join, &m->ev_ks[0], &m->ev_ks[1], &m->ev_tri[0], &m->ev_tri[1]})
                SEGP_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        if (!forked) {
            SEGP_CUDA_CHECK(cudaEventRecord(m->ev_fork, st));
            SEGP_CUDA_CHECK(cudaStreamWaitEvent(m->s_hi, m->ev_fork, 0));
            SEGP_CUDA_CHECK(cudaStreamWaitEvent(m->s_lo, m->ev_fork, 0));
            forked = true;
        }
        // debugging aid (SEGP_TIMELINE=1): device timestamps of every launch of the first pipelined steps, on stderr
        static const bool timeline = getenv("SEGP_TIMELINE") != nullptr;
        struct Mark {
            const char* what;
            int t, h;
            cudaEvent_t e0, e1;
        };
        std::vector<Mark> marks;
        cudaEvent_t tl_base = nullptr;
        auto mark_begin = [&](const char* what, int t, int h, cudaStream_t s) {
            if (!timeline || t > 2) return;
            Mark mk{what, t, h, nullptr, nullptr};
            cudaEventCreate(&mk.e0);
            cudaEventCreate(&mk.e1);
            cudaEventRecord(mk.e0, s);
            marks.push_back(mk);
        };
        auto mark_end = [&](int t, cudaStream_t s) {
            if (!timeline || t > 2) return;
            cudaEventRecord(marks.back().e1, s);
        };
        if (timeline) {
            cudaEventCreate(&tl_base);
            cudaEventRecord(tl_base, m->s_lo);
        }
        const int pa = ((npanels / 2 + 1) / 2) * 2;   // even, so the cluster pairs of the first half are complete
        const int p_begin[2] = {0, pa}, p_end[2] = {pa, npanels};
        const long b_begin[2] = {0, (long)pa * I8_N}, b_end[2] = {std::min<long>((long)pa * I8_N, nb), nb};
        for (int t = 0; t <= horizon; ++t) {
            for (int h = 0; h < 2; ++h) {
                if (t > 0) {
                    SEGP_CUDA_CHECK(cudaStreamWaitEvent(m->s_lo, m->ev_tri[h], 0));
                    mark_begin("ell", t - 1, h, m->s_lo);
                    SEGP_CHECK(launch_ellipsoid_step(step_args(c0, t - 1, b_begin[h], b_end[h]), m->s_lo));
                    mark_end(t - 1, m->s_lo);
                    ++m->launches;
                }
                if (t == horizon) continue;
                mark_begin("kstar", t, h, m->s_lo);
                SEGP_CHECK(run_kstar(m, kstar_args(c0, t, b_end[h]), m->s_lo, p_begin[h], true));
                mark_end(t, m->s_lo);
                SEGP_CUDA_CHECK(cudaEventRecord(m->ev_ks[h], m->s_lo));
                SEGP_CUDA_CHECK(cudaStreamWaitEvent(m->s_hi, m->ev_ks[h], 0));
                mark_begin("tri", t, h, m->s_hi);
                SEGP_CHECK(run_tri(m, (long)p_end[h] * I8_N, m->s_hi, p_begin[h], true));
                mark_end(t, m->s_hi);
                SEGP_CUDA_CHECK(cudaEventRecord(m->ev_tri[h], m->s_hi));
                m->launches += 2;
            }
        }
        if (timeline) {
            cudaDeviceSynchronize();
            for (const Mark& mk : marks) {
                float t0 = 0.f, t1 = 0.f;
                cudaEventElapsedTime(&t0, tl_base, mk.e0);
                cudaEventElapsedTime(&t1, tl_base, mk.e1);
                fprintf(stderr, "[segp timeline] %-5s step %d half %d: %8.3f -> %8.3f ms (%.3f)\n", mk.what, mk.t, mk.h, t0,
                        t1, t1 - t0);
                cudaEventDestroy(mk.e0);
                cudaEventDestroy(mk.e1);
            }
            cudaEventDestroy(tl_base);
        }
    }
    if (forked) {   // everything issued on s_hi has been waited for by s_lo
        SEGP_CUDA_CHECK(cudaEventRecord(m->ev_join, m->s_lo));
        SEGP_CUDA_CHECK(cudaStreamWaitEvent(st, m->ev_join, 0));
    }
    return SEGP_OK;
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
    cudaStream_t st = nullptr;
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
    SEGP_CHECK(segp_multistep(m, n_batch, horizon, base + o_p0, p0_stride, h_q0 ? base + o_q0 : nullptr, q0_stride,
                              base + o_kff, h_k_fb ? base + o_kfb : nullptr, kfb_stride,
                              h_k_fb_init ? base + o_kfbi : nullptr, kfb_init_stride, params, base + o_pall,
                              base + o_qall, h_var_all ? base + o_var : nullptr, h_status ? d_stat : nullptr, st));
    SEGP_CUDA_CHECK(cudaMemcpyAsync(h_p_all, base + o_pall, n_pall * sizeof(double), cudaMemcpyDeviceToHost, st));
    SEGP_CUDA_CHECK(cudaMemcpyAsync(h_q_all, base + o_qall, n_qall * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h_var_all != nullptr)
        SEGP_CUDA_CHECK(cudaMemcpyAsync(h_var_all, base + o_var, n_pall * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h_status != nullptr)
        SEGP_CUDA_CHECK(cudaMemcpyAsync(h_status, d_stat, n_batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
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

int segp_i8_selftest(int device, int variant, int k_blocks, const int8_t* h_a, const int8_t* h_b, int32_t* h_acc,
                     double* h_colsum) {
    if ((variant < 1 || variant > 3) || k_blocks < 1 || k_blocks > 64 || (variant >= 2 && (k_blocks & 1)) ||
        h_a == nullptr || h_b == nullptr || h_acc == nullptr || h_colsum == nullptr) {
        set_error("segp_i8_selftest: bad argument (variant 1|2, k_blocks in [1,64], even for variant 2)");
        return SEGP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("segp_i8_selftest: cudaSetDevice(%d) failed", device);
        return SEGP_ERR_CUDA;
    }
    SEGP_CHECK(tri_i8_init());
    // one tile of a (128 k_blocks)-point model, panel 0, K = 128 k_blocks:
    //   variant 1: block row k_blocks - 1 (128 rows);  variant 2: block rows k_blocks - 2 and k_blocks - 1 (256 rows),
    //   where the upper block row sees zeros in the last 128 columns (its own k-range ends one block earlier)
    const int nblk = k_blocks, kdim = TILE * k_blocks, nkb = 2 * k_blocks;
    const int rows = variant >= 2 ? 2 * TILE : TILE;
    const int bi0 = variant >= 2 ? k_blocks - 2 : k_blocks - 1;
    const size_t a_bytes = (size_t)nblk * (nblk + 1) * (I8_S * I8_A_TILE);
    const size_t b_bytes = (size_t)nkb * (I8_S * I8_B_TILE);
    std::vector<int8_t> a_img(a_bytes, 0), b_img(b_bytes, 0);
    auto sw = [](int r, int k) { return r * I8_KB + ((((k >> 4) ^ ((r >> 1) & 3)) << 4) | (k & 15)); };
    // same image formats as pack_w_i8_kernel / kstar_i8_kernel
    for (int rb = 0; rb < rows / TILE; ++rb) {
        const int bi = bi0 + rb;
        for (int kb = 0; kb < 2 * (bi + 1); ++kb)
            for (int pl = 0; pl < I8_S; ++pl) {
                int8_t* at = a_img.data() + ((size_t)bi * (bi + 1) + kb) * (I8_S * I8_A_TILE) + (size_t)pl * I8_A_TILE;
                for (int k = 0; k < I8_KB; ++k)
                    for (int r = 0; r < TILE; ++r)
                        at[sw(r, k)] = h_a[((size_t)pl * rows + rb * TILE + r) * kdim + (size_t)kb * I8_KB + k];
            }
    }
    for (int kb = 0; kb < nkb; ++kb)
        for (int pl = 0; pl < I8_S; ++pl)
            for (int k = 0; k < I8_KB; ++k)
                for (int r = 0; r < I8_N; ++r) {
                    const int8_t v = h_b[((size_t)pl * I8_N + r) * kdim + (size_t)kb * I8_KB + k];
                    int8_t* kbase = b_img.data() + (size_t)kb * (I8_S * I8_B_TILE);
                    if (variant >= 2)
                        kbase[(size_t)(r / (I8_N / 2)) * (I8_S * (I8_B_TILE / 2)) + (size_t)pl * (I8_B_TILE / 2) +
                              sw(r % (I8_N / 2), k)] = v;
                    else
                        kbase[(size_t)pl * I8_B_TILE + sw(r, k)] = v;
                }
    int8_t *d_a = nullptr, *d_b = nullptr, *d_zero = nullptr;
    double *d_rf = nullptr, *d_q = nullptr;
    int32_t* d_dbg = nullptr;
    const size_t n_dbg = (size_t)I8_S * rows * I8_N;
    int rc = SEGP_OK;
    do {
        if ((rc = dev_alloc(&d_a, a_bytes)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_b, b_bytes)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_zero, (size_t)I8_S * I8_A_TILE)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_rf, (size_t)kdim)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_q, (size_t)nblk * I8_N)) != SEGP_OK) break;
        if ((rc = dev_alloc(&d_dbg, n_dbg)) != SEGP_OK) break;
        std::vector<double> ones((size_t)kdim, 1.0);
        cudaError_t e = cudaMemcpy(d_a, a_img.data(), a_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_b, b_img.data(), b_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(d_zero, 0, (size_t)I8_S * I8_A_TILE);
        if (e == cudaSuccess) e = cudaMemcpy(d_rf, ones.data(), kdim * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(d_q, 0, (size_t)nblk * I8_N * sizeof(double));
        if (e != cudaSuccess) {
            set_error("segp_i8_selftest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
            break;
        }
        TriI8Args t{};
        t.wi8 = d_a;
        t.rowfac = d_rf;
        t.ki8 = d_b;
        t.qpart = d_q;
        t.nblk = nblk;
        t.npanels = 1;
        t.npanel_cap = 1;
        t.b_cap = I8_N;
        t.dbg = d_dbg;
        t.zero_a = d_zero;
        t.fix_bi = variant >= 2 ? bi0 / 2 : bi0;
        if ((rc = (variant == 3   ? launch_tri_i8x2p(t, 1, nullptr)
                   : variant == 2 ? launch_tri_i8x2(t, 1, nullptr)
                                  : launch_tri_i8(t, 1, nullptr))) != SEGP_OK)
            break;
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(h_acc, d_dbg, n_dbg * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)   // column sums of the block row(s): [rows / 128][96]
            e = cudaMemcpy(h_colsum, d_q + (size_t)bi0 * I8_N, (size_t)(rows / TILE) * I8_N * sizeof(double),
                           cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("segp_i8_selftest: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
        }
    } while (0);
    dev_free(d_a);
    dev_free(d_b);
    dev_free(d_zero);
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
    if (strcmp(name, "tri_mode") == 0 && value >= -1 && value <= 5) {
        if (value >= 1 && m->has_composite) {
            set_error("tri_mode=%ld (int8 tcgen05) is not available with composite (lin_*) kernels: float64 only", value);
            return SEGP_ERR_UNSUPPORTED;
        }
        if (value >= 1 && m->has_data && !i8_capable(m)) {
            set_error("tri_mode=%ld (int8 tcgen05) needs n_train_padded <= %ld; this model has %d", value, I8_MAX_NPAD,
                      m->n_pad);
            return SEGP_ERR_UNSUPPORTED;
        }
        m->opt_tri_mode = value;
        return SEGP_OK;
    }
    if (strcmp(name, "overlap") == 0 && (value == 0 || value == 1)) {
        m->opt_overlap = value;
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
    if (strcmp(name, "i8_ablate") == 0) {
        m->opt_i8_ablate = value;
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
    else if (strcmp(name, "overlap") == 0) *value = m->opt_overlap;
    else if (strcmp(name, "i8_prof_ptr") == 0) *value = (long)(uintptr_t)m->i8_prof;
    else if (strcmp(name, "tri_mode_effective") == 0) *value = tri_mode(m);
    else if (strcmp(name, "tri_persistent") == 0) *value = m->last_tri_persistent ? 1 : 0;
    else if (strcmp(name, "append_incremental") == 0) *value = m->last_append_incremental ? 1 : 0;
    else if (strcmp(name, "keep_w") == 0) *value = m->opt_keep_w;
    else if (strcmp(name, "n_train") == 0) *value = m->n_train;
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

}  // extern "C"
