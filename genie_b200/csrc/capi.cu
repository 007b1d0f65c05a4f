// extern "C" surface of libgenie_b200.so (declared in include/genie_b200.h).
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "internal.h"
#include <cstring>
#include "input.cuh"

using namespace gl;

namespace {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<uint64_t> g_plan_counter{0};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel device timing ---------------------------------------------------------------------------------
namespace {
struct TimingSlot {
    int kid;
    cudaEvent_t beg, end;
};
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mu;
std::vector<TimingSlot*> g_timing_pending, g_timing_free;
double g_timing_ms[KID_COUNT];
int64_t g_timing_n[KID_COUNT];
const char* const kKernelNames[KID_COUNT] = {"pack_weights_kernel", "input_series_kernel", "input_gather_kernel",
                                             "da_init_kernel",      "da_layer1_kernel",    "da_layer1_tc_kernel", "da_layer2_readin_kernel",
                                             "readin_finalize_kernel", "sa_pre_kernel",    "sa_main_kernel",
                                             "src_mean32_kernel",   "src_mean16_kernel",   "da_layer1_s_kernel",
                                             "da_layer2_s_kernel",  "heads_grid_kernel",   "heads_query_kernel",
                                             "assoc_grid_pre_kernel", "assoc_init_kernel", "assoc_layer1_kernel",
                                             "assoc_layer2_kernel", "assoc_collapse_kernel", "knn_kernel", "stack_output_kernel", "kron_spmm_kernel",
                                             "node_mlp_fwd_kernel", "node_mlp_bwd_kernel"};
}  // namespace

TimedLaunch::TimedLaunch(int kid_, cudaStream_t st_) : kid(kid_), st(st_), slot(nullptr) {
    if (!g_timing_on.load(std::memory_order_relaxed)) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    TimingSlot* s = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        if (!g_timing_free.empty()) {
            s = g_timing_free.back();
            g_timing_free.pop_back();
        }
    }
    if (!s) {
        s = new TimingSlot();
        if (cudaEventCreate(&s->beg) != cudaSuccess || cudaEventCreate(&s->end) != cudaSuccess) {
            delete s;
            return;
        }
    }
    s->kid = kid;
    cudaEventRecord(s->beg, st);
    slot = s;
}

TimedLaunch::~TimedLaunch() {
    if (!slot) return;
    TimingSlot* s = static_cast<TimingSlot*>(slot);
    cudaEventRecord(s->end, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing_pending.push_back(s);
}

Workspace carve_workspace(const genie_plan* p, void* base) {
    Workspace w;
    const size_t P = (size_t)p->g.n_prod, G = (size_t)p->g.n_grid;
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        float* ptr = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
        off += align_up(n_floats * sizeof(float), 256);
        return ptr;
    };
    // bf16 storage (genie_plan_set_storage): tr0 / msrc / vb (and mean_src(vb), which re-uses tr0) hold bf16 rows
    const size_t half = p->storage == GENIE_STORAGE_BF16 ? 2 : 1;
    w.tr0 = take(P * LD_TR0 / half);
    w.msrc = split_supported(p) ? take(P * LD_TR0 / half) : nullptr;
    w.zc = take(P * LD_ZC);
    w.va = take(P * LD_V);
    w.vb = take(P * LD_V / half);
    w.xg = take(G * 32);
    w.r = take(G * 16);
    w.px = take(G * 32);
    w.sa_a = take(G * 32);
    w.sa_b = take(G * 32);
    w.partial = take(1024 * 8);
    w.bytes = off;
    return w;
}

// Layer 0 + layer 1 of DataAggregation.  Which kernels do the work is decided ON THE DEVICE from the packed weights
// (layout.h TCS_OK: the tensor-core kernels need PReLU12 to be invertible): both families are launched, one exits at once.
//   plans with tiling tables: source pass (src_mean) + station pass (da_layer1_s) on tcgen05
//   other CARTESIAN plans:    one-pass tcgen05 kernel (da_layer1_tc)
//   fallback / EXPLICIT:      generic FFMA kernel (da_layer1)
static int launch_da_layers01(const genie_plan* plan, const float* packed, const float* slice, const float* mask,
                              const Workspace& w, cudaStream_t st) {
    const bool split = split_supported(plan);
    if (plan->storage == GENIE_STORAGE_BF16) {
        // bf16 rows exist only on the split tensor-core path (the caller checked the PReLU slopes, layout.h TCS_OK)
        int rc;
        if ((rc = launch_da_init(plan, packed, slice, mask, w.tr0, true, st))) return rc;
        if ((rc = launch_src_mean(plan, 32, w.tr0, w.msrc, nullptr, st, plan->storage))) return rc;
        return launch_da_layer1_s(plan, packed, w.tr0, w.msrc, mask, w.zc, w.va, w.vb, st);
    }
    // the edge-feature terms are implemented in the split and the generic kernels, not in the one-pass tensor-core kernel
    const bool tc = split || (da_tc_supported(plan) && plan->edge_sta == nullptr);
    int rc;
    if ((rc = launch_da_init(plan, packed, slice, mask, w.tr0, tc, st))) return rc;
    if (split) {
        const float* gate = packed + TC_BASE + TC_SCAL + TCS_OK;
        if ((rc = launch_src_mean(plan, 32, w.tr0, w.msrc, gate, st, plan->storage))) return rc;
        if ((rc = launch_da_layer1_s(plan, packed, w.tr0, w.msrc, mask, w.zc, w.va, w.vb, st))) return rc;
    } else if (tc) {
        if ((rc = launch_da_layer1_tc(plan, packed, w.tr0, mask, w.zc, w.va, w.vb, st))) return rc;
    }
    return launch_da_layer1(plan, packed, w.tr0, mask, w.zc, w.va, w.vb, tc, st);
}

// The same with a1 folded into layer 0 (genie_window_fwd): split plans only; the mask travels inside the feature rows.
static int launch_da_layers01_window(const genie_plan* plan, const float* packed, const WindowParamSrc& ws,
                                     const int32_t* ind_use, const float* trv, const float* series, float* slice_out,
                                     float* mask_out, const Workspace& w, cudaStream_t st) {
    int rc;
    if ((rc = launch_da_init_fused(plan, packed, ws, ind_use, trv, series, slice_out, mask_out, w.tr0, true, st))) return rc;
    const float* gate = packed + TC_BASE + TC_SCAL + TCS_OK;
    if ((rc = launch_src_mean(plan, 32, w.tr0, w.msrc, gate, st, plan->storage))) return rc;
    if ((rc = launch_da_layer1_s(plan, packed, w.tr0, w.msrc, nullptr, w.zc, w.va, w.vb, st))) return rc;
    if (plan->storage == GENIE_STORAGE_BF16) return GENIE_OK;       // no generic fallback on bf16 rows
    return launch_da_layer1(plan, packed, w.tr0, nullptr, w.zc, w.va, w.vb, true, st);
}

// Layer 2 of DataAggregation (+ Bipartite_ReadIn when `readin_out` is given) from zc / va / vb.
static int launch_da_layer2(const genie_plan* plan, const float* packed, const Workspace& w, const float* mask,
                            const float* edge_attr, float* latent_out, float* readin_out, int ld_r, cudaStream_t st) {
    int rc;
    if (split_supported(plan) && plan->g.n_sta_tiles <= 32 && edge_attr != nullptr) {
        float* m2 = w.tr0;                                 // layer-0 features are dead: re-use their buffer
        if ((rc = launch_src_mean(plan, 16, w.vb, m2, nullptr, st, plan->storage, plan->halo_vb))) return rc;
        return launch_da_layer2_s(plan, packed, w.zc, w.va, m2, mask, edge_attr, latent_out, readin_out ? readin_out : w.r,
                                  readin_out ? ld_r : 16, st);
    }
    if (plan->storage == GENIE_STORAGE_BF16) {
        set_error("bf16 storage: layer 2 runs only fused with the read-in on a plan with tiling tables (genie_frontend_fwd / "
                  "genie_window_fwd)");
        return GENIE_ERR_UNSUPPORTED;
    }
    if (readin_out) {
        GENIE_CUDA_CHECK(cudaMemsetAsync(w.xg, 0, (size_t)plan->g.n_grid * 32 * sizeof(float), st));
        const int mode = L2_GATHER | L2_READIN | (latent_out ? L2_STORE_LATENT : 0);
        if ((rc = launch_da_layer2_readin(plan, packed, mode, w.zc, w.va, w.vb, nullptr, latent_out, edge_attr, mask, w.xg,
                                          st)))
            return rc;
        return launch_readin_finalize(plan, packed, w.xg, readin_out, ld_r, st);
    }
    return launch_da_layer2_readin(plan, packed, L2_GATHER | L2_STORE_LATENT, w.zc, w.va, w.vb, nullptr, latent_out,
                                   nullptr, mask, nullptr, st);
}

extern "C" {

const char* genie_last_error(void) { return g_last_error.c_str(); }
int genie_abi_version(void) { return GENIE_B200_ABI_VERSION; }
int64_t genie_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int genie_timing_enable(int on) {
    g_timing_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return GENIE_OK;
}
int genie_timing_kernel_count(void) { return (int)KID_COUNT; }
const char* genie_timing_kernel_name(int k) { return (k >= 0 && k < KID_COUNT) ? kKernelNames[k] : ""; }
int genie_timing_collect(double* total_ms, int64_t* launches, int reset) {
    std::vector<TimingSlot*> pend;
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        pend.swap(g_timing_pending);
    }
    for (TimingSlot* s : pend) {
        float ms = 0.f;
        GENIE_CUDA_CHECK(cudaEventSynchronize(s->end));
        GENIE_CUDA_CHECK(cudaEventElapsedTime(&ms, s->beg, s->end));
        g_timing_ms[s->kid] += (double)ms;
        g_timing_n[s->kid] += 1;
    }
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        for (TimingSlot* s : pend) g_timing_free.push_back(s);
    }
    for (int k = 0; k < KID_COUNT; ++k) {
        if (total_ms) total_ms[k] = g_timing_ms[k];
        if (launches) launches[k] = g_timing_n[k];
        if (reset) {
            g_timing_ms[k] = 0.0;
            g_timing_n[k] = 0;
        }
    }
    return GENIE_OK;
}

int genie_plan_set_edge_terms(genie_plan_t* plan, const float* edge_sta_dev, const float* edge_src_dev) {
    if (!plan || ((edge_sta_dev == nullptr) != (edge_src_dev == nullptr))) {
        set_error("genie_plan_set_edge_terms: null plan, or only one of the two tables given");
        return GENIE_ERR_INVALID;
    }
    plan->edge_sta = edge_sta_dev;
    plan->edge_src = edge_src_dev;
    return GENIE_OK;
}

int genie_plan_set_init_terms(genie_plan_t* plan, const float* init_sta_dev, const float* init_src_dev) {
    if (!plan || (init_sta_dev == nullptr && init_src_dev != nullptr) ||
        (init_sta_dev != nullptr && (plan->g.mode == GENIE_GRAPH_CARTESIAN) != (init_src_dev != nullptr))) {
        set_error("genie_plan_set_init_terms: CARTESIAN plans take a station and a grid table, EXPLICIT plans one per-node table");
        return GENIE_ERR_INVALID;
    }
    plan->init_sta = init_sta_dev;
    plan->init_src = init_src_dev;
    return GENIE_OK;
}

int genie_plan_set_storage(genie_plan_t* plan, int32_t storage) {
    if (!plan || (storage != GENIE_STORAGE_FP32 && storage != GENIE_STORAGE_BF16)) {
        set_error("genie_plan_set_storage: bad argument");
        return GENIE_ERR_INVALID;
    }
    if (storage == GENIE_STORAGE_BF16 && (!split_supported(plan) || plan->g.n_sta_tiles > 32)) {
        set_error("genie_plan_set_storage: bf16 storage needs a CARTESIAN plan with tiling tables");
        return GENIE_ERR_UNSUPPORTED;
    }
    plan->storage = storage;
    return GENIE_OK;
}

int genie_debug_trace(int64_t* trace_dev, int tiles) {
    set_s1_trace(reinterpret_cast<long long*>(trace_dev), trace_dev ? tiles : 0);
    return GENIE_OK;
}

int genie_plan_create(const genie_graph_desc_t* d, genie_plan_t** out) {
    if (!d || !out) {
        set_error("genie_plan_create: null argument");
        return GENIE_ERR_INVALID;
    }
    if (d->mode != GENIE_GRAPH_CARTESIAN && d->mode != GENIE_GRAPH_EXPLICIT) {
        set_error("genie_plan_create: unknown graph mode");
        return GENIE_ERR_INVALID;
    }
    if (d->n_grid < 0 || d->n_prod < 0 || d->n_sta < 0) {
        set_error("genie_plan_create: negative size");
        return GENIE_ERR_INVALID;
    }
    if (d->mode == GENIE_GRAPH_CARTESIAN && (int64_t)d->n_sta * d->n_grid != d->n_prod) {
        set_error("genie_plan_create: CARTESIAN mode needs n_prod == n_sta * n_grid");
        return GENIE_ERR_INVALID;
    }
    if (d->mode == GENIE_GRAPH_EXPLICIT && d->n_prod > 0 && d->prod_grid == nullptr) {
        set_error("genie_plan_create: EXPLICIT mode needs prod_grid");
        return GENIE_ERR_INVALID;
    }
    if (d->n_prod > 0 && (!d->sta_rowptr || !d->src_rowptr)) {
        set_error("genie_plan_create: missing product-graph row pointers");
        return GENIE_ERR_INVALID;
    }
    if (d->n_grid_owned < 0 || d->n_grid_owned > d->n_grid || (d->n_grid_owned > 0 && d->mode != GENIE_GRAPH_CARTESIAN)) {
        set_error("genie_plan_create: n_grid_owned must be in [0, n_grid] and needs CARTESIAN mode");
        return GENIE_ERR_INVALID;
    }
    if (d->n_grid > 0 && (!d->grid_rowptr || !d->grid_outdeg)) {
        set_error("genie_plan_create: missing grid-graph arrays");
        return GENIE_ERR_INVALID;
    }
    genie_plan* p = new (std::nothrow) genie_plan();
    if (!p) {
        set_error("genie_plan_create: out of host memory");
        return GENIE_ERR_INVALID;
    }
    p->g = *d;
    p->n_edges_grid = -1;
    p->storage = GENIE_STORAGE_FP32;
    p->cslot = (int)(g_plan_counter.fetch_add(1, std::memory_order_relaxed) % GENIE_CSLOTS);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        set_error(std::string("genie_plan_create: no usable CUDA device: ") + cudaGetErrorString(e));
        delete p;
        return GENIE_ERR_CUDA;
    }
    *out = p;
    return GENIE_OK;
}

void genie_plan_destroy(genie_plan_t* plan) { delete plan; }

size_t genie_plan_workspace_bytes(const genie_plan_t* plan) {
    if (!plan) return 0;
    return carve_workspace(plan, nullptr).bytes;
}

size_t genie_frontend_packed_floats(void) { return (size_t)PACKED_FLOATS; }

int genie_frontend_pack_weights(const genie_frontend_weights_t* w, float* packed_dev, void* stream) {
    if (!w || !packed_dev) {
        set_error("genie_frontend_pack_weights: null argument");
        return GENIE_ERR_INVALID;
    }
    const void* const* ptrs = reinterpret_cast<const void* const*>(w);
    for (size_t i = 0; i < sizeof(*w) / sizeof(void*); ++i) {
        if (ptrs[i] == nullptr) {
            set_error("genie_frontend_pack_weights: null weight pointer at slot " + std::to_string(i));
            return GENIE_ERR_INVALID;
        }
    }
    return launch_pack_weights(w, packed_dev, static_cast<cudaStream_t>(stream));
}

int genie_input_scatter_fwd(const genie_plan_t* plan, const genie_input_params_t* prm, const double* picks_dev,
                            int64_t n_picks, const int32_t* sta_perm_dev, const int32_t* ind_use_dev,
                            const float* trv_times_dev, const int32_t* node_sta_dev, const int32_t* node_grid_dev,
                            float* series_dev, float* slice_out_dev, float* mask_out_dev, int64_t* time_bin_out_dev,
                            void* stream) {
    if (!plan || !prm || !sta_perm_dev || !ind_use_dev || !trv_times_dev || !series_dev || !slice_out_dev ||
        !mask_out_dev || (n_picks > 0 && !picks_dev)) {
        set_error("genie_input_scatter_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    if (prm->n_ts < 2 || prm->n_extra < 0 || prm->n_sta_use <= 0 || prm->n_locs <= 0 || !(prm->dt > 0.0)) {
        set_error("genie_input_scatter_fwd: bad input parameters");
        return GENIE_ERR_INVALID;
    }
    const bool have_nodes = node_sta_dev != nullptr && node_grid_dev != nullptr;
    if ((node_sta_dev != nullptr) != (node_grid_dev != nullptr) ||
        (plan->g.mode == GENIE_GRAPH_EXPLICIT && !have_nodes)) {
        set_error("genie_input_scatter_fwd: node_sta/node_grid must both be given (required in EXPLICIT mode)");
        return GENIE_ERR_INVALID;
    }
    if (plan->g.mode == GENIE_GRAPH_CARTESIAN && !have_nodes && plan->g.n_sta != prm->n_sta_use) {
        set_error("genie_input_scatter_fwd: n_sta_use differs from the plan's n_sta");
        return GENIE_ERR_INVALID;
    }
    return launch_input_scatter(plan, prm, picks_dev, n_picks, sta_perm_dev, ind_use_dev, trv_times_dev, node_sta_dev,
                                node_grid_dev, series_dev, slice_out_dev, mask_out_dev, time_bin_out_dev,
                                static_cast<cudaStream_t>(stream));
}

size_t genie_heads_packed_floats(void) { return (size_t)HD_FLOATS; }

int genie_heads_layout(int32_t* offsets_out, int n) {
    static const int32_t k[] = {HD_SD_W, HD_SD_B, HD_SD_SL, HD_TA_WC1, HD_TA_BC1, HD_TA_WV1, HD_TA_BV1, HD_TA_WV2, HD_TA_BV2,
                                HD_TA_WP1, HD_TA_BP1, HD_TA_WP2, HD_TA_BP2, HD_TA_SL, HD_SA_WQ, HD_SA_BQ, HD_SA_WC, HD_SA_BC,
                                HD_SA_WV, HD_SA_BV, HD_SA_WP, HD_SA_BP, HD_SA_SL};
    const int count = (int)(sizeof(k) / sizeof(k[0]));
    if (!offsets_out || n < count) {
        set_error("genie_heads_layout: need room for " + std::to_string(count) + " offsets");
        return GENIE_ERR_INVALID;
    }
    for (int i = 0; i < count; ++i) offsets_out[i] = k[i];
    return GENIE_OK;
}

int genie_heads_grid_fwd(const float* heads_packed_dev, const float* fold_dev, int n_t, const float* x_spatial_dev, int ld_x,
                         int n_grid, float* y_out_dev, float* proj_out_dev, void* stream) {
    if (!heads_packed_dev || !fold_dev || !x_spatial_dev || !y_out_dev || n_t < 0 || n_grid < 0 || ld_x < 30) {
        set_error("genie_heads_grid_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    return launch_heads_grid(heads_packed_dev, fold_dev, n_t, x_spatial_dev, ld_x, n_grid, y_out_dev, proj_out_dev,
                             static_cast<cudaStream_t>(stream));
}

int genie_heads_query_fwd(const float* heads_packed_dev, const float* fold_dev, int n_t, const float* x_spatial_dev, int ld_x,
                          const float* x_context_dev, const float* x_query_dev, const int64_t* nbr_dev, int k_nbr, int n_query,
                          float scale_rel, float* x_out_dev, const float* proj_dev, void* stream) {
    if (!heads_packed_dev || !fold_dev || !x_spatial_dev || !x_context_dev || !x_query_dev || !nbr_dev || !x_out_dev ||
        n_t < 0 || n_query < 0 || ld_x < 30 || !(scale_rel > 0.f)) {
        set_error("genie_heads_query_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    return launch_heads_query(heads_packed_dev, fold_dev, n_t, x_spatial_dev, ld_x, x_context_dev, x_query_dev, nbr_dev, k_nbr,
                              n_query, scale_rel, x_out_dev, proj_dev, static_cast<cudaStream_t>(stream));
}

int genie_input_nearest_fwd(const genie_nearest_params_t* prm, const double* times_all_dev, const double* times_p_dev,
                            const double* times_s_dev, const int32_t* ind_use_dev, const float* trv_times_dev,
                            float* slice_out_dev, float* mask_out_dev, void* stream) {
    if (!prm || !ind_use_dev || !trv_times_dev || !slice_out_dev || !mask_out_dev || (prm->n_all > 0 && !times_all_dev) ||
        (prm->n_p > 0 && !times_p_dev) || (prm->n_s > 0 && !times_s_dev)) {
        set_error("genie_input_nearest_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    if (prm->n_batch < 0 || prm->n_grid < 0 || prm->n_sta_use <= 0 || prm->n_locs <= 0 || prm->n_all < 0 || prm->n_p < 0 ||
        prm->n_s < 0 || !(prm->kernel_sig_t > 0.0)) {
        set_error("genie_input_nearest_fwd: bad parameters");
        return GENIE_ERR_INVALID;
    }
    return launch_input_nearest(prm, times_all_dev, times_p_dev, times_s_dev, ind_use_dev, trv_times_dev, slice_out_dev,
                                mask_out_dev, static_cast<cudaStream_t>(stream));
}

// ---- product-graph message passing of the training path -----------------------------------------------------------------------
int genie_kron_spmm_fwd(int mode, int n_sta, int n_grid, int64_t n_prod, const int64_t* rowptr_dev, const int32_t* col_dev,
                        const float* val_dev, const float* x_dev, int ld_x, int n_ch, float* out_dev, int ld_out, void* stream) {
    if (mode < 0 || mode > 2 || n_prod < 0 || n_ch < 0 || ld_x < n_ch || ld_out < n_ch || !rowptr_dev ||
        (mode != 2 && (n_sta <= 0 || n_grid <= 0 || n_prod != (int64_t)n_sta * n_grid)) ||
        (n_prod > 0 && n_ch > 0 && (!x_dev || !out_dev))) {
        set_error("genie_kron_spmm_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return launch_kron_spmm(mode, n_sta, n_prod, rowptr_dev, col_dev, val_dev, x_dev, ld_x, n_ch, out_dev, ld_out, sms,
                            static_cast<cudaStream_t>(stream));
}

// ---- per-node dense layers of the training path ----------------------------------------------------------------------------
static int current_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
int genie_node_mlp_partial_rows(void) { return mlp_partial_rows(current_sm_count()); }
int genie_node_mlp_fwd(const genie_mlp_desc_t* desc, float* y_dev, int32_t ld_y, uint32_t* neg_mask_dev, void* stream) {
    return launch_node_mlp_fwd(desc, y_dev, ld_y, neg_mask_dev, current_sm_count(), static_cast<cudaStream_t>(stream));
}
int genie_node_mlp_bwd(const genie_mlp_desc_t* desc, const float* y_dev, int32_t ld_y, const uint32_t* neg_mask_dev,
                       const float* gy_dev, int32_t ld_gy, float* const* gx_dev, const int32_t* ld_gx, float* partial_dev,
                       void* stream) {
    return launch_node_mlp_bwd(desc, y_dev, ld_y, neg_mask_dev, gy_dev, ld_gy, gx_dev, ld_gx, partial_dev, current_sm_count(),
                               static_cast<cudaStream_t>(stream));
}

// ---- output stacking of the streaming loop ------------------------------------------------------------
int genie_stack_output_fwd(const float* x_dev, int n_query, int n_t, int n_use, const int32_t* col_dev, float scale,
                           float* out_dev, int64_t ld_out, void* stream) {
    if (n_query < 0 || n_t < 0 || n_use < 0 || n_use > n_t || ld_out < 0 ||
        ((int64_t)n_query * n_use > 0 && (!x_dev || !col_dev || !out_dev))) {
        set_error("genie_stack_output_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    return launch_stack_output(x_dev, n_query, n_t, n_use, col_dev, scale, out_dev, ld_out, static_cast<cudaStream_t>(stream));
}

// ---- device kNN -------------------------------------------------------------------------------------
int genie_knn_fwd(const float* x_dev, int n_x, const float* y_dev, int n_y, int k, int64_t* idx_out_dev, void* stream) {
    if (n_x < 0 || n_y < 0 || k < 1 || k > 32 || k > n_x || (n_x > 0 && !x_dev) || (n_y > 0 && (!y_dev || !idx_out_dev))) {
        set_error("genie_knn_fwd: bad argument (1 <= k <= min(32, n_x))");
        return GENIE_ERR_INVALID;
    }
    return launch_knn(x_dev, n_x, y_dev, n_y, k, idx_out_dev, static_cast<cudaStream_t>(stream));
}

// ---- association branch ---------------------------------------------------------------------------
size_t genie_assoc_packed_floats(void) { return assoc_packed_floats(); }

int genie_assoc_layout(int32_t* offsets_out, int n) { return assoc_layout(offsets_out, n); }

int genie_assoc_set_terms(genie_plan_t* plan, const float* init_sta_dev, const float* init_src_dev, const float* edge_sta_dev,
                          const float* edge_src_dev) {
    if (!plan || (init_sta_dev == nullptr && init_src_dev != nullptr) ||
        (init_sta_dev != nullptr && (plan->g.mode == GENIE_GRAPH_CARTESIAN) != (init_src_dev != nullptr)) ||
        ((edge_sta_dev == nullptr) != (edge_src_dev == nullptr))) {
        set_error("genie_assoc_set_terms: table combination does not match the plan (see genie_plan_set_init_terms / "
                  "genie_plan_set_edge_terms)");
        return GENIE_ERR_INVALID;
    }
    plan->assoc_init_sta = init_sta_dev;
    plan->assoc_init_src = init_src_dev;
    plan->assoc_edge_sta = edge_sta_dev;
    plan->assoc_edge_src = edge_src_dev;
    return GENIE_OK;
}

size_t genie_assoc_workspace_bytes(const genie_plan_t* plan) {
    if (!plan) return 0;
    return carve_assoc_workspace(plan, nullptr).bytes;
}

int genie_assoc_product_fwd(const genie_plan_t* plan, const float* assoc_packed_dev, const float* x_spatial_dev, int ld_x,
                            const float* y_dev, int n_t, float mask_thresh, const float* edge_attr_dev,
                            const float* x_latent_dev, const float* mask_dev, void* assoc_workspace_dev,
                            float* s0_out_dev, float* mask_out_dev, float** s_rows_out, void* stream) {
    if (!plan || !assoc_packed_dev || !x_spatial_dev || !y_dev || !edge_attr_dev || !x_latent_dev || !mask_dev ||
        !assoc_workspace_dev || ld_x < 30 || n_t < 0) {
        set_error("genie_assoc_product_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    const AssocWorkspace w = carve_assoc_workspace(plan, assoc_workspace_dev);
    if (s_rows_out) *s_rows_out = w.tr;
    return launch_assoc_product(plan, assoc_packed_dev, x_spatial_dev, ld_x, y_dev, n_t, mask_thresh, edge_attr_dev,
                                x_latent_dev, mask_dev, w, s0_out_dev, mask_out_dev, static_cast<cudaStream_t>(stream));
}

int genie_assoc_collapse_fwd(const float* assoc_packed_dev, const float* s_rows_dev, int64_t n_prod, const int64_t* edges_p_dev,
                             const int64_t* edges_s_dev, int64_t n_edges, const float* tlatent_dev, const float* tpick_dev,
                             const int64_t* ipick_dev, const float* phase_dev, int n_arv, int n_sta, int l_dt, int k_infer,
                             float dt0, float dt_step, float eps, float* arrival_out_dev, void* stream) {
    if (!assoc_packed_dev || !s_rows_dev || !edges_p_dev || !edges_s_dev || !tlatent_dev || !arrival_out_dev || n_arv < 0 ||
        (n_arv > 0 && (!tpick_dev || !ipick_dev || !phase_dev)) || l_dt < 2 || k_infer < 1 || !(dt_step > 0.f) ||
        !(eps > 0.f) || n_edges != (int64_t)n_sta * l_dt * k_infer) {
        set_error("genie_assoc_collapse_fwd: bad argument (the pointer tables must hold n_sta * len(dt_partition) * k_infer "
                  "entries, module.py:624)");
        return GENIE_ERR_INVALID;
    }
    return launch_assoc_collapse(assoc_packed_dev, s_rows_dev, n_prod, edges_p_dev, edges_s_dev, tlatent_dev, tpick_dev,
                                 ipick_dev, phase_dev, n_arv, l_dt, k_infer, dt0, dt_step, eps, arrival_out_dev,
                                 static_cast<cudaStream_t>(stream));
}

int genie_data_aggregation_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev,
                               const float* mask_dev, float* x_latent_out_dev, void* workspace_dev, void* stream) {
    if (!plan || !packed_dev || !slice_dev || !mask_dev || !x_latent_out_dev || !workspace_dev) {
        set_error("genie_data_aggregation_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace w = carve_workspace(plan, workspace_dev);
    int rc;
    if ((rc = launch_da_layers01(plan, packed_dev, slice_dev, mask_dev, w, st))) return rc;
    return launch_da_layer2(plan, packed_dev, w, mask_dev, nullptr, x_latent_out_dev, nullptr, 0, st);
}

int genie_bipartite_readin_fwd(const genie_plan_t* plan, const float* packed_dev, const float* x_latent_dev,
                               const float* edge_attr_dev, const float* mask_dev, float* out_dev, void* workspace_dev,
                               void* stream) {
    if (!plan || !packed_dev || !x_latent_dev || !edge_attr_dev || !mask_dev || !out_dev || !workspace_dev) {
        set_error("genie_bipartite_readin_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace w = carve_workspace(plan, workspace_dev);
    GENIE_CUDA_CHECK(cudaMemsetAsync(w.xg, 0, (size_t)plan->g.n_grid * 32 * sizeof(float), st));
    int rc;
    if ((rc = launch_da_layer2_readin(plan, packed_dev, L2_READIN, nullptr, nullptr, nullptr, x_latent_dev, nullptr,
                                      edge_attr_dev, mask_dev, w.xg, st)))
        return rc;
    return launch_readin_finalize(plan, packed_dev, w.xg, out_dev, 15, st);
}

int genie_da_layer1_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev, const float* mask_dev,
                        void* workspace_dev, void* stream) {
    if (!plan || !packed_dev || !slice_dev || !mask_dev || !workspace_dev) {
        set_error("genie_da_layer1_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    Workspace w = carve_workspace(plan, workspace_dev);
    return launch_da_layers01(plan, packed_dev, slice_dev, mask_dev, w, static_cast<cudaStream_t>(stream));
}

// ---- grid sharding: halo rows over peer memory ------------------------------------------------------------------------------
int genie_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
    if (!ptr_out || !handle_out || bytes == 0) {
        set_error("genie_peer_alloc: bad argument");
        return GENIE_ERR_INVALID;
    }
    void* ptr = nullptr;
    GENIE_CUDA_CHECK(cudaMalloc(&ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        set_error(std::string("genie_peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
        return GENIE_ERR_CUDA;
    }
    static_assert(sizeof(h) == GENIE_PEER_HANDLE_BYTES, "handle size");
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = ptr;
    return GENIE_OK;
}

int genie_peer_open(const unsigned char* handle, void** ptr_out) {
    if (!handle || !ptr_out) {
        set_error("genie_peer_open: bad argument");
        return GENIE_ERR_INVALID;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GENIE_CUDA_CHECK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GENIE_OK;
}

int genie_peer_close(void* ptr) {
    if (ptr) GENIE_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return GENIE_OK;
}

int genie_peer_free(void* ptr) {
    if (ptr) GENIE_CUDA_CHECK(cudaFree(ptr));
    return GENIE_OK;
}

int genie_plan_set_halo_export(genie_plan_t* plan, const int32_t* exp_ptr_dev, const int32_t* exp_peer_dev,
                               const int32_t* exp_row_dev, float* const* peer_base_dev, const float* halo_vb_dev) {
    if (!plan) {
        set_error("genie_plan_set_halo_export: null plan");
        return GENIE_ERR_INVALID;
    }
    if (exp_ptr_dev == nullptr && halo_vb_dev == nullptr) {
        plan->exp_ptr = plan->exp_peer = plan->exp_row = nullptr;
        plan->peer_base = nullptr;
        plan->halo_vb = nullptr;
        return GENIE_OK;
    }
    if (!exp_ptr_dev || !exp_peer_dev || !exp_row_dev || !peer_base_dev || !halo_vb_dev || plan->g.n_grid_owned <= 0 ||
        !split_supported(plan)) {
        set_error("genie_plan_set_halo_export: needs all five tables and a CARTESIAN plan with tiling tables and n_grid_owned > 0");
        return GENIE_ERR_INVALID;
    }
    plan->exp_ptr = exp_ptr_dev;
    plan->exp_peer = exp_peer_dev;
    plan->exp_row = exp_row_dev;
    plan->peer_base = peer_base_dev;
    plan->halo_vb = halo_vb_dev;
    return GENIE_OK;
}

int genie_workspace_region(const genie_plan_t* plan, void* workspace_dev, int32_t which, void** ptr_out,
                           size_t* bytes_out) {
    if (!plan || !workspace_dev || !ptr_out || !bytes_out || which != GENIE_WS_VB) {
        set_error("genie_workspace_region: bad argument");
        return GENIE_ERR_INVALID;
    }
    Workspace w = carve_workspace(plan, workspace_dev);
    *ptr_out = w.vb;
    *bytes_out = (size_t)plan->g.n_prod * LD_V * (plan->storage == GENIE_STORAGE_BF16 ? 2 : sizeof(float));
    return GENIE_OK;
}

int genie_da_layer2_readin_fwd(const genie_plan_t* plan, const float* packed_dev, const float* mask_dev,
                               const float* edge_attr_dev, float* x_latent_out_dev, float* readin_out_dev,
                               void* workspace_dev, void* stream) {
    if (!plan || !packed_dev || !mask_dev || !edge_attr_dev || !readin_out_dev || !workspace_dev) {
        set_error("genie_da_layer2_readin_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    Workspace w = carve_workspace(plan, workspace_dev);
    return launch_da_layer2(plan, packed_dev, w, mask_dev, edge_attr_dev, x_latent_out_dev, readin_out_dev, 15,
                            static_cast<cudaStream_t>(stream));
}

int genie_spatial_aggregation_fwd(const genie_plan_t* plan, const float* packed_dev, int32_t layer, const float* x_dev,
                                  const float* pos_dev, float scale_rel, float* out_dev, void* workspace_dev,
                                  void* stream) {
    if (!plan || !packed_dev || !x_dev || !pos_dev || !out_dev || !workspace_dev) {
        set_error("genie_spatial_aggregation_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    Workspace w = carve_workspace(plan, workspace_dev);
    return launch_spatial_aggregation(plan, packed_dev, layer, x_dev, layer == 0 ? 15 : 30, pos_dev, scale_rel, w.px,
                                      w.partial, out_dev, 30, static_cast<cudaStream_t>(stream));
}

int genie_window_fwd(const genie_plan_t* plan, const float* packed_dev, const genie_window_params_t* wp_dev,
                     int64_t max_window_picks, int32_t n_extra, const double* picks_dev, const int32_t* sta_perm_dev,
                     const int32_t* ind_use_dev, const float* trv_times_dev, float* series_dev, int32_t n_ts_max,
                     const float* edge_attr_dev, const float* pos_dev, float scale_rel, float* slice_out_dev,
                     float* mask_out_dev, float* x_latent_out_dev, float* readin_out_dev, float* x_spatial_out_dev,
                     void* workspace_dev, void* stream) {
    if (!plan || !packed_dev || !wp_dev || !sta_perm_dev || !ind_use_dev || !trv_times_dev || !series_dev || !edge_attr_dev ||
        !pos_dev || !x_spatial_out_dev || !workspace_dev || (max_window_picks > 0 && !picks_dev) || max_window_picks < 0 ||
        n_extra < 0 || n_ts_max < 2 || ((slice_out_dev == nullptr) != (mask_out_dev == nullptr))) {
        set_error("genie_window_fwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    if (!split_supported(plan) || plan->g.n_sta_tiles > 32 || plan->g.n_prod >= (int64_t)0x7fffffff) {
        set_error("genie_window_fwd: needs a CARTESIAN plan with tiling tables (use genie_input_scatter_fwd + genie_frontend_fwd)");
        return GENIE_ERR_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace w = carve_workspace(plan, workspace_dev);
    WindowParamSrc ws;
    ws.host = genie_input_params_t();
    ws.dev = wp_dev;
    ws.n_picks = 0;
    int rc;
    if ((rc = launch_input_series(ws, max_window_picks, picks_dev, sta_perm_dev, series_dev,
                                  (size_t)2 * plan->g.n_sta * n_ts_max * sizeof(float), n_extra, st)))
        return rc;
    if ((rc = launch_da_layers01_window(plan, packed_dev, ws, ind_use_dev, trv_times_dev, series_dev, slice_out_dev,
                                        mask_out_dev, w, st)))
        return rc;
    float* r = readin_out_dev ? readin_out_dev : w.r;
    const int ld_r = readin_out_dev ? 15 : 16;
    if ((rc = launch_da_layer2(plan, packed_dev, w, nullptr, edge_attr_dev, x_latent_out_dev, r, ld_r, st))) return rc;
    if ((rc = launch_spatial_aggregation(plan, packed_dev, 0, r, ld_r, pos_dev, scale_rel, w.px, w.partial, w.sa_a, 32, st)))
        return rc;
    if ((rc = launch_spatial_aggregation(plan, packed_dev, 1, w.sa_a, 32, pos_dev, scale_rel, w.px, w.partial, w.sa_b, 32, st)))
        return rc;
    return launch_spatial_aggregation(plan, packed_dev, 2, w.sa_b, 32, pos_dev, scale_rel, w.px, w.partial, x_spatial_out_dev,
                                      30, st);
}

int genie_frontend_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev,
                       const float* mask_dev, const float* edge_attr_dev, const float* pos_dev, float scale_rel,
                       float* x_latent_out_dev, float* readin_out_dev, float* x_spatial_out_dev, void* workspace_dev,
                       void* stream) {
    if (!plan || !packed_dev || !slice_dev || !mask_dev || !edge_attr_dev || !pos_dev || !x_spatial_out_dev ||
        !workspace_dev) {
        set_error("genie_frontend_fwd: null argument");
        return GENIE_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace w = carve_workspace(plan, workspace_dev);
    int rc;
    if ((rc = launch_da_layers01(plan, packed_dev, slice_dev, mask_dev, w, st))) return rc;
    float* r = readin_out_dev ? readin_out_dev : w.r;
    const int ld_r = readin_out_dev ? 15 : 16;
    if ((rc = launch_da_layer2(plan, packed_dev, w, mask_dev, edge_attr_dev, x_latent_out_dev, r, ld_r, st))) return rc;
    if ((rc = launch_spatial_aggregation(plan, packed_dev, 0, r, ld_r, pos_dev, scale_rel, w.px, w.partial, w.sa_a, 32,
                                         st)))
        return rc;
    if ((rc = launch_spatial_aggregation(plan, packed_dev, 1, w.sa_a, 32, pos_dev, scale_rel, w.px, w.partial, w.sa_b,
                                         32, st)))
        return rc;
    return launch_spatial_aggregation(plan, packed_dev, 2, w.sa_b, 32, pos_dev, scale_rel, w.px, w.partial,
                                      x_spatial_out_dev, 30, st);
}

}  // extern "C"
