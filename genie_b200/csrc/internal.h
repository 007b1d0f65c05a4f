// Internal declarations shared by the translation units of libgenie_b200.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>

#include "../../include/genie_b200.h"
#include "layout.h"

// Weight blocks that kernels read as constant-bank operands (init_trns, read-in fc1) live in per-plan slots of a __constant__
// array, refreshed in stream order before each launch: plans (models, e.g. the reference's mz_list, one per grid) that run
// on different streams do not share a slot unless more than GENIE_CSLOTS plans are alive at once.
constexpr int GENIE_CSLOTS = 4;     // the slot is a template parameter of the kernels that use it

struct genie_plan {
    genie_graph_desc_t g;
    int sm_count;
    int cslot;              // constant-bank slot of this plan
    int storage;            // GENIE_STORAGE_* of the gathered intermediate rows
    int64_t n_edges_grid;   // number of grid-graph edges (host copy of grid_rowptr[G]), fetched lazily
    // Edge-feature model (genie_plan_set_edge_terms): per-node additive terms of layer 1, NULL = off.
    const float* edge_sta;  // [S or P][GENIE_EDGE_TERM_LD]
    const float* edge_src;  // [G or P][GENIE_EDGE_TERM_LD]
    // use_absolute_pos (genie_plan_set_init_terms): additive terms of init_trns before its activation, NULL = off.
    const float* init_sta;  // [S][32] (CARTESIAN) or [P][32] (EXPLICIT, init_src NULL)
    const float* init_src;  // [G][32] or NULL
    // the same two kinds of tables for the association branch (genie_assoc_set_terms), NULL = off
    const float* assoc_init_sta;
    const float* assoc_init_src;
    const float* assoc_edge_sta;
    const float* assoc_edge_src;
    // Grid-sharded plans (genie_plan_set_halo_export): the layer-1 station pass stores the layer-2 message rows v_b of the owned
    // grid nodes that peers hold as halo straight into the peers' landing buffers (NVLink peer stores), and the layer-2 source
    // pass reads this rank's halo rows from its own landing buffer.  NULL = off.
    const int32_t* exp_ptr;    // [n_grid_owned + 1] CSR over the owned grid nodes
    const int32_t* exp_peer;   // [n_exports] peer slot of every export
    const int32_t* exp_row;    // [n_exports] row (halo position) of the node in that peer's landing buffer
    float* const* peer_base;   // [n_peers] device array of the peers' landing-buffer base addresses
    const float* halo_vb;      // this rank's landing buffer: [n_grid - n_grid_owned][S][16 fp32 | 16 bf16]
};

// Device-side view of the product graph.  CARTESIAN: node i = g*S + s; sta neighbours g*S + col, src neighbours
// col*S + s.  EXPLICIT: neighbours are col directly.
struct GraphView {
    int mode;
    int S;
    int G;
    int64_t P;
    const int64_t* sta_rowptr;
    const int32_t* sta_col;
    const int64_t* src_rowptr;
    const int32_t* src_col;
    const int32_t* prod_grid;
    const int32_t* grid_order;
    const float* edge_sta;   // edge-feature model: rows indexed by the station (CARTESIAN) / product node (EXPLICIT), or NULL
    const float* edge_src;   //                     rows indexed by the grid node (CARTESIAN) / product node (EXPLICIT), or NULL
};

inline GraphView make_view(const genie_plan* p) {
    GraphView v;
    v.mode = p->g.mode;
    v.S = p->g.n_sta;
    v.G = p->g.n_grid;
    v.P = p->g.n_prod;
    v.sta_rowptr = p->g.sta_rowptr;
    v.sta_col = p->g.sta_col;
    v.src_rowptr = p->g.src_rowptr;
    v.src_col = p->g.src_col;
    v.prod_grid = p->g.prod_grid;
    v.grid_order = p->g.grid_order;
    v.edge_sta = p->edge_sta;
    v.edge_src = p->edge_src;
    return v;
}

// Workspace carve-up (floats).  All regions 256-byte aligned.
struct Workspace {
    float* tr0;      // [P][32]   layer-0 features; re-used for mean_src(v_b) [P][16] once layer 1 is done
    float* msrc;     // [P][32]   mean over source neighbours of the layer-0 features (split kernels only)
    float* zc;       // [P][32]
    float* va;       // [P][16]
    float* vb;       // [P][16]
    float* xg;       // [G][32]   read-in accumulator (sum over stations)
    float* r;        // [G][16]   read-in output (15 used)
    float* px;       // [G][32]   SpatialAggregation: W_x x_j
    float* sa_a;     // [G][32]   ping
    float* sa_b;     // [G][32]   pong
    float* partial;  // [1024][8] per-CTA partial sums of the global feature
    size_t bytes;
};
Workspace carve_workspace(const genie_plan* p, void* base);

void set_error(const std::string& msg);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device (context): a launcher sets it once per device.
// need() is true until mark() has been called for the CURRENT device; safe from several host threads (worst case the
// attribute is set twice).
struct PerDeviceOnce {
    std::atomic<unsigned char> done[64];
    PerDeviceOnce() { for (auto& d : done) d.store(0); }
    static int dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < 64) ? d : 0; }
    bool need() const { return done[dev()].load(std::memory_order_acquire) == 0; }
    void mark() { done[dev()].store(1, std::memory_order_release); }
};
void count_launch(int n = 1);

// Kernel ids of the optional per-kernel device timing (genie_timing_* in the C-ABI).
enum KernelId {
    KID_PACK = 0,
    KID_INPUT_SERIES,
    KID_INPUT_GATHER,
    KID_DA_INIT,
    KID_DA_LAYER1,
    KID_DA_LAYER1_TC,
    KID_DA_LAYER2_READIN,
    KID_READIN_FINALIZE,
    KID_SA_PRE,
    KID_SA_MAIN,
    KID_SRC_MEAN32,
    KID_SRC_MEAN16,
    KID_DA_LAYER1_S,
    KID_DA_LAYER2_S,
    KID_HEADS_GRID,
    KID_HEADS_QUERY,
    KID_ASSOC_GRID_PRE,
    KID_ASSOC_INIT,
    KID_ASSOC_LAYER1,
    KID_ASSOC_LAYER2,
    KID_ASSOC_COLLAPSE,
    KID_KNN,
    KID_STACK,
    KID_KRON_SPMM,
    KID_NODE_MLP_FWD,
    KID_NODE_MLP_BWD,
    KID_COUNT
};
// Brackets one kernel launch with cudaEvents on its stream when timing is enabled (no-op otherwise).
struct TimedLaunch {
    TimedLaunch(int kid, cudaStream_t st);
    ~TimedLaunch();
    int kid;
    cudaStream_t st;
    void* slot;
};

#define GENIE_CUDA_CHECK(expr)                                                                         \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
            return GENIE_ERR_CUDA;                                                                      \
        }                                                                                               \
    } while (0)

#define GENIE_LAUNCH_CHECK()                                                                            \
    do {                                                                                                \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess) {                                                                        \
            set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e));                  \
            return GENIE_ERR_CUDA;                                                                      \
        }                                                                                               \
        count_launch();                                                                                 \
    } while (0)

// ---- launchers (each returns a GENIE_* status) -----------------------------------------------------------------------
int launch_pack_weights(const genie_frontend_weights_t* w, float* packed, cudaStream_t st);

// tc_plan: the plan is eligible for the tensor-core layer-1 kernel (da_tc_supported).  Whether that kernel or the generic
// one does the work is then decided ON THE DEVICE from the packed weights (layout.h TCS_OK): both are launched, one exits.
bool da_tc_supported(const genie_plan* p);
int launch_da_init(const genie_plan* p, const float* packed, const float* slice, const float* mask, float* tr0,
                   bool tc_plan, cudaStream_t st);
struct WindowParamSrc;
int launch_da_init_fused(const genie_plan* p, const float* packed, const WindowParamSrc& ws, const int32_t* ind_use,
                         const float* trv, const float* series, float* slice_out, float* mask_out, float* tr0,
                         bool tc_plan, cudaStream_t st);
int launch_input_series(const WindowParamSrc& ws, int64_t max_picks, const double* picks, const int32_t* sta_perm, float* series,
                        size_t series_bytes, int n_extra, cudaStream_t st);
int launch_da_layer1(const genie_plan* p, const float* packed, const float* tr0, const float* mask, float* zc, float* va,
                     float* vb, bool tc_plan, cudaStream_t st);
int launch_da_layer1_tc(const genie_plan* p, const float* packed, const float* pfeat, const float* mask, float* zc,
                        float* va, float* vb, cudaStream_t st);
// mode bits for the layer-2 / read-in kernel
enum { L2_GATHER = 1, L2_STORE_LATENT = 2, L2_READIN = 4 };
int launch_da_layer2_readin(const genie_plan* p, const float* packed, int mode, const float* zc, const float* va,
                            const float* vb, const float* latent_in, float* latent_out, const float* edge_attr,
                            const float* mask, float* xg, cudaStream_t st);
int launch_readin_finalize(const genie_plan* p, const float* packed, const float* xg, float* out, int ld_out,
                           cudaStream_t st);
int launch_spatial_aggregation(const genie_plan* p, const float* packed, int layer, const float* x, int ld_x,
                               const float* pos, float scale_rel, float* px, float* partial, float* out, int ld_out,
                               cudaStream_t st);
// split source-pass / station-pass kernels (plans with tiling tables; src_mean_kernels.cu, da_s1_kernel.cu, da_s2_kernel.cu)
bool split_supported(const genie_plan* p);
// out[g,s,:] = mean_{g' in N_src(g)} X[g',s,:], rows of `width` floats (32 or 16); gate: optional device flag, 0 = skip
// halo != NULL: rows of grid nodes >= n_grid_owned are read from `halo` (row 0 = node n_grid_owned) instead of X
int launch_src_mean(const genie_plan* p, int width, const float* X, float* out, const float* gate, cudaStream_t st,
                    int storage = GENIE_STORAGE_FP32, const float* halo = nullptr);
int launch_da_layer1_s(const genie_plan* p, const float* packed, const float* pfeat, const float* msrc,
                       const float* mask, float* zc, float* va, float* vb, cudaStream_t st);
void set_s1_trace(long long* buf, int tiles);
int launch_da_layer2_s(const genie_plan* p, const float* packed, const float* zc, const float* va, const float* m2,
                       const float* mask, const float* edge_attr, float* latent_out, float* out, int ld_out,
                       cudaStream_t st);
int launch_heads_grid(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, int G, float* y,
                      float* proj, cudaStream_t st);
int launch_heads_query(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, const float* x_context,
                       const float* x_query, const int64_t* nbr, int k_nbr, int Q, float scale_rel, float* x_out,
                       const float* proj, cudaStream_t st);
// association branch (assoc_kernels.cu)
struct AssocWorkspace {
    float* tr;        // [P][32]   init_trns output; re-used for the branch output s [P][32] = [o1(15) 0 | o2(15) 0]
    float* a1;        // [P][32]   PReLU11(l1_t1_1 tr)
    float* a2;        // [P][32]   PReLU12(l1_t2_1 tr)
    float* zc;        // [P][32]
    float* va;        // [P][16]
    float* vb;        // [P][16]
    float* msrc;      // [P][32]   mean over source neighbours of a2 (plans with tiling tables)
    float* yfc1;      // [G][32]   read-out fc1, y_latent half
    float* mask_out;  // [G]
    float* t2;        // tensor-core blob of layer 1 (plans with tiling tables), layout.h T2_FLOATS + 96 floats
    size_t bytes;
};
AssocWorkspace carve_assoc_workspace(const genie_plan* p, void* base);
int launch_assoc_pack_t2(const float* assoc_packed, float* blob, cudaStream_t st);
int launch_assoc_layer1_s(const genie_plan* p, const float* blob, const float* tr, const float* a1, const float* msrc,
                          const float* mask, const float* mask_out, float* zc, float* va, float* vb, cudaStream_t st);
int launch_assoc_layer2_s(const genie_plan* p, const float* slope_dev, const float* zc, const float* va, const float* m2, float* s_out,
                          cudaStream_t st);
size_t assoc_packed_floats();
int assoc_layout(int32_t* out, int n);
int launch_assoc_product(const genie_plan* p, const float* packed, const float* x_spatial, int ld_x, const float* y, int T,
                         float thresh, const float* edge_attr, const float* x_latent, const float* mask,
                         const AssocWorkspace& w, float* s0_out, float* mask_out_copy, cudaStream_t st);
int launch_assoc_collapse(const float* packed, const float* s_rows, int64_t P, const int64_t* edges_p, const int64_t* edges_s,
                          const float* tlatent, const float* tpick, const int64_t* ipick, const float* phase_label, int n_arv,
                          int l_dt, int k_infer, float dt0, float dt_step, float eps, float* arrival, cudaStream_t st);
int launch_stack_output(const float* x, int Q, int T, int n_use, const int32_t* col, float scale, float* out, int64_t ld_out,
                        cudaStream_t st);
int launch_kron_spmm(int mode, int S, int64_t P, const int64_t* rowptr, const int32_t* col, const float* val, const float* X,
                     int ld_x, int C, float* out, int ld_o, int sm_count, cudaStream_t st);
int mlp_partial_rows(int sm_count);
int launch_node_mlp_fwd(const genie_mlp_desc_t* d, float* y, int ld_y, uint32_t* neg_mask, int sm_count, cudaStream_t st);
int launch_node_mlp_bwd(const genie_mlp_desc_t* d, const float* y, int ld_y, const uint32_t* neg_mask, const float* gy, int ld_gy,
                        float* const* gx,
                        const int* ld_gx, float* partial, int sm_count, cudaStream_t st);
int launch_knn(const float* x, int n_x, const float* y, int n_y, int k, int64_t* idx_out, cudaStream_t st);
int launch_input_nearest(const genie_nearest_params_t* prm, const double* t_all, const double* t_p, const double* t_s,
                         const int32_t* ind_use, const float* trv_times, float* slice_out, float* mask_out, cudaStream_t st);
int launch_input_scatter(const genie_plan* p, const genie_input_params_t* prm, const double* picks, int64_t n_picks,
                         const int32_t* sta_perm, const int32_t* ind_use, const float* trv_times,
                         const int32_t* node_sta, const int32_t* node_grid, float* series, float* slice_out,
                         float* mask_out, int64_t* time_bin_out, cudaStream_t st);
