/*
 * genie_b200.h — C-ABI of libgenie_b200.so: the B200 (sm_100a) product-graph front end of GENIE.
 *
 * The reference (imcbrearty/GENIE) has no FFI of its own: the hot path is reached through the Python classes of
 * Code/module.py and the helper functions of Code/process_utils.py, whose arithmetic runs inside torch_geometric /
 * torch_scatter / torch_cluster.  Every entry point below replaces one of those call sites; the citation after each
 * declaration is the reference interface it stands in for.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add (genie_b200/capi.py is that binding).
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev / every `const float*` tensor argument is a DEVICE pointer owned
 *     by the caller (torch allocates them); the library never allocates or frees device memory and never synchronises
 *     the stream.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all floating point tensors are fp32, row-major, dense; index arrays are int32 (columns) / int64 (row pointers).
 *   - return value: 0 = ok, non-zero = error; genie_last_error() returns a thread-local message.
 *   - re-entrant across distinct plans / workspaces; calls that share a workspace must be stream-ordered.  Two small
 *     weight blocks (init_trns, read-in fc1) are refreshed in __constant__ memory, in stream order, before the kernels that
 *     read them: forward calls with DIFFERENT packed weights must therefore not overlap on one device (same weights: fine).
 */
#ifndef GENIE_B200_H_
#define GENIE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENIE_B200_ABI_VERSION 8

#if defined(__GNUC__)
#define GENIE_API __attribute__((visibility("default")))
#else
#define GENIE_API
#endif

enum { GENIE_OK = 0, GENIE_ERR_INVALID = 1, GENIE_ERR_CUDA = 2, GENIE_ERR_UNSUPPORTED = 3 };

/* ---- graph description -------------------------------------------------------------------------------------------
 * Replaces the explicit int64 [2,E] edge lists the reference builds once per station set
 * (process_utils.py:718-722 extract_inputs_adjacencies, :744-849 extract_inputs_adjacencies_subgraph) and hands to
 * GCN_Detection_Network_extended.set_adjacencies (module.py:941).  All graphs are stored CSR *by destination*
 * (row i lists the message sources j of target i), which is what mean/sum aggregation needs.
 *
 * GENIE_GRAPH_CARTESIAN: product node id = g*n_sta + s (process_continuous_days.py:629).  The product edges are never
 *   materialised: sta_* is the station kNN graph A_sta_sta over n_sta nodes and src_* the grid kNN graph A_src_src
 *   over n_grid nodes; (g,s) <- (g,s') for s' in sta(s) and (g,s) <- (g',s) for g' in src(g)  (process_utils.py:720-721).
 * GENIE_GRAPH_EXPLICIT: arbitrary product graph (sub-graph mode): sta_* and src_* are CSR over the n_prod product
 *   nodes themselves (A_in_sta / A_in_src of module.py:85), prod_grid[i] is the grid node product node i feeds
 *   (A_src_in_edges.edge_index[1], module.py:229).
 * grid_* is always the grid-node graph A_src used by SpatialAggregation (module.py:243); grid_outdeg[j] is the number
 * of edges that leave grid node j (the 'global' feature of module.py:249 is a mean over edges).
 */
enum { GENIE_GRAPH_CARTESIAN = 0, GENIE_GRAPH_EXPLICIT = 1 };

typedef struct genie_graph_desc {
    int32_t mode;
    int32_t n_sta;              /* S (CARTESIAN only, else 0) */
    int32_t n_grid;             /* G */
    int32_t sta_max_deg;        /* CARTESIAN: largest in-degree of the station graph, or 0 if unknown (generic kernels) */
    int64_t n_prod;             /* P (= S*G in CARTESIAN mode) */
    const int64_t* sta_rowptr;  /* [S+1] or [P+1] */
    const int32_t* sta_col;
    const int64_t* src_rowptr;  /* [G+1] or [P+1] */
    const int32_t* src_col;
    const int64_t* grid_rowptr; /* [G+1] */
    const int32_t* grid_col;
    const int32_t* grid_outdeg; /* [G] */
    const int32_t* prod_grid;   /* [P], EXPLICIT only (NULL otherwise); must be < n_grid */
    const int32_t* grid_order;  /* [G] optional (NULL = 0..G-1): a permutation of the grid nodes in which graph neighbours
                                   are close together (e.g. reverse Cuthill-McKee of grid/src graph).  Only the ORDER in
                                   which tiles are processed follows it (L2 reuse of neighbour tiles); data layout and
                                   results do not depend on it. */
    /* Optional on-chip tiling tables of the split source-pass / station-pass kernels (CARTESIAN only; leave the
     * counts 0 to run the one-pass kernels).  They only say HOW the work is tiled; results do not depend on them.
     *   station tiles: NT compact sets of <= 128 stations with the halo of their station-graph in-neighbours, at most
     *     GENIE_TILE_ROWS_MAX staged rows each:  sta_tile_rows int32 [NT][ROWS_MAX] station id per staged row (the
     *     tile's own stations first), sta_tile_meta int32 [NT][2] = (own stations, staged rows), sta_tile_nbr uint16
     *     [NT][128][16] staged-row index of each in-neighbour (padding = ROWS_MAX, the zero row), sta_tile_invdeg fp32
     *     [NT][128] = 1 / in-degree (0 if isolated).  Every station must belong to exactly one tile.
     *   grid groups: NG compact groups of grid nodes (any sizes), grid_grp_ptr int32 [NG+1] into grid_grp_nodes int32 [G]
     *     (a permutation of the grid nodes); the source-neighbour rows of one group are re-used through L1. */
    int32_t n_sta_tiles;
    int32_t n_grid_groups;
    const int32_t* sta_tile_rows;
    const int32_t* sta_tile_meta;
    const uint16_t* sta_tile_nbr;
    const float* sta_tile_invdeg;
    const int32_t* grid_grp_ptr;
    const int32_t* grid_grp_nodes;
    /* Grid sharding (CARTESIAN only): the first n_grid_owned grid nodes are owned by this rank, the remaining ones are
     * halo copies of other ranks' nodes (their inputs are present, their layer-1 outputs arrive by exchange: see
     * genie_da_layer1_fwd).  0 = all n_grid nodes are owned.  grid_grp_nodes must then list owned nodes only. */
    int32_t n_grid_owned;
} genie_graph_desc_t;

#define GENIE_TILE_ROWS_MAX 288

typedef struct genie_plan genie_plan_t;

/* Validates `desc` (host-side checks only) and keeps a copy.  Replaces GCN_Detection_Network_extended.set_adjacencies,
 * module.py:941-961. */
GENIE_API int genie_plan_create(const genie_graph_desc_t* desc, genie_plan_t** plan_out);
GENIE_API void genie_plan_destroy(genie_plan_t* plan);
/* Edge-feature model (`use_updated_model_definition: True`: DataAggregationEdges, module.py:102-174).  Every message of
 * that model is [x_j | pos_rel(edge)], and the mean over a node's in-edges of the 4 pos_rel channels does not depend on the
 * window, so the four extra input columns of l1_t*_2 / l2_t*_2 reduce to per-node additive terms (exact, by linearity):
 *   edge_sta_dev [S (CARTESIAN) or P (EXPLICIT)][GENIE_EDGE_TERM_LD]:
 *       [0,30)  = l1_t1_2.weight[:, 60:64] . mean_{station in-edges} pos_rel_sta      (added to tr1 before activate1)
 *       [32,47) = l2_t1_2.weight[:, 90:94] . mean_{station in-edges} pos_rel_sta      (added to the l2_t1_2 output)
 *   edge_src_dev [G or P][GENIE_EDGE_TERM_LD]: the same with l1_t2_2 / l2_t2_2 and the source in-edges.
 * The weights structure then carries l1_t*_2 / l2_t*_2 WITHOUT those four columns ([30][64] / [15][94]).  The tables are
 * the caller's (they must outlive the forward calls); NULL, NULL switches the terms off.  genie_b200/module.py builds
 * them in set_adjacencies / whenever the weights change. */
#define GENIE_EDGE_TERM_LD 48
GENIE_API int genie_plan_set_edge_terms(genie_plan_t* plan, const float* edge_sta_dev, const float* edge_src_dev);
/* `use_absolute_pos: True` (module.py:913-914, 56-57): Slice gains the six channels [locs_use_cart[s] | x_temp_cuda_cart[g]] /
 * (3 scale_rel) before init_trns (4+6+4 inputs).  They do not depend on the window, so — exact, by linearity — they reduce to
 * additive terms of init_trns before its activation:
 *   CARTESIAN plans: init_sta_dev [S][32] = init_trns.weight[:, 4:7] . locs[s] / (3 scale_rel)   (30 used, padding zero)
 *                    init_src_dev [G][32] = init_trns.weight[:, 7:10] . x_temp[g] / (3 scale_rel)
 *   EXPLICIT plans:  init_sta_dev [P][32] = the sum of the two for the node's (station, grid node); init_src_dev NULL.
 * The weights structure then carries init_trns WITHOUT those six columns ([30][8]).  Tables are the caller's; NULL, NULL = off. */
GENIE_API int genie_plan_set_init_terms(genie_plan_t* plan, const float* init_sta_dev, const float* init_src_dev);
/* Bytes of caller-provided scratch the forward entry points need for this plan (intermediate node features). */
GENIE_API size_t genie_plan_workspace_bytes(const genie_plan_t* plan);

/* ---- weights -------------------------------------------------------------------------------------------------------
 * Device pointers straight into the reference's state_dict tensors (nn.Linear weight [out,in] row-major, bias [out],
 * nn.PReLU weight [1]); names are the reference's attribute names (module.py:53-83, 214-222, 231-241).
 * DataAggregation.l1_t1_1 / l1_t2_1 are unused by the reference's forward (module.py:90-91) and so absent here.
 */
typedef struct genie_linear { const float* weight; const float* bias; } genie_linear_t;

typedef struct genie_frontend_weights {
    /* DataAggregation(4, 15) — module.py:52-98 */
    genie_linear_t da_init_trns;                 /* 8  -> 30 */
    genie_linear_t da_l1_t1_2, da_l1_t2_2;       /* 64 -> 30 */
    genie_linear_t da_l2_t1_1, da_l2_t2_1;       /* 60 -> 30 */
    genie_linear_t da_l2_t1_2, da_l2_t2_2;       /* 94 -> 15 */
    const float *da_activate, *da_activate11, *da_activate12, *da_activate1, *da_activate21, *da_activate22,
        *da_activate2;
    /* BipartiteGraphOperator(30, 15, ndim_edges = 3) — module.py:214-229 */
    genie_linear_t ri_fc1;                       /* 33 -> 30 */
    genie_linear_t ri_fc2;                       /* 30 -> 15 */
    const float *ri_activate1, *ri_activate2;
    /* SpatialAggregation(15,30), (30,30), (30,30) — module.py:231-249 */
    struct {
        genie_linear_t fc1;                      /* C+8  -> 30 */
        genie_linear_t fc2;                      /* 30+C -> 30 */
        genie_linear_t fglobal;                  /* C    -> 5  */
        const float *activate1, *activate2, *activate3;
    } sa[3];
} genie_frontend_weights_t;

/* ---- storage mode of the intermediate node-feature rows -----------------------------------------------------------------
 * GENIE_STORAGE_FP32 (default): every intermediate is fp32; results match the reference's fp32 path within 1e-4.
 * GENIE_STORAGE_BF16: the GATHERED rows — layer-0 features, their source-neighbour mean, the layer-2 source messages and
 *   their mean — are kept as bf16 in HBM and in shared memory (half the bytes on the gather paths that bound the kernels);
 *   all arithmetic stays fp32 (sums, means, 3xTF32 tensor-core stages).  A fast inference mode (BASELINE.json configs[1]),
 *   NOT within the 1e-4 bar (~1e-3 relative).  Needs a CARTESIAN plan with tiling tables and PReLU slopes eligible for the
 *   tensor-core path (layout.h TCS_OK); changes genie_plan_workspace_bytes — set it before sizing the workspace. */
#define GENIE_STORAGE_FP32 0
#define GENIE_STORAGE_BF16 1
GENIE_API int genie_plan_set_storage(genie_plan_t* plan, int32_t storage);

/* ---- grid sharding over the GPUs of one node: halo rows over peer memory ----------------------------------------------
 * The reference runs on one device (module.py:2) and has no counterpart.  On a grid-sharded plan (n_grid_owned > 0: the local
 * grid = owned grid nodes followed by the 1-hop halo of their source-graph in-neighbours) layer 2 of DataAggregation
 * (module.py:95: propagate(A_in_src, ...)) needs the layer-2 message rows v_b of the halo nodes, which their owners compute.
 * Instead of a collective after the kernel, the layer-1 station pass of the OWNER stores those rows straight into the peers'
 * landing buffers while it produces them (peer stores over NVLink / NVSwitch), tile by tile:
 *   genie_peer_alloc   cudaMalloc of a landing buffer + its inter-process handle (GENIE_PEER_HANDLE_BYTES bytes, to be passed
 *                      to the other ranks by any host channel); genie_peer_free releases it
 *   genie_peer_open    maps a peer's buffer from its handle (lazy peer access); genie_peer_close unmaps it
 *   genie_plan_set_halo_export
 *       exp_ptr_dev   int32 [n_grid_owned + 1]  CSR over the owned grid nodes: exports of node g are [exp_ptr[g], exp_ptr[g+1])
 *       exp_peer_dev  int32 [n_exports]         index into peer_base_dev
 *       exp_row_dev   int32 [n_exports]         position of the node in that peer's halo list (row of its landing buffer)
 *       peer_base_dev float* [n_peers]          DEVICE array of the mapped base addresses of the peers' landing buffers
 *       halo_vb_dev   this rank's landing buffer, [n_grid - n_grid_owned][n_sta][16] fp32 (bf16 storage: 16 x 2 bytes): the
 *                     layer-2 source pass reads the rows of halo grid nodes from it
 *   All NULL switches the export off.  The caller orders "every peer has finished its layer-1 pass" before layer 2 (any
 *   stream-ordered barrier, e.g. a one-element all-reduce) and "every peer has finished its layer-2 source pass" before the
 *   next window's layer 1 (the all-gather of the read-in rows does it). */
#define GENIE_PEER_HANDLE_BYTES 64
GENIE_API int genie_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out);
GENIE_API int genie_peer_open(const unsigned char* handle, void** ptr_out);
GENIE_API int genie_peer_close(void* ptr);
GENIE_API int genie_peer_free(void* ptr);
GENIE_API int genie_plan_set_halo_export(genie_plan_t* plan, const int32_t* exp_ptr_dev, const int32_t* exp_peer_dev,
                                         const int32_t* exp_row_dev, float* const* peer_base_dev, const float* halo_vb_dev);

/* Number of floats of the packed (kernel-layout) weight buffer. */
GENIE_API size_t genie_frontend_packed_floats(void);
/* Re-lays the reference-layout weights into `packed_dev` (one small kernel; call again whenever weights change). */
GENIE_API int genie_frontend_pack_weights(const genie_frontend_weights_t* w, float* packed_dev, void* stream);

/* ---- a1: pick window -> Slice / Mask ---------------------------------------------------------------------------------
 * Replaces process_utils.extract_input_from_data (process_utils.py:460-629; trv_times given — the host mirror builds the
 * table from `trv_pairwise` when the caller passes none, :594-596).
 * The fp64 quantities that the reference derives on the host with numpy (`abs_time_ref[0]`, the arange step
 * `(start+dt)-start`, `len(abs_time_ref)`, `ceil(3*sigma/dt)`) are computed by the caller with the same expressions and
 * passed in, so the integer pick->bin and node->bin maps are bit-identical to numpy's.
 */
typedef struct genie_input_params {
    double t0;              /* window origin time (s) */
    double max_t;           /* maximum moveout (s) */
    double kernel_sig_t;    /* sigma (s) */
    double dt;              /* series sampling (s) */
    double ref0;            /* abs_time_ref[0] = t0 - 3*sigma            (process_utils.py:500-502) */
    double ref_step;        /* abs_time_ref[1] - abs_time_ref[0] as numpy.arange produces it */
    int32_t n_ts;           /* len(abs_time_ref) */
    int32_t n_extra;        /* ceil(3*sigma/dt)                           (process_utils.py:520) */
    int32_t n_locs;         /* number of stations of the absolute station table */
    int32_t n_sta_use;      /* number of used stations (len(ind_use)) */
    int32_t use_sign_input; /* process_utils.py:610-614: features times sign(-diff) of the series they were read from */
    int32_t reserved_;
} genie_input_params_t;

/* picks_dev      fp64 [n_picks,5] (time, absolute station, amp, prob, phase 0/1), any order.
 * sta_perm_dev   int32 [n_locs]: absolute station -> index into ind_use, or -1 (process_utils.py:485-486).
 * ind_use_dev    int32 [n_sta_use]: used station -> absolute station.
 * trv_times_dev  fp32 [n_grid, n_locs, 2] travel times (s).
 * node_sta_dev / node_grid_dev: int32 [n_prod] (A_src_in_sta rows 0/1) or both NULL in CARTESIAN mode.
 * series_dev     fp32 scratch, 2 * n_sta_use * n_ts floats (overwritten; internal layout [station][bin][phase]).
 * slice_out_dev, mask_out_dev: fp32 [n_prod,4].  time_bin_out_dev: optional int64 [n_prod,2] (the integer index map).
 */
GENIE_API int genie_input_scatter_fwd(const genie_plan_t* plan, const genie_input_params_t* prm, const double* picks_dev,
                            int64_t n_picks, const int32_t* sta_perm_dev, const int32_t* ind_use_dev,
                            const float* trv_times_dev, const int32_t* node_sta_dev, const int32_t* node_grid_dev,
                            float* series_dev, float* slice_out_dev, float* mask_out_dev, int64_t* time_bin_out_dev,
                            void* stream);

/* ---- read-out heads (the step right after the front end) ---------------------------------------------
 * y = TemporalAttention(SpatialDirect(x_spatial), t_query)                         (module.py:251-260, 299-331, 1015-1016)
 * x = TemporalAttention(SpatialAttention(x_spatial, x_query, x_context), t_query)  (module.py:262-297, 1018-1020)
 *   heads_packed_dev  fp32 [genie_heads_packed_floats()]: every nn.Linear of the three modules K-major [n_in][ld] at the
 *                     offsets genie_heads_layout reports (order: SpatialDirect W, b, slope; TemporalAttention f_context_1 W, b,
 *                     f_values_1 W, b, f_values_2 W, b, proj_1 W, b, proj_2 W, b, slopes {1,2,4,5}; SpatialAttention f_queries
 *                     W, b, f_context W, b, f_values W, b, proj W, b, slopes {1,2}); ld = 32 / 76 for 30 / 75 outputs.
 *   fold_dev          fp32 [n_t*5*32 + n_t*5]: the t_query branch folded into f_context_2 — row (t*5+h) of A (32 floats, 30
 *                     used) = q[t,h,:] . f_context_2.weight[h*15:(h+1)*15, :] / sqrt(15), then a0[t*5+h] = q[t,h,:] .
 *                     f_context_2.bias[h*15:(h+1)*15] / sqrt(15), q = temporal_query_2(activate3(temporal_query_1(t/scale_t))).
 *   nbr_dev           int64 [n_query][k_nbr] context node of every query edge, nearest first (knn(...).flip(0), module.py:282).
 *   y_out_dev [n_grid][n_t], x_out_dev [n_query][n_t].  n_t <= 25 per call (GENIE_ERR_UNSUPPORTED beyond: split the query
 *                     times into blocks — the fold table is per query time — as genie_b200/ops.py heads_fwd does).
 *   proj_out_dev / proj_dev  optional fp32 [n_grid][GENIE_HEADS_PROJ_LD] (NULL = off): f_context and f_values of SpatialAttention
 *                     (module.py:288-290) are linear in [x_j | edge attr], and their x_j parts do not depend on the query — the
 *                     grid kernel computes them once per context node (columns 0-74 and 80-154) and the query kernel adds the
 *                     edge-attribute part per (query, neighbour) pair, instead of two 33 -> 75 layers per pair.
 */
#define GENIE_HEADS_PROJ_LD 160
GENIE_API size_t genie_heads_packed_floats(void);
GENIE_API int genie_heads_layout(int32_t* offsets_out, int n);
GENIE_API int genie_heads_grid_fwd(const float* heads_packed_dev, const float* fold_dev, int n_t, const float* x_spatial_dev,
                                   int ld_x, int n_grid, float* y_out_dev, float* proj_out_dev, void* stream);
GENIE_API int genie_heads_query_fwd(const float* heads_packed_dev, const float* fold_dev, int n_t, const float* x_spatial_dev,
                                    int ld_x, const float* x_context_dev, const float* x_query_dev, const int64_t* nbr_dev,
                                    int k_nbr, int n_query, float scale_rel, float* x_out_dev, const float* proj_dev,
                                    void* stream);

/* ---- product-graph message passing of the training path (BASELINE.json configs[2]) -----------------------------------------
 * Replaces `MessagePassing.propagate(A_in_sta / A_in_src, x=...)` with aggr='mean' (module.py:90-95, 394-400) AND its
 * gradient, i.e. the reference's index_select + scatter_add_ pair in both directions:
 *     out[i, 0:n_ch] = sum_{e in row(i)} val[e] * x[nbr(i, e), 0:n_ch]
 * with a small CSR matrix (rowptr int64, col int32, val fp32, all device pointers) applied as a Kronecker product:
 *   mode 0: node i = g*n_sta + s, row(i) = row s of an [n_sta x n_sta] matrix, nbr = g*n_sta + col   (station edges, :720)
 *   mode 1: node i = g*n_sta + s, row(i) = row g of an [n_grid x n_grid] matrix, nbr = col*n_sta + s (source edges, :721)
 *   mode 2: explicit [n_prod x n_prod] matrix (sub-graph mode).
 * Forward: CSR by target, val = 1/in-degree(target).  Backward of the same op: CSR by source, val = 1/in-degree(edge target).
 * Both are gathers: no atomics, bit-reproducible.  x_dev [n_prod][ld_x], out_dev [n_prod][ld_out]. */
GENIE_API int genie_kron_spmm_fwd(int mode, int n_sta, int n_grid, int64_t n_prod, const int64_t* rowptr_dev,
                                  const int32_t* col_dev, const float* val_dev, const float* x_dev, int ld_x, int n_ch,
                                  float* out_dev, int ld_out, void* stream);

/* ---- per-product-node dense layers of the training path (BASELINE.json configs[2]) -----------------------------------------
 * Replace `activate(Linear(torch.cat((x_0, x_1, ...), dim=1)))` over product-node-sized tensors — every layer of
 * DataAggregation (module.py:87-96), DataAggregationAssociationPhase (:387-403), BipartiteGraphOperator.fc1 (:227) and
 * BipartiteGraphReadOutOperator (:349-351) — and its gradient (train_GENIE_model.py:1786-1861: loss.backward()).
 *   forward   y = PReLU_a(W [x_0 | x_1 | ...] + b)        the concatenation is never materialised; slope NULL = no activation
 *   backward  g = gy * PReLU_a'(y);  gx_p = g W[:, columns of part p] (gx_dev[p] NULL = not wanted);
 *             per-CTA partial sums of  gW = g^T [x_0 | x_1 | ...],  gb = sum g,  ga = sum_{z<0} gy * y / a
 *             neg_mask_dev: uint32 [n_rows], bit o = the pre-activation z of output o is negative — written by the forward
 *             call (NULL = not kept), read by the backward call (needed whenever slope is given: a slope may be negative,
 *             so the sign of y does not tell)
 *             partial_dev: fp32 [genie_node_mlp_partial_rows()][n_out * n_in + n_out + 1] (gW row-major, then gb, then ga); the
 *             caller sums over the first dimension (fixed order: bit-reproducible, no atomics).
 * Limits: 1..4 parts, sum of widths <= GENIE_MLP_MAX_IN, n_out <= GENIE_MLP_MAX_OUT, slope != 0 when given. */
#define GENIE_MLP_MAX_IN 104
#define GENIE_MLP_MAX_OUT 32
typedef struct genie_mlp_desc {
    int64_t n_rows;
    int32_t n_parts;
    int32_t n_out;
    int32_t width[4];       /* columns of every part */
    int32_t ld[4];          /* row stride (floats) of every part */
    const float* x[4];      /* device pointers */
    const float* weight;    /* [n_out][n_in] row-major (nn.Linear.weight), n_in = sum(width) */
    const float* bias;      /* [n_out] or NULL */
    const float* slope;     /* 1-element nn.PReLU weight or NULL */
} genie_mlp_desc_t;
GENIE_API int genie_node_mlp_partial_rows(void);
GENIE_API int genie_node_mlp_fwd(const genie_mlp_desc_t* desc, float* y_dev, int32_t ld_y, uint32_t* neg_mask_dev, void* stream);
GENIE_API int genie_node_mlp_bwd(const genie_mlp_desc_t* desc, const float* y_dev, int32_t ld_y, const uint32_t* neg_mask_dev,
                                 const float* gy_dev, int32_t ld_gy, float* const* gx_dev, const int32_t* ld_gx,
                                 float* partial_dev, void* stream);

/* ---- output stacking of the streaming loop --------------------------------------------------------
 * process_continuous_days.py:797-805: Out_2[:, ip_need[t]] += x[:, t, 0] / n_overlap / n_scale_x_grid for the first n_use
 * (all, or all but the last when step_size == 'half') query times of one window, on the device.
 *   x_dev [n_query][n_t] (the query prediction of forward_fixed_source); col_dev int32 [n_use] distinct columns of out_dev
 *   [n_query][ld_out] (the nearest solution-grid step of every query time, the caller's cKDTree look-up at :793);
 *   scale = 1 / (n_overlap * n_scale_x_grid).  Launch windows in stream order: columns of consecutive windows overlap. */
GENIE_API int genie_stack_output_fwd(const float* x_dev, int n_query, int n_t, int n_use, const int32_t* col_dev, float scale,
                                     float* out_dev, int64_t ld_out, void* stream);

/* ---- device kNN ---------------------------------------------------------------------------------
 * Replaces torch_cluster.knn(x, y, k) at process_utils.py:718-719 (station / source graphs) and module.py:282 (query edges):
 * idx_out_dev int64 [n_y][k] = the k rows of x_dev [n_x][3] nearest to every row of y_dev [n_y][3], nearest first (the
 * reference's row 1 of knn(x, y, k), whose row 0 is repeat(arange(n_y), k)).  Coordinates fp32 exactly as the reference passes
 * them (kilometres); distances are evaluated in fp64, ties keep the lower index.  1 <= k <= min(32, n_x). */
GENIE_API int genie_knn_fwd(const float* x_dev, int n_x, const float* y_dev, int n_y, int k, int64_t* idx_out_dev, void* stream);

/* ---- association branch (forward / forward_fixed, module.py:983-991) -------------------------------
 * The product-node-sized part of the association branch:
 *   mask_out = max_t y[g,t] > mask_thresh                                                              (module.py:983)
 *   s0 = BipartiteGraphReadOutOperator(SpatialDirect(x_spatial), A_Lg_in_src, mask_out)                (module.py:333-352)
 *   s  = DataAggregationAssociationPhase(s0, x_latent, mask_out[g(i)], Mask, A_in_sta, A_in_src)       (module.py:356-403)
 *   arrival[a] = [LocalSliceLgCollapseP(...)[a] | LocalSliceLgCollapseS(...)[a]], a null row appended  (module.py:604-653, 709-711)
 * A_Lg_in_src must be the flipped read-in edge list (edge e: grid node g(e) -> product node e, process_continuous_days.py:632)
 * and carry the same [P,3] edge features as A_src_in_edges; the graphs are the plan's.
 *   assoc_packed_dev   fp32 [genie_assoc_packed_floats()], every nn.Linear K-major [n_in][ld] at the offsets genie_assoc_layout
 *                      reports (order: SpatialDirect W, b; read-out fc1[:, :30], fc1.bias, fc1[:, 30:33], fc2 W, b; init_trns W,
 *                      b; l1_t1_1 W, b; l1_t2_1 W, b; l1_t1_2, l1_t2_2 W, their biases; l2_t1_1, l2_t2_1 W, their biases;
 *                      l2_t1_2[:, 60:90], l2_t2_2[:, 60:90]; l2_t1_2[:, 0:60 | 90:95], l2_t2_2[...], their biases; collapse P
 *                      fc1 W, b, fc2 W, b; collapse S the same; 16 slopes {SpatialDirect, read-out 1, 2, association activate,
 *                      11, 12, 1, 21, 22, 2, collapse P 1, 2, collapse S 1, 2}); ld = 32 / 16 for 30 / 15 outputs.
 *   x_spatial_dev [n_grid][ld_x]; y_dev [n_grid][n_t] (the grid prediction of the heads); x_latent_dev [P][30] (DataAggregation
 *   output, genie_frontend_fwd's x_latent_out_dev); s0_out_dev [P][15] or NULL; mask_out_dev [n_grid] or NULL;
 *   *s_rows_out (optional) receives the address, inside the workspace, of s as [P][32] rows = [s[0:15] 0 | s[15:30] 0].
 *   genie_assoc_collapse_fwd: edges_{p,s}_dev int64 [n_sta * l_dt * k_infer] product-node pointers (A_edges_p / A_edges_s);
 *   tlatent_dev [P][2]; tpick fp32, ipick int64, phase fp32 [n_arv]; dt0 = dt_partition[0], dt_step = dt_partition[1] -
 *   dt_partition[0] (fp32 difference, as torch forms it); arrival_out_dev [n_arv + 1][30].
 */
GENIE_API size_t genie_assoc_packed_floats(void);
GENIE_API int genie_assoc_layout(int32_t* offsets_out, int n);
GENIE_API size_t genie_assoc_workspace_bytes(const genie_plan_t* plan);
/* Model variants of the association branch, by the same linearity arguments as genie_plan_set_init_terms /
 * genie_plan_set_edge_terms (all tables the caller's, window-independent; NULL = off):
 *   `use_absolute_pos: True` (module.py:987-988): init_sta_dev / init_src_dev [S][32] / [G][32] (CARTESIAN) or [P][32] / NULL
 *       = DataAggregationAssociationPhase.init_trns.weight[:, 15:18] . locs / (3 scale_rel) and [:, 18:21] . x_temp / (3 scale_rel);
 *       the packed init_trns then has its six position columns removed ([30][50]).
 *   `use_updated_model_definition: True` (DataAggregationAssociationPhaseEdges, module.py:406-481): edge_sta_dev / edge_src_dev
 *       [S or P][GENIE_EDGE_TERM_LD] / [G or P][GENIE_EDGE_TERM_LD] with [0,30) = l1_t*_2.weight[:, 60:64] . mean pos_rel and
 *       [32,47) = l2_t*_2.weight[:, 90:94] . mean pos_rel; the packed l1_t*_2 / l2_t*_2 then lack those four columns. */
GENIE_API int genie_assoc_set_terms(genie_plan_t* plan, const float* init_sta_dev, const float* init_src_dev,
                                    const float* edge_sta_dev, const float* edge_src_dev);
GENIE_API int genie_assoc_product_fwd(const genie_plan_t* plan, const float* assoc_packed_dev, const float* x_spatial_dev,
                                      int ld_x, const float* y_dev, int n_t, float mask_thresh, const float* edge_attr_dev,
                                      const float* x_latent_dev, const float* mask_dev, void* assoc_workspace_dev,
                                      float* s0_out_dev, float* mask_out_dev, float** s_rows_out, void* stream);
GENIE_API int genie_assoc_collapse_fwd(const float* assoc_packed_dev, const float* s_rows_dev, int64_t n_prod,
                                       const int64_t* edges_p_dev, const int64_t* edges_s_dev, int64_t n_edges,
                                       const float* tlatent_dev, const float* tpick_dev, const int64_t* ipick_dev,
                                       const float* phase_dev, int n_arv, int n_sta, int l_dt, int k_infer, float dt0,
                                       float dt_step, float eps, float* arrival_out_dev, void* stream);

/* ---- a1': nearest-pick input features -----------------------------------------------------------------------------------
 * Replaces the device-sized part of process_utils.extract_inputs_from_data_fixed_grids_with_phase_type
 * (process_utils.py:194-268; the input features used when `use_updated_input: False` and in training): per sample b, product
 * node (g, s) and phase, the query time  (trv[g, sta, phase] + b * offset_per_batch) + sta * offset_per_station  (fp64, in
 * that order) is located with searchsorted(left) on a sorted time axis; the distance to the nearer of its two neighbours
 * (indices clipped to the axis, :199-203) goes through exp((-0.5 * d^2) / sigma^2) in fp64.  Channels 0, 1: the axis of all
 * selected picks at the P / S query; channels 2, 3: the axes of the P picks / S picks (0 when the axis is empty).
 * The caller builds the axes on the host exactly as the reference does (:137-189, a few thousand picks).
 *   times_all_dev / times_p_dev / times_s_dev   fp64, sorted ascending, n_all / n_p / n_s entries (NULL allowed when 0).
 *   ind_use_dev    int32 [n_sta_use] used station -> absolute station (sorted: np.unique(ind_use), :152).
 *   trv_times_dev  fp32 [G, n_locs, 2].
 *   slice_out_dev, mask_out_dev   fp32 [n_batch, G * n_sta_use, 4] (grid-major as :268); Mask = value > 0.01 on the fp64 value.
 */
typedef struct genie_nearest_params {
    double offset_per_batch;     /* 1.5 * max_t                                  (process_utils.py:177) */
    double offset_per_station;   /* 1.5 * n_batch * offset_per_batch             (process_utils.py:178) */
    double kernel_sig_t;
    int64_t n_all, n_p, n_s;
    int32_t n_batch;
    int32_t n_grid;
    int32_t n_locs;
    int32_t n_sta_use;
} genie_nearest_params_t;
GENIE_API int genie_input_nearest_fwd(const genie_nearest_params_t* prm, const double* times_all_dev, const double* times_p_dev,
                                      const double* times_s_dev, const int32_t* ind_use_dev, const float* trv_times_dev,
                                      float* slice_out_dev, float* mask_out_dev, void* stream);

/* ---- a2: DataAggregation.forward — module.py:85-98 -----------------------------------------------------------------
 * slice_dev, mask_dev fp32 [P,4] -> x_latent_out_dev fp32 [P,30]. */
GENIE_API int genie_data_aggregation_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev,
                               const float* mask_dev, float* x_latent_out_dev, void* workspace_dev, void* stream);

/* ---- a3: BipartiteGraphOperator.forward — module.py:224-229 --------------------------------------------------------
 * x_latent_dev [P,30], edge_attr_dev [P,3] (A_src_in_edges.x), mask_dev [P,4] -> out_dev [G,15]. */
GENIE_API int genie_bipartite_readin_fwd(const genie_plan_t* plan, const float* packed_dev, const float* x_latent_dev,
                               const float* edge_attr_dev, const float* mask_dev, float* out_dev,
                               void* workspace_dev, void* stream);

/* ---- a4: SpatialAggregation.forward — module.py:243-249 ------------------------------------------------------------
 * layer = 0,1,2 (SpatialAggregation1..3): x_dev [G,C] (C = 15,30,30), pos_dev [G,3] Cartesian metres -> out_dev [G,30]. */
GENIE_API int genie_spatial_aggregation_fwd(const genie_plan_t* plan, const float* packed_dev, int32_t layer, const float* x_dev,
                                  const float* pos_dev, float scale_rel, float* out_dev, void* workspace_dev,
                                  void* stream);

/* ---- fused a2+a3+a4x3: the front end of forward_fixed_source — module.py:1010-1014 ---------------------------------
 * x_latent_out_dev [P,30] and readin_out_dev [G,15] are optional (NULL = not materialised);
 * x_spatial_out_dev [G,30] is SpatialAggregation3's output. */
GENIE_API int genie_frontend_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev,
                       const float* mask_dev, const float* edge_attr_dev, const float* pos_dev, float scale_rel,
                       float* x_latent_out_dev, float* readin_out_dev, float* x_spatial_out_dev,
                       void* workspace_dev, void* stream);

/* ---- a1 fused into the front end: one window of the streaming loop — process_continuous_days.py:776-797 -------------
 * extract_input_from_data + forward_fixed_source's front end for CARTESIAN plans whose pick table, travel times and graphs
 * are resident on the device.  Slice / Mask never reach HBM: the per-station series is built as in genie_input_scatter_fwd,
 * and the layer-0 kernel computes every node's Slice / Mask row in registers (same fp64 bin arithmetic) right before
 * init_trns.  Everything that changes from window to window sits in ONE device-resident block, so that the whole window
 * can be captured in a CUDA graph once and replayed (the caller refreshes the block, e.g. by a captured copy from pinned
 * host memory):
 *   wp_dev            DEVICE pointer to the window's parameters and the row range [pick_lo, pick_hi) of its picks inside
 *                     picks_dev (the reference's selection P[:,0] in (t0 - 2 sigma, t0 + max_t + 2 sigma), :476, is applied
 *                     per pick again, so any superset range is fine)
 *   max_window_picks  launch bound of the series kernel (a frozen graph cannot resize its grids): pick_hi - pick_lo must not
 *                     exceed it — rows beyond it are NOT processed; n_extra = the host's copy of prm.n_extra (a constant
 *                     of sigma and dt), which sizes that grid as well
 *   series_dev        fp32 scratch for 2 * n_sta_use * n_ts_max floats; n_ts_max >= every window's prm.n_ts
 *   slice_out_dev / mask_out_dev   optional [P,4] copies of the inputs (NULL = not materialised)
 * Other arguments as genie_frontend_fwd.  Plans without tiling tables and EXPLICIT plans: GENIE_ERR_UNSUPPORTED
 * (use genie_input_scatter_fwd + genie_frontend_fwd). */
typedef struct genie_window_params {
    genie_input_params_t prm;
    int64_t pick_lo, pick_hi;
} genie_window_params_t;

GENIE_API int genie_window_fwd(const genie_plan_t* plan, const float* packed_dev, const genie_window_params_t* wp_dev,
                               int64_t max_window_picks, int32_t n_extra, const double* picks_dev, const int32_t* sta_perm_dev,
                               const int32_t* ind_use_dev, const float* trv_times_dev, float* series_dev, int32_t n_ts_max,
                               const float* edge_attr_dev, const float* pos_dev, float scale_rel, float* slice_out_dev,
                               float* mask_out_dev, float* x_latent_out_dev, float* readin_out_dev,
                               float* x_spatial_out_dev, void* workspace_dev, void* stream);

/* ---- the same front end in two halves, for grid-sharded plans (genie_b200/sharded.py) ----------------------------------
 * genie_da_layer1_fwd runs DataAggregation up to the layer-2 messages (module.py:88-93) and leaves them in the workspace;
 * genie_workspace_region exposes the message rows v_b (`which` = GENIE_WS_VB: fp32 [n_prod,16], one row per product node)
 * so that the caller can fill the rows of its halo grid nodes from their owners (NCCL all-to-all);
 * genie_da_layer2_readin_fwd then finishes DataAggregation (module.py:94-98) and applies Bipartite_ReadIn
 * (module.py:224-229) for the owned grid nodes: readin_out_dev fp32 [n_grid_owned,15], x_latent_out_dev optional. */
enum { GENIE_WS_VB = 0 };
GENIE_API int genie_da_layer1_fwd(const genie_plan_t* plan, const float* packed_dev, const float* slice_dev,
                                  const float* mask_dev, void* workspace_dev, void* stream);
GENIE_API int genie_workspace_region(const genie_plan_t* plan, void* workspace_dev, int32_t which, void** ptr_out,
                                     size_t* bytes_out);
GENIE_API int genie_da_layer2_readin_fwd(const genie_plan_t* plan, const float* packed_dev, const float* mask_dev,
                                         const float* edge_attr_dev, float* x_latent_out_dev, float* readin_out_dev,
                                         void* workspace_dev, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------------------------- */
GENIE_API const char* genie_last_error(void);
GENIE_API int genie_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
GENIE_API int64_t genie_launch_count(void);

/* Optional per-kernel device timing (bench.py's roofline leg; the reference has no counterpart: it only prints
 * time.time() deltas behind `verbose`, process_utils.py:469-470, 639-640).  While enabled, every kernel launch is
 * bracketed by cudaEvents recorded on the launching stream (skipped while that stream is being captured into a CUDA
 * graph).  genie_timing_collect waits for the recorded events, adds their durations to per-kernel totals and copies the
 * totals into total_ms[genie_timing_kernel_count()] / launches[...] (either may be NULL); reset != 0 clears them. */
GENIE_API int genie_timing_enable(int on);
GENIE_API int genie_timing_kernel_count(void);
GENIE_API const char* genie_timing_kernel_name(int k);
GENIE_API int genie_timing_collect(double* total_ms, int64_t* launches, int reset);

/* Development aid: while trace_dev != NULL, CTA 0 of the layer-1 station-pass kernel stamps clock64() at its pipeline
 * hand-off points into trace_dev[(tile * 2 + pipeline) * 24 + slot] for `tiles` tiles per pipeline (slots: scripts/trace_s1.py). */
GENIE_API int genie_debug_trace(int64_t* trace_dev, int tiles);

#ifdef __cplusplus
}
#endif
#endif /* GENIE_B200_H_ */
