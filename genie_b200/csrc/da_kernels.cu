// DataAggregation (module.py:52-98) and BipartiteGraphOperator (module.py:214-229) forward kernels — generic path.
//
// The reference evaluates, per product node i (Slice/Mask rows x0 = [slice ‖ mask]):
//   tr0 = PReLU_a (W0 x0)
//   tr  = PReLU_1([W11 [tr0 ‖ mean_sta PReLU_11(tr0_j) ‖ mask] ‖ W12 [tr0 ‖ mean_src PReLU_12(tr0_j) ‖ mask]])
//   out = PReLU_2([W21b [tr ‖ mean_sta a_j ‖ mask] ‖ W22b [tr ‖ mean_src b_j ‖ mask]]),  a = PReLU_21(W21a tr), b = ...
// Three kernels, separated by the two global dependencies (neighbours' tr0, neighbours' a/b):
//   da_init_kernel     x0 -> tr0                                                     (row stride 32 floats)
//   da_layer1_kernel   gather tr0 -> tr -> {ca, cb, va, vb}, where the 15-wide halves of the last linear layer are
//                      split by linearity:  W21b[tr ‖ mean a ‖ mask] = (W_tr tr + W_m mask + b) + mean_j (W_agg a_j)
//                      = ca + mean_sta(va_j), so layer 2 gathers 15-wide rows instead of 30-wide ones.
//   da_layer2_readin_kernel   gather va/vb -> x_latent -> (optional store) -> fc1, mask, sum over the node's grid
//                      node (BipartiteGraphOperator) -> xg accumulator; readin_finalize_kernel applies fc2.
// Warps gather (one product node per warp, lanes = channels: every neighbour row is one coalesced 128-byte / 64-byte
// request); the dense per-node MLPs run thread-per-node on a transposed shared-memory feature tile with the weights
// broadcast from shared memory.
#include "common.cuh"
#include "gather.cuh"
#include "bf16.cuh"
#include "input.cuh"

using namespace gl;

namespace {

constexpr int TM = 128;        // product nodes per tile
constexpr int LDF = TM + 1;    // feature-tile row stride (floats): conflict-free for both access patterns

// --------------------------------------------------------------------------------------------------------------------
// K1: tr0 = PReLU(init_trns([slice ‖ mask]))
// --------------------------------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 256;
// init_trns weights + bias in the constant bank (FFMA reads them as uniform operands: no shared-memory broadcasts);
// refreshed from the packed weights by a stream-ordered device-to-device copy before every launch.
__constant__ float c_init_slots[GENIE_CSLOTS][8 * LD + LD];     // one slot per plan (genie_plan::cslot)

// FUSED: a1 is folded in (genie_window_fwd) — the thread computes its node's Slice / Mask row from the per-station series and
// the travel-time table (input.cuh, the arithmetic of input_gather_kernel) instead of loading it; the rows reach HBM only
// when the caller asks for copies.  The four mask values ride in padding channel 30 of the stored row (pack_mask): the
// station-pass kernels take the mask from the rows they stage anyway.
struct InitInputs {
    const float* slice;          // !FUSED: [P,4]
    const float* mask;           // !FUSED: [P,4]
    WindowParamSrc ws;           // FUSED
    const int32_t* ind_use;      // FUSED: used station -> absolute station
    const float* trv;            // FUSED: [G, n_locs, 2]
    const float* series;         // FUSED: [S][n_ts][2]
    float* slice_out;            // FUSED, optional
    float* mask_out;             // FUSED, optional
};

// BF16: the row is stored as 32 bf16 (64 bytes; genie_plan_set_storage) instead of 32 floats.
template <int CSLOT, bool FUSED, bool BF16>
__global__ void __launch_bounds__(K1_THREADS) da_init_kernel(const float* __restrict__ packed, const InitInputs in,
                                                             float* __restrict__ tr0, int64_t P, int tc_plan,
                                                             const float* __restrict__ init_sta,
                                                             const float* __restrict__ init_src, int S) {
    __shared__ __align__(16) float sOut[K1_THREADS * LD_TR0];
    const float* c_init = c_init_slots[CSLOT];       // compile-time slot: the weights stay immediate constant-bank operands
    const float a0 = packed[DA_SLOPES + SL_A0];
    // tensor-core path (da_tc_kernels.cu): store p = PReLU12(tr0) instead of tr0
    const bool post = tc_plan && packed[TC_BASE + TC_SCAL + TCS_OK] != 0.f;
    const float a12 = post ? packed[DA_SLOPES + SL_A12] : 1.f;

    const int64_t i0 = (int64_t)blockIdx.x * K1_THREADS;
    const int n = threadIdx.x;
    const int64_t i = i0 + n;
    f32x2_t acc2[15];
#pragma unroll
    for (int o = 0; o < 15; ++o) acc2[o] = pack2(c_init[8 * LD + 2 * o], c_init[8 * LD + 2 * o + 1]);
    uint32_t g32 = 0, s32 = 0;
    if ((FUSED || init_src != nullptr) && i < P) {           // 32-bit division (P < 2^31 is checked by the launcher)
        g32 = (uint32_t)i / (uint32_t)S;
        s32 = (uint32_t)i - g32 * (uint32_t)S;
    }
    if (init_sta != nullptr && i < P) {
        // use_absolute_pos (module.py:913-914): the six position channels of init_trns, by linearity a per-station plus a
        // per-grid-node term (CARTESIAN) or one per-node term (EXPLICIT) added before the activation
        const float2* ts = reinterpret_cast<const float2*>(init_sta + (init_src != nullptr ? (int64_t)s32 : i) * 32);
#pragma unroll
        for (int o = 0; o < 15; ++o) {
            const float2 v = __ldg(ts + o);
            fadd2(acc2[o], pack2(v.x, v.y));
        }
        if (init_src != nullptr) {
            const float2* tg = reinterpret_cast<const float2*>(init_src + (int64_t)g32 * 32);
#pragma unroll
            for (int o = 0; o < 15; ++o) {
                const float2 v = __ldg(tg + o);
                fadd2(acc2[o], pack2(v.x, v.y));
            }
        }
    }
    float mpack = 0.f;
    if (i < P) {
        float4 sv, mv;
        if (FUSED) {
            int64_t lo, hi;
            const genie_input_params_t prm = load_params(in.ws, lo, hi);
            const int sta_abs = __ldg(in.ind_use + s32);
            const float2 tt = __ldg(reinterpret_cast<const float2*>(in.trv + ((int64_t)g32 * prm.n_locs + sta_abs) * 2));
            long long bp, bs;
            sv = input_slice_row(prm, (int)s32, tt, in.series, bp, bs);
            mv = input_mask_row(sv);
            if (in.slice_out != nullptr) {
                __stcs(reinterpret_cast<float4*>(in.slice_out) + i, sv);
                __stcs(reinterpret_cast<float4*>(in.mask_out) + i, mv);
            }
        } else {
            sv = __ldcs(reinterpret_cast<const float4*>(in.slice) + i);
            mv = __ldg(reinterpret_cast<const float4*>(in.mask) + i);
        }
        mpack = pack_mask(make_float4(mv.x != 0.f ? 1.f : 0.f, mv.y != 0.f ? 1.f : 0.f, mv.z != 0.f ? 1.f : 0.f,
                                      mv.w != 0.f ? 1.f : 0.f));
        const float inp[8] = {sv.x, sv.y, sv.z, sv.w, mv.x, mv.y, mv.z, mv.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 0; o < 15; ++o) ffma2(acc2[o], inp[k], c_init[k * LD + 2 * o], c_init[k * LD + 2 * o + 1]);
        }
    }
    float acc[30];
#pragma unroll
    for (int o = 0; o < 15; ++o) unpack2(acc2[o], acc[2 * o], acc[2 * o + 1]);
    // stage the row in shared memory (16-byte chunks XOR-swizzled by the row index: conflict-free), then write the
    // tile out as one contiguous, fully coalesced block.
    float row[32];
#pragma unroll
    for (int o = 0; o < 30; ++o) row[o] = prelu(prelu(acc[o], a0), a12);
    row[30] = mpack;                                                                                   // channel 30
    row[31] = 0.f;
    if (BF16) {
        uint4* srow = reinterpret_cast<uint4*>(sOut) + n * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v8[8] = {row[8 * c], row[8 * c + 1], row[8 * c + 2], row[8 * c + 3],
                                 row[8 * c + 4], row[8 * c + 5], row[8 * c + 6], row[8 * c + 7]};
            srow[c ^ ((n >> 1) & 3)] = bf16_pack8(v8);
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(tr0) + i0 * 4;          // bf16 rows: 4 chunks of 16 bytes
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int f = it * K1_THREADS + threadIdx.x;
            const int r = f >> 2, pos = f & 3;
            if (i0 + r < P) dst[r * 4 + (pos ^ ((r >> 1) & 3))] = reinterpret_cast<const uint4*>(sOut)[f];
        }
    } else {
        float4* srow = reinterpret_cast<float4*>(sOut + n * LD_TR0);
#pragma unroll
        for (int c = 0; c < 8; ++c) srow[c ^ (n & 7)] = make_float4(row[4 * c], row[4 * c + 1], row[4 * c + 2], row[4 * c + 3]);
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(tr0 + i0 * LD_TR0);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int f = it * K1_THREADS + threadIdx.x;   // float4 index inside the tile
            const int r = f >> 3, pos = f & 7;
            if (i0 + r < P) dst[r * 8 + (pos ^ (r & 7))] = reinterpret_cast<const float4*>(sOut)[f];
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K2: layer 1 (+ the node-local part of layer 2)
// --------------------------------------------------------------------------------------------------------------------
constexpr int K2_THREADS = 256;
constexpr int K2_W_FLOATS = DA_END - DA_W11;
constexpr int K2_F_ROWS = 94;   // 0-29 tr0 | 30-59 mean_sta | 60-89 mean_src | 90-93 mask ; rows 0-59 later hold tr
constexpr size_t K2_SMEM = (size_t)(K2_W_FLOATS + K2_F_ROWS * LDF) * sizeof(float);

__global__ void __launch_bounds__(K2_THREADS, 2)
    da_layer1_kernel(const GraphView gv, const float* __restrict__ packed, const float* __restrict__ tr0,
                     const float* __restrict__ mask, float* __restrict__ zc, float* __restrict__ va,
                     float* __restrict__ vb, int64_t n_tiles, int tc_plan) {
    extern __shared__ __align__(16) float smem[];
    if (tc_plan && packed[TC_BASE + TC_SCAL + TCS_OK] != 0.f) return;   // da_layer1_tc_kernel did the work
    float* sW = smem;                  // packed[DA_W11 .. DA_END)
    float* F = smem + K2_W_FLOATS;     // [94][LDF]
    {
        const float4* src = reinterpret_cast<const float4*>(packed + DA_W11);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < K2_W_FLOATS / 4; i += K2_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const float* sl = sW + (DA_SLOPES - DA_W11);
    const float a11 = sl[SL_A11], a12 = sl[SL_A12], a1 = sl[SL_A1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = threadIdx.x & (TM - 1);
    const int br = threadIdx.x >> 7;   // 0: station-edge branch (l*_t1_*), 1: source-edge branch (l*_t2_*)

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * TM;
        // ---- stage A: gather ------------------------------------------------------------------------------------------
        for (int m = warp; m < TM; m += K2_THREADS / 32) {
            const int64_t i = i0 + m;
            float own = 0.f, m1 = 0.f, m2 = 0.f, mk = 0.f;
            if (i < gv.P) {
                NbrRange rs, rg;
                node_ranges(gv, i, rs, rg);
                own = tr0[i * LD_TR0 + lane];
                m1 = gather_mean32(tr0, rs, gv.sta_col, a11, lane);
                m2 = gather_mean32(tr0, rg, gv.src_col, a12, lane);
                if (lane < 4)     // mask == NULL (a1 fused into layer 0): the four values ride bit-packed in channel 30 of the row
                    mk = mask != nullptr ? mask[i * 4 + lane] : (float)(((int)tr0[i * LD_TR0 + 30] >> lane) & 1);
            }
            if (lane < 30) {
                F[lane * LDF + m] = own;
                F[(30 + lane) * LDF + m] = m1;
                F[(60 + lane) * LDF + m] = m2;
            }
            if (lane < 4) F[(90 + lane) * LDF + m] = mk;
        }
        __syncthreads();
        // edge-feature model (genie_plan_set_edge_terms): per-node additive terms of this thread's branch, or NULL
        const float* et = nullptr;
        if (gv.edge_sta != nullptr && i0 + n < gv.P) {
            int64_t idx = i0 + n;
            if (gv.mode == GENIE_GRAPH_CARTESIAN) {
                const int64_t g = idx / gv.S;
                idx = br ? g : idx - g * gv.S;
            }
            et = (br ? gv.edge_src : gv.edge_sta) + idx * GENIE_EDGE_TERM_LD;
        }
        // ---- stage B: tr = PReLU1([l1_t1_2(..) ‖ l1_t2_2(..)]) ------------------------------------------------------
        {
            const float* W = sW + (br ? (DA_W12 - DA_W11) : 0);
            const float* B = sW + ((br ? DA_B12 : DA_B11) - DA_W11);
            float acc[30];
#pragma unroll
            for (int o = 0; o < 30; ++o) acc[o] = B[o];
            if (et != nullptr) {
#pragma unroll
                for (int o = 0; o < 30; ++o) acc[o] += __ldg(et + o);
            }
#pragma unroll 2
            for (int k = 0; k < 30; ++k) fma_row30(acc, F[k * LDF + n], W + k * LD);
            const float* Fm = F + (30 + 30 * br) * LDF;
#pragma unroll 2
            for (int k = 0; k < 30; ++k) fma_row30(acc, Fm[k * LDF + n], W + (30 + k) * LD);
#pragma unroll
            for (int k = 0; k < 4; ++k) fma_row30(acc, F[(90 + k) * LDF + n], W + (60 + k) * LD);
            __syncthreads();   // every thread has finished reading rows 0-89
#pragma unroll
            for (int o = 0; o < 30; ++o) F[(br * 30 + o) * LDF + n] = prelu(acc[o], a1);
        }
        __syncthreads();
        // ---- stage C: h = PReLU(l2_t*_1 tr);  v = W_agg h;  c = W_tr tr + W_m mask + b ---------------------------------
        {
            const float* Wa = sW + ((br ? DA_W22A : DA_W21A) - DA_W11);
            const float* Ba = sW + ((br ? DA_B22A : DA_B21A) - DA_W11);
            const float ah = br ? sl[SL_A22] : sl[SL_A21];
            float h[30];
#pragma unroll
            for (int o = 0; o < 30; ++o) h[o] = Ba[o];
#pragma unroll 2
            for (int k = 0; k < 60; ++k) fma_row30(h, F[k * LDF + n], Wa + k * LD);
            float v[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) v[o] = 0.f;
            const float* Wv = sW + ((br ? DA_WVB : DA_WVA) - DA_W11);
#pragma unroll
            for (int k = 0; k < 30; ++k) fma_row16(v, prelu(h[k], ah), Wv + k * LD16);
            const float* Wc = sW + ((br ? DA_WCB : DA_WCA) - DA_W11);
            const float* Bc = sW + ((br ? DA_BCB : DA_BCA) - DA_W11);
            float c[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) c[o] = Bc[o];
            if (et != nullptr) {
#pragma unroll
                for (int o = 0; o < 15; ++o) c[o] += __ldg(et + 32 + o);
            }
#pragma unroll 2
            for (int k = 0; k < 60; ++k) fma_row16(c, F[k * LDF + n], Wc + k * LD16);
#pragma unroll
            for (int k = 0; k < 4; ++k) fma_row16(c, F[(90 + k) * LDF + n], Wc + (60 + k) * LD16);
            // padding channel 15 of zc carries max_c(mask) for the read-in of the split layer-2 kernel (da_s2_kernel.cu)
            if (br == 0)
                c[15] = fmaxf(fmaxf(F[90 * LDF + n], F[91 * LDF + n]), fmaxf(F[92 * LDF + n], F[93 * LDF + n]));
            const int64_t i = i0 + n;
            if (i < gv.P) {
                float4* zp = reinterpret_cast<float4*>(zc + i * LD_ZC + br * 16);
                float4* vp = reinterpret_cast<float4*>((br ? vb : va) + i * LD_V);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    zp[q] = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
                    vp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
        __syncthreads();   // F is rewritten by the next tile's gather
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K3: layer 2 aggregation -> x_latent -> BipartiteGraphOperator.fc1 / mask / sum over stations
// --------------------------------------------------------------------------------------------------------------------
constexpr int K3_THREADS = 256;
constexpr int K3_W_FLOATS = RI_END - RI_WFC1;

template <int MODE>
__global__ void __launch_bounds__(K3_THREADS)
    da_layer2_readin_kernel(const GraphView gv, const float* __restrict__ packed, const float* __restrict__ zc,
                            const float* __restrict__ va, const float* __restrict__ vb,
                            const float* __restrict__ latent_in, float* __restrict__ latent_out,
                            const float* __restrict__ edge_attr, const float* __restrict__ mask,
                            float* __restrict__ xg, int64_t n_tiles) {
    __shared__ __align__(16) float sW[K3_W_FLOATS];
    __shared__ float L[30 * LDF];     // x_latent tile, transposed
    __shared__ float Hs[30 * LDF];    // masked fc1 output, transposed
    __shared__ int gid[TM];
    for (int i = threadIdx.x; i < K3_W_FLOATS; i += K3_THREADS) sW[i] = packed[RI_WFC1 + i];
    const float a2 = packed[DA_SLOPES + SL_A2];
    __syncthreads();
    const float ri_a1 = sW[RI_SLOPES - RI_WFC1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * TM;
        if (MODE & L2_GATHER) {
            const int half = lane >> 4, l = lane & 15;
            for (int m = warp; m < TM; m += K3_THREADS / 32) {
                const int64_t i = i0 + m;
                float x = 0.f;
                if (i < gv.P) {
                    NbrRange rs, rg;
                    node_ranges(gv, i, rs, rg);
                    const float own = zc[i * LD_ZC + lane];
                    const float mean = gather_mean16x2(va, vb, rs, rg, gv.sta_col, gv.src_col, lane);
                    x = prelu(own + mean, a2);
                }
                if (l < 15) L[(half * 15 + l) * LDF + m] = x;
            }
        } else {
            for (int idx = threadIdx.x; idx < TM * 30; idx += K3_THREADS) {
                const int m = idx / 30, ch = idx - m * 30;
                L[ch * LDF + m] = (i0 + m) < gv.P ? latent_in[i0 * 30 + idx] : 0.f;
            }
        }
        __syncthreads();
        if (MODE & L2_STORE_LATENT) {
            for (int idx = threadIdx.x; idx < TM * 30; idx += K3_THREADS) {
                const int m = idx / 30, ch = idx - m * 30;
                if ((i0 + m) < gv.P) latent_out[i0 * 30 + idx] = L[ch * LDF + m];
            }
        }
        if (MODE & L2_READIN) {
            // fc1 split over two threads per node: outputs [0,16) and [16,32) (30, 31 are zero padding)
            const int n = threadIdx.x & (TM - 1);
            const int br = threadIdx.x >> 7;
            const int64_t i = i0 + n;
            float acc[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[o] = sW[(RI_BFC1 - RI_WFC1) + br * 16 + o];
#pragma unroll 2
            for (int k = 0; k < 30; ++k) fma_row16(acc, L[k * LDF + n], sW + k * LD + br * 16);
            float mmax = 0.f;
            if (i < gv.P) {
                const float e0 = edge_attr[i * 3 + 0], e1 = edge_attr[i * 3 + 1], e2 = edge_attr[i * 3 + 2];
                fma_row16(acc, e0, sW + 30 * LD + br * 16);
                fma_row16(acc, e1, sW + 31 * LD + br * 16);
                fma_row16(acc, e2, sW + 32 * LD + br * 16);
                const float4 mv = reinterpret_cast<const float4*>(mask)[i];
                mmax = fmaxf(fmaxf(mv.x, mv.y), fmaxf(mv.z, mv.w));
            }
#pragma unroll
            for (int o = 0; o < 16; ++o)
                if (br * 16 + o < 30) Hs[(br * 16 + o) * LDF + n] = (i < gv.P) ? mmax * prelu(acc[o], ri_a1) : 0.f;
            if (br == 0) gid[n] = (i < gv.P) ? node_grid(gv, i) : -1;
            __syncthreads();
            // segmented sum over runs of equal grid node inside the tile; one atomic per (run, channel, 16-node strip)
            if (threadIdx.x < 240) {
                const int c = threadIdx.x >> 3, q = threadIdx.x & 7;
                int cur = gid[q * 16];
                float sum = 0.f;
#pragma unroll 4
                for (int t = 0; t < 16; ++t) {
                    const int m = q * 16 + t;
                    const int g = gid[m];
                    if (g != cur) {
                        if (cur >= 0) atomicAdd(&xg[(int64_t)cur * 32 + c], sum);
                        cur = g;
                        sum = 0.f;
                    }
                    sum += Hs[c * LDF + m];
                }
                if (cur >= 0) atomicAdd(&xg[(int64_t)cur * 32 + c], sum);
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K4: out = PReLU(fc2 xg)
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) readin_finalize_kernel(const float* __restrict__ packed,
                                                              const float* __restrict__ xg, float* __restrict__ out,
                                                              int ld_out, int G) {
    __shared__ __align__(16) float sW[30 * LD16 + LD16];
    for (int i = threadIdx.x; i < 30 * LD16 + LD16; i += 128) sW[i] = packed[RI_WFC2 + i];
    const float a2 = packed[RI_SLOPES + 1];
    __syncthreads();
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= G) return;
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = sW[30 * LD16 + o];
    const float4* row = reinterpret_cast<const float4*>(xg + (int64_t)g * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 x = row[c];
        fma_row16(acc, x.x, sW + (4 * c + 0) * LD16);
        fma_row16(acc, x.y, sW + (4 * c + 1) * LD16);
        if (4 * c + 2 < 30) fma_row16(acc, x.z, sW + (4 * c + 2) * LD16);
        if (4 * c + 3 < 30) fma_row16(acc, x.w, sW + (4 * c + 3) * LD16);
    }
#pragma unroll
    for (int o = 0; o < 15; ++o) out[(int64_t)g * ld_out + o] = prelu(acc[o], a2);
}

}  // namespace

// --------------------------------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------------------------------
template <bool FUSED, bool BF16>
static int launch_da_init_t(const genie_plan* p, const float* packed, const InitInputs& in, float* tr0, bool tc_plan,
                            cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    if (P == 0) return GENIE_OK;
    const int64_t blocks = (P + K1_THREADS - 1) / K1_THREADS;
    GENIE_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_init_slots, packed + DA_W0, sizeof(float) * (8 * LD + LD),
                                             sizeof(float) * (8 * LD + LD) * p->cslot, cudaMemcpyDeviceToDevice, st));
    TimedLaunch tl(KID_DA_INIT, st);
#define GENIE_INIT_CASE(C)                                                                                          \
    case C:                                                                                                         \
        da_init_kernel<C, FUSED, BF16><<<(unsigned)blocks, K1_THREADS, 0, st>>>(packed, in, tr0, P, tc_plan ? 1 : 0, \
                                                                                p->init_sta, p->init_src, p->g.n_sta); \
        break;
    switch (p->cslot) {
        GENIE_INIT_CASE(0) GENIE_INIT_CASE(1) GENIE_INIT_CASE(2) GENIE_INIT_CASE(3)
        default: set_error("launch_da_init: bad constant slot"); return GENIE_ERR_INVALID;
    }
#undef GENIE_INIT_CASE
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_da_init(const genie_plan* p, const float* packed, const float* slice, const float* mask, float* tr0,
                   bool tc_plan, cudaStream_t st) {
    InitInputs in = {};
    in.slice = slice;
    in.mask = mask;
    return p->storage == GENIE_STORAGE_BF16 ? launch_da_init_t<false, true>(p, packed, in, tr0, tc_plan, st)
                                            : launch_da_init_t<false, false>(p, packed, in, tr0, tc_plan, st);
}

// a1 fused into layer 0 (genie_window_fwd): CARTESIAN plans with P < 2^31.
int launch_da_init_fused(const genie_plan* p, const float* packed, const WindowParamSrc& ws, const int32_t* ind_use,
                         const float* trv, const float* series, float* slice_out, float* mask_out, float* tr0,
                         bool tc_plan, cudaStream_t st) {
    if (p->g.mode != GENIE_GRAPH_CARTESIAN || p->g.n_prod >= (int64_t)0x7fffffff) {
        set_error("launch_da_init_fused: needs a CARTESIAN plan with fewer than 2^31 product nodes");
        return GENIE_ERR_UNSUPPORTED;
    }
    InitInputs in = {};
    in.ws = ws;
    in.ind_use = ind_use;
    in.trv = trv;
    in.series = series;
    in.slice_out = slice_out;
    in.mask_out = slice_out ? mask_out : nullptr;
    return p->storage == GENIE_STORAGE_BF16 ? launch_da_init_t<true, true>(p, packed, in, tr0, tc_plan, st)
                                            : launch_da_init_t<true, false>(p, packed, in, tr0, tc_plan, st);
}

int launch_da_layer1(const genie_plan* p, const float* packed, const float* tr0, const float* mask, float* zc, float* va,
                     float* vb, bool tc_plan, cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    if (P == 0) return GENIE_OK;
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)K2_SMEM));
        attr_set.mark();
    }
    const int64_t n_tiles = (P + TM - 1) / TM;
    const int64_t grid = n_tiles < (int64_t)p->sm_count * 2 ? n_tiles : (int64_t)p->sm_count * 2;
    TimedLaunch tl(KID_DA_LAYER1, st);
    da_layer1_kernel<<<(unsigned)grid, K2_THREADS, K2_SMEM, st>>>(make_view(p), packed, tr0, mask, zc, va, vb, n_tiles,
                                                              tc_plan ? 1 : 0);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_da_layer2_readin(const genie_plan* p, const float* packed, int mode, const float* zc, const float* va,
                            const float* vb, const float* latent_in, float* latent_out, const float* edge_attr,
                            const float* mask, float* xg, cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    if (P == 0) return GENIE_OK;
    const int64_t n_tiles = (P + TM - 1) / TM;
    const int64_t cap = (int64_t)p->sm_count * 4;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    const GraphView gv = make_view(p);
    TimedLaunch tl(KID_DA_LAYER2_READIN, st);
#define GENIE_L2_CASE(M)                                                                                              \
    case M:                                                                                                            \
        da_layer2_readin_kernel<M><<<grid, K3_THREADS, 0, st>>>(gv, packed, zc, va, vb, latent_in, latent_out,      \
                                                                  edge_attr, mask, xg, n_tiles);                     \
        break;
    switch (mode) {
        GENIE_L2_CASE(L2_GATHER | L2_READIN)
        GENIE_L2_CASE(L2_GATHER | L2_STORE_LATENT | L2_READIN)
        GENIE_L2_CASE(L2_GATHER | L2_STORE_LATENT)
        GENIE_L2_CASE(L2_READIN)
        default:
            set_error("launch_da_layer2_readin: unsupported mode");
            return GENIE_ERR_INVALID;
    }
#undef GENIE_L2_CASE
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_readin_finalize(const genie_plan* p, const float* packed, const float* xg, float* out, int ld_out,
                           cudaStream_t st) {
    const int G = p->g.n_grid_owned > 0 ? p->g.n_grid_owned : p->g.n_grid;     // halo grid nodes have no read-in row
    if (G == 0) return GENIE_OK;
    TimedLaunch tl(KID_READIN_FINALIZE, st);
    readin_finalize_kernel<<<(G + 127) / 128, 128, 0, st>>>(packed, xg, out, ld_out, G);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
