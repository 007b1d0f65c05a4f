// Association branch of forward / forward_fixed (module.py:983-991): the product-node-sized part.
//
//   mask_out = (max_t y[g,t] > 0.01)                                                            module.py:983
//   s0  = BipartiteGraphReadOutOperator(SpatialDirect(x_spatial), A_Lg_in_src, mask_out)        module.py:333-352
//   s   = DataAggregationAssociationPhase(s0, x_latent, mask_out[g(i)], Mask, A_in_sta, A_in_src)   module.py:356-403
//   arv = LocalSliceLgCollapse{P,S}(A_edges_{p,s}, dt_partition, tpick, ipick, phase, s, tlatent)   module.py:604-653
//
// Five kernels, split at the global dependencies:
//   assoc_grid_pre_kernel   per grid node: y_latent = SpatialDirect(x_spatial), the y_latent half of the read-out fc1
//                           (every product node of a grid node shares it) and the source mask
//   assoc_init_kernel       per product node (thread per node, weights broadcast from shared memory): read-out MLP ->
//                           init_trns (50 -> 30) -> the two layer-1 message maps l1_t1_1 / l1_t2_1 (unlike DataAggregation,
//                           the association phase DOES use them, module.py:394-395)
//   assoc_layer1_kernel     gather the two means -> l1_t*_2 -> l2_t*_1 -> the layer-2 linear maps split by linearity exactly as
//                           in da_kernels.cu (15-wide messages va / vb, node-local part zc)
//   assoc_layer2_kernel     gather va / vb means, add zc, PReLU -> s rows [o1(15) 0 | o2(15) 0]
//   assoc_collapse_kernel   one warp per (pick, phase): the k_infer = 10 product nodes the pointer table names, kept when their
//                           theoretical time is within 2 eps, PReLU(fc1 [s_j | dt/eps | phase]) averaged, fc2 -> arrival rows
// Layout of the packed weights: AS_* below, reported to the host by genie_assoc_layout (packed in genie_b200/ops.py).
#include "common.cuh"
#include "assoc_layout.h"
#include "gather.cuh"

using namespace gl;


namespace {

// Two-wide accumulators (Blackwell FFMA2, common.cuh): acc[o] holds outputs 2o, 2o+1.  acc += a * wrow[0..29]; wrow is a
// 16-byte aligned shared-memory row read with the same address by all lanes (broadcast).  Same rounding as scalar FMAs.
__device__ __forceinline__ void fma2_row30(f32x2_t (&acc)[15], float a, const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const float4 w = w4[c];
        ffma2(acc[2 * c], a, w.x, w.y);
        ffma2(acc[2 * c + 1], a, w.z, w.w);
    }
    const float2 w = *reinterpret_cast<const float2*>(wrow + 28);
    ffma2(acc[14], a, w.x, w.y);
}
__device__ __forceinline__ void fma2_row16(f32x2_t (&acc)[8], float a, const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 w = w4[c];
        ffma2(acc[2 * c], a, w.x, w.y);
        ffma2(acc[2 * c + 1], a, w.z, w.w);
    }
}
__device__ __forceinline__ void init2_30(f32x2_t (&acc)[15], const float* __restrict__ b) {
#pragma unroll
    for (int o = 0; o < 15; ++o) acc[o] = pack2(b[2 * o], b[2 * o + 1]);
}
__device__ __forceinline__ void init2_16(f32x2_t (&acc)[8], const float* __restrict__ b) {
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = pack2(b[2 * o], b[2 * o + 1]);
}

constexpr int TM = 128;
constexpr int LDF = TM + 1;

// --------------------------------------------------------------------------------------------------------------------
// per grid node
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) assoc_grid_pre_kernel(const float* __restrict__ packed, const float* __restrict__ x_spatial,
                                                             int ld_x, const float* __restrict__ y, int T, int G,
                                                             float thresh, float* __restrict__ yfc1,
                                                             float* __restrict__ mask_out) {
    __shared__ __align__(16) float sW[as::RO_WA];    // SD_W .. RO_B1
    for (int i = threadIdx.x; i < as::RO_WA; i += 128) sW[i] = packed[i];
    const float a_sd = packed[as::SL + as::SL_SD];
    __syncthreads();
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= G) return;
    float yl[30];
#pragma unroll
    for (int o = 0; o < 30; ++o) yl[o] = sW[as::SD_B + o];
    for (int k = 0; k < 30; ++k) fma_row30(yl, x_spatial[(int64_t)g * ld_x + k], sW + as::SD_W + k * 32);
    float acc[30];
#pragma unroll
    for (int o = 0; o < 30; ++o) acc[o] = sW[as::RO_B1 + o];
#pragma unroll
    for (int k = 0; k < 30; ++k) fma_row30(acc, prelu(yl[k], a_sd), sW + as::RO_WY + k * 32);
#pragma unroll
    for (int o = 0; o < 30; ++o) yfc1[(int64_t)g * 32 + o] = acc[o];
    yfc1[(int64_t)g * 32 + 30] = 0.f;
    yfc1[(int64_t)g * 32 + 31] = 0.f;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, y[(int64_t)g * T + t]);
    mask_out[g] = (T > 0 && m > thresh) ? 1.f : 0.f;
}

// --------------------------------------------------------------------------------------------------------------------
// per product node: read-out MLP -> init_trns -> layer-1 message maps
// --------------------------------------------------------------------------------------------------------------------
constexpr int K0_THREADS = 128;
constexpr int K0_W0 = as::RO_WA;                       // first packed float staged in shared memory
constexpr int K0_W_FLOATS = as::INIT_END - K0_W0;
// Every thread owns NPT = 2 nodes (rows lane and lane + 32 of the warp's 64 consecutive nodes): a weight row read from shared
// memory (a broadcast LDS.128 still returns 512 bytes per warp — the pipe that bounded the one-node version at 18.6 ms at
// 1000 x 50000) feeds two accumulator sets.
constexpr int K0_NPT = 2, K0_WROWS = 32 * K0_NPT;
// per-warp staging (floats): the warp's x_latent rows and edge-attribute rows, and one 32 x 32 output tile
constexpr int K0_XL = K0_WROWS * 30, K0_EA = K0_WROWS * 3, K0_OUT = 32 * 32;
constexpr int K0_WARP_FLOATS = K0_XL + K0_EA + K0_OUT;
constexpr size_t K0_SMEM = (size_t)(K0_W_FLOATS + (K0_THREADS / 32) * K0_WARP_FLOATS) * sizeof(float);
static_assert(K0_W_FLOATS % 4 == 0 && K0_WARP_FLOATS % 4 == 0, "16-byte alignment of the staging areas");

// acc_u[0..N) += a_u * wrow[0..N) for the thread's two nodes: one shared-memory read per weight chunk
// (two-wide FFMA2 on adjacent output pairs: the same per-element fused multiply-adds in half the issue slots)
__device__ __forceinline__ void fma2_pair(float& o0, float& o1, f32x2_t aa, float w0, float w1) {
    f32x2_t acc = pack2(o0, o1);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(pack2(w0, w1)));
    unpack2(acc, o0, o1);
}
__device__ __forceinline__ void fma_row30x2(float (&acc)[K0_NPT][30], const float (&a)[K0_NPT], const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    const f32x2_t aa[K0_NPT] = {pack2(a[0], a[0]), pack2(a[1], a[1])};
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const float4 w = w4[c];
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
            fma2_pair(acc[u][4 * c + 0], acc[u][4 * c + 1], aa[u], w.x, w.y);
            fma2_pair(acc[u][4 * c + 2], acc[u][4 * c + 3], aa[u], w.z, w.w);
        }
    }
    const float2 w = *reinterpret_cast<const float2*>(wrow + 28);
#pragma unroll
    for (int u = 0; u < K0_NPT; ++u) fma2_pair(acc[u][28], acc[u][29], aa[u], w.x, w.y);
}
__device__ __forceinline__ void fma_row16x2(float (&acc)[K0_NPT][16], const float (&a)[K0_NPT], const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    const f32x2_t aa[K0_NPT] = {pack2(a[0], a[0]), pack2(a[1], a[1])};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 w = w4[c];
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
            fma2_pair(acc[u][4 * c + 0], acc[u][4 * c + 1], aa[u], w.x, w.y);
            fma2_pair(acc[u][4 * c + 2], acc[u][4 * c + 3], aa[u], w.z, w.w);
        }
    }
}

// Rows of 32 floats of 32 consecutive nodes (thread = node) -> global memory as 512-byte runs: the rows pass through a
// warp-private 4 KB tile (16-byte chunks XOR-swizzled by row: conflict free both ways).  A thread-per-row STG.128 touches
// 32 lines per instruction; this is 8 instructions of 4 lines each.
__device__ __forceinline__ void store_rows32(float* __restrict__ dst_row0, const float (&v)[30], float* tile, int lane, int n_rows) {
    float4* t4 = reinterpret_cast<float4*>(tile);
#pragma unroll
    for (int c = 0; c < 7; ++c) t4[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    t4[lane * 8 + (7 ^ (lane & 7))] = make_float4(v[28], v[29], 0.f, 0.f);
    __syncwarp();
    float4* d = reinterpret_cast<float4*>(dst_row0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int row = 4 * j + (lane >> 3), c = lane & 7;
        const float4 x = t4[row * 8 + (c ^ (row & 7))];
        if (row < n_rows) __stcs(d + j * 32 + lane, x);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(K0_THREADS)
    assoc_init_kernel(const GraphView gv, const float* __restrict__ packed, const float* __restrict__ yfc1,
                      const float* __restrict__ mask_out, const float* __restrict__ edge_attr,
                      const float* __restrict__ x_latent, const float* __restrict__ mask, const float* __restrict__ init_sta,
                      const float* __restrict__ init_src, float* __restrict__ s0_out, float* __restrict__ tr_out,
                      float* __restrict__ a1_out, float* __restrict__ a2_out) {
    extern __shared__ __align__(16) float sW0[];
    for (int i = threadIdx.x; i < K0_W_FLOATS; i += K0_THREADS) sW0[i] = packed[K0_W0 + i];
    const float a_ro1 = packed[as::SL + as::SL_RO1], a_ro2 = packed[as::SL + as::SL_RO2];
    const float a_in = packed[as::SL + as::SL_A], a11 = packed[as::SL + as::SL_A11], a12 = packed[as::SL + as::SL_A12];
    __syncthreads();
    const float* sW = sW0 - K0_W0;                     // index with the packed offsets
    const int lane = threadIdx.x & 31;
    const int64_t i0 = ((int64_t)blockIdx.x * (K0_THREADS / 32) + (threadIdx.x >> 5)) * K0_WROWS;      // first node of the warp
    if (i0 >= gv.P) return;
    const int n_rows = (int)min((int64_t)K0_WROWS, gv.P - i0);
    // the warp's x_latent rows (30 floats each) and edge-attribute rows are contiguous: coalesced 16-byte loads into shared
    // memory, then every thread reads its own rows (8-byte reads at a 120-byte stride: conflict free)
    float* sXL = sW0 + K0_W_FLOATS + (threadIdx.x >> 5) * K0_WARP_FLOATS;
    float* sEA = sXL + K0_XL;
    float* sOUT = sEA + K0_EA;
    {
        const float4* src = reinterpret_cast<const float4*>(x_latent + i0 * 30);     // i0 is a multiple of 64: 16-byte aligned
        const int n4 = n_rows * 30 / 4, rem = n_rows * 30 - 4 * n4;
#pragma unroll
        for (int j = 0; j < K0_XL / 128; ++j) {
            const int k = j * 32 + lane;
            if (k < n4) reinterpret_cast<float4*>(sXL)[k] = __ldg(src + k);
        }
        if (lane < rem) sXL[4 * n4 + lane] = __ldg(x_latent + i0 * 30 + 4 * n4 + lane);
        const float* ea = edge_attr + i0 * 3;
#pragma unroll
        for (int j = 0; j < K0_EA / 32; ++j)
            if (j * 32 + lane < n_rows * 3) sEA[j * 32 + lane] = __ldg(ea + j * 32 + lane);
    }
    __syncwarp();
    bool live[K0_NPT];
    int row[K0_NPT], g[K0_NPT];
    int64_t node[K0_NPT];
    float mo[K0_NPT];
#pragma unroll
    for (int u = 0; u < K0_NPT; ++u) {
        live[u] = lane + 32 * u < n_rows;
        row[u] = live[u] ? lane + 32 * u : 0;          // idle threads of the last warp shadow its first node (no stores)
        node[u] = i0 + row[u];
        g[u] = node_grid(gv, node[u]);
        mo[u] = mask_out[g[u]];
    }
    float tr[K0_NPT][30];
    {
        // read-out: s0 = PReLU(fc2(mask_j * PReLU(fc1 [y_latent_g | attr_i])))       (module.py:346, 350-352)
        float h[K0_NPT][30];
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
            const float4* yr = reinterpret_cast<const float4*>(yfc1 + (int64_t)g[u] * 32);
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const float4 v = __ldg(yr + c);
                h[u][4 * c] = v.x; h[u][4 * c + 1] = v.y; h[u][4 * c + 2] = v.z; h[u][4 * c + 3] = v.w;
            }
            const float4 v = __ldg(yr + 7);
            h[u][28] = v.x; h[u][29] = v.y;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float a[K0_NPT] = {sEA[row[0] * 3 + k], sEA[row[1] * 3 + k]};
            fma_row30x2(h, a, sW + as::RO_WA + k * 32);
        }
        float s[K0_NPT][16];
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u)
#pragma unroll
            for (int o = 0; o < 16; ++o) s[u][o] = sW[as::RO_B2 + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) {
            const float a[K0_NPT] = {mo[0] * prelu(h[0][k], a_ro1), mo[1] * prelu(h[1][k], a_ro1)};
            fma_row16x2(s, a, sW + as::RO_W2 + k * 16);
        }
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
#pragma unroll
            for (int o = 0; o < 15; ++o) s[u][o] = prelu(s[u][o], a_ro2);
            if (s0_out != nullptr && live[u]) {
#pragma unroll
                for (int o = 0; o < 15; ++o) s0_out[node[u] * 15 + o] = s[u][o];
            }
        }
        // init_trns [s0 | x_latent | mask_out | Mask]                                (module.py:389-391)
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
#pragma unroll
            for (int o = 0; o < 30; ++o) tr[u][o] = sW[as::AI_B + o];
            if (init_sta != nullptr) {
                // use_absolute_pos (module.py:987-988): the six position channels of init_trns as additive terms
                const float* ts = init_sta + (init_src != nullptr ? node[u] - (int64_t)g[u] * gv.S : node[u]) * 32;
#pragma unroll
                for (int o = 0; o < 30; ++o) tr[u][o] += __ldg(ts + o);
                if (init_src != nullptr) {
                    const float* tg = init_src + (int64_t)g[u] * 32;
#pragma unroll
                    for (int o = 0; o < 30; ++o) tr[u][o] += __ldg(tg + o);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            const float a[K0_NPT] = {s[0][k], s[1][k]};
            fma_row30x2(tr, a, sW + as::AI_W + k * 32);
        }
        const float2* xl0 = reinterpret_cast<const float2*>(sXL + row[0] * 30);
        const float2* xl1 = reinterpret_cast<const float2*>(sXL + row[1] * 30);
#pragma unroll 5
        for (int k = 0; k < 15; ++k) {
            const float2 v0 = xl0[k], v1 = xl1[k];
            const float ax[K0_NPT] = {v0.x, v1.x}, ay[K0_NPT] = {v0.y, v1.y};
            fma_row30x2(tr, ax, sW + as::AI_W + (15 + 2 * k) * 32);
            fma_row30x2(tr, ay, sW + as::AI_W + (16 + 2 * k) * 32);
        }
        fma_row30x2(tr, mo, sW + as::AI_W + 45 * 32);
        const float4 mv0 = __ldg(reinterpret_cast<const float4*>(mask) + node[0]);
        const float4 mv1 = __ldg(reinterpret_cast<const float4*>(mask) + node[1]);
        {
            const float m0[K0_NPT] = {mv0.x, mv1.x}, m1[K0_NPT] = {mv0.y, mv1.y}, m2[K0_NPT] = {mv0.z, mv1.z}, m3[K0_NPT] = {mv0.w, mv1.w};
            fma_row30x2(tr, m0, sW + as::AI_W + 46 * 32);
            fma_row30x2(tr, m1, sW + as::AI_W + 47 * 32);
            fma_row30x2(tr, m2, sW + as::AI_W + 48 * 32);
            fma_row30x2(tr, m3, sW + as::AI_W + 49 * 32);
        }
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u)
#pragma unroll
            for (int o = 0; o < 30; ++o) tr[u][o] = prelu(tr[u][o], a_in);
    }
#pragma unroll
    for (int u = 0; u < K0_NPT; ++u) store_rows32(tr_out + (i0 + 32 * u) * 32, tr[u], sOUT, lane, n_rows - 32 * u);
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {                   // the two layer-1 message maps, one after the other (register budget)
        const int wo = br ? as::M12_W : as::M11_W, bo = br ? as::M12_B : as::M11_B;
        const float sl = br ? a12 : a11;
        float m[K0_NPT][30];
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u)
#pragma unroll
            for (int o = 0; o < 30; ++o) m[u][o] = sW[bo + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) {
            const float a[K0_NPT] = {tr[0][k], tr[1][k]};
            fma_row30x2(m, a, sW + wo + k * 32);
        }
#pragma unroll
        for (int u = 0; u < K0_NPT; ++u) {
#pragma unroll
            for (int o = 0; o < 30; ++o) m[u][o] = prelu(m[u][o], sl);
            store_rows32((br ? a2_out : a1_out) + (i0 + 32 * u) * 32, m[u], sOUT, lane, n_rows - 32 * u);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// layer 1 (+ the node-local part of layer 2); the structure of da_layer1_kernel with five mask channels
// --------------------------------------------------------------------------------------------------------------------
constexpr int K2_THREADS = 256;
constexpr int K2_W_FLOATS = as::L1_END - as::W11;
constexpr int K2_F_ROWS = 95;   // 0-29 tr | 30-59 mean_sta | 60-89 mean_src | 90-94 mask5 ; rows 0-59 later hold the new tr
constexpr size_t K2_SMEM = (size_t)(K2_W_FLOATS + K2_F_ROWS * LDF) * sizeof(float);

__global__ void __launch_bounds__(K2_THREADS, 2)
    assoc_layer1_kernel(const GraphView gv, const float* __restrict__ packed, const float* __restrict__ tr_in,
                        const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ msrc,
                        const float* __restrict__ mask_out, const float* __restrict__ mask, const float* __restrict__ edge_sta,
                        const float* __restrict__ edge_src, float* __restrict__ zc, float* __restrict__ va,
                        float* __restrict__ vb, int64_t n_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* sW0 = smem;
    float* F = smem + K2_W_FLOATS;
    {
        const float4* src = reinterpret_cast<const float4*>(packed + as::W11);
        float4* dst = reinterpret_cast<float4*>(sW0);
        for (int i = threadIdx.x; i < K2_W_FLOATS / 4; i += K2_THREADS) dst[i] = src[i];
    }
    const float s_a1 = packed[as::SL + as::SL_A1], s_a21 = packed[as::SL + as::SL_A21], s_a22 = packed[as::SL + as::SL_A22];
    __syncthreads();
    const float* sW = sW0 - as::W11;                   // index with the packed offsets
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
                own = tr_in[i * 32 + lane];
                m1 = gather_mean32(a1, rs, gv.sta_col, 1.f, lane);
                // source neighbours: pre-averaged by the source pass (src_mean_kernels.cu) on plans with tiling tables
                m2 = msrc != nullptr ? msrc[i * 32 + lane] : gather_mean32(a2, rg, gv.src_col, 1.f, lane);
                if (lane == 0) mk = mask_out[node_grid(gv, i)];
                else if (lane < 5) mk = mask[i * 4 + lane - 1];
            }
            if (lane < 30) {
                F[lane * LDF + m] = own;
                F[(30 + lane) * LDF + m] = m1;
                F[(60 + lane) * LDF + m] = m2;
            }
            if (lane < 5) F[(90 + lane) * LDF + m] = mk;
        }
        __syncthreads();
        // edge-feature model (genie_assoc_set_terms): per-node additive terms of this thread's branch, or NULL
        const float* et = nullptr;
        if (edge_sta != nullptr && i0 + n < gv.P) {
            int64_t idx = i0 + n;
            if (gv.mode == GENIE_GRAPH_CARTESIAN) {
                const int64_t g = idx / gv.S;
                idx = br ? g : idx - g * gv.S;
            }
            et = (br ? edge_src : edge_sta) + idx * GENIE_EDGE_TERM_LD;
        }
        // ---- stage B: tr = PReLU1([l1_t1_2(..) | l1_t2_2(..)]) --------------------------------------------------------
        {
            const float* W = sW + (br ? as::W12 : as::W11);
            f32x2_t acc[15];
            init2_30(acc, sW + (br ? as::B12 : as::B11));
            if (et != nullptr) {
#pragma unroll
                for (int o = 0; o < 15; ++o) fadd2(acc[o], pack2(__ldg(et + 2 * o), __ldg(et + 2 * o + 1)));
            }
#pragma unroll 2
            for (int k = 0; k < 30; ++k) fma2_row30(acc, F[k * LDF + n], W + k * 32);
            const float* Fm = F + (30 + 30 * br) * LDF;
#pragma unroll 2
            for (int k = 0; k < 30; ++k) fma2_row30(acc, Fm[k * LDF + n], W + (30 + k) * 32);
#pragma unroll
            for (int k = 0; k < 5; ++k) fma2_row30(acc, F[(90 + k) * LDF + n], W + (60 + k) * 32);
            __syncthreads();   // every thread has finished reading rows 0-89
#pragma unroll
            for (int o = 0; o < 15; ++o) {
                float lo, hi;
                unpack2(acc[o], lo, hi);
                F[(br * 30 + 2 * o) * LDF + n] = prelu(lo, s_a1);
                F[(br * 30 + 2 * o + 1) * LDF + n] = prelu(hi, s_a1);
            }
        }
        __syncthreads();
        // ---- stage C: h = PReLU(l2_t*_1 tr);  v = W_agg h;  c = W_tr tr + W_m mask + b ---------------------------------
        {
            const float* Wa = sW + (br ? as::W22A : as::W21A);
            const float ah = br ? s_a22 : s_a21;
            f32x2_t h[15];
            init2_30(h, sW + (br ? as::B22A : as::B21A));
            const float* Wc = sW + (br ? as::WCB : as::WCA);
            f32x2_t c[8];
            init2_16(c, sW + (br ? as::BCB : as::BCA));
            if (et != nullptr) {
#pragma unroll
                for (int o = 0; o < 8; ++o) fadd2(c[o], pack2(__ldg(et + 32 + 2 * o), __ldg(et + 32 + 2 * o + 1)));
            }
#pragma unroll 2
            for (int k = 0; k < 60; ++k) {
                const float x = F[k * LDF + n];
                fma2_row30(h, x, Wa + k * 32);
                fma2_row16(c, x, Wc + k * 16);
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) fma2_row16(c, F[(90 + k) * LDF + n], Wc + (60 + k) * 16);
            f32x2_t v[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) v[o] = pack2(0.f, 0.f);
            const float* Wv = sW + (br ? as::WVB : as::WVA);
#pragma unroll
            for (int k = 0; k < 15; ++k) {
                float lo, hi;
                unpack2(h[k], lo, hi);
                fma2_row16(v, prelu(lo, ah), Wv + (2 * k) * 16);
                fma2_row16(v, prelu(hi, ah), Wv + (2 * k + 1) * 16);
            }
            const int64_t i = i0 + n;
            if (i < gv.P) {
                float4* zp = reinterpret_cast<float4*>(zc + i * LD_ZC + br * 16);
                float4* vp = reinterpret_cast<float4*>((br ? vb : va) + i * LD_V);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 cz, vz;
                    unpack2(c[2 * q], cz.x, cz.y);
                    unpack2(c[2 * q + 1], cz.z, cz.w);
                    unpack2(v[2 * q], vz.x, vz.y);
                    unpack2(v[2 * q + 1], vz.z, vz.w);
                    zp[q] = cz;
                    vp[q] = vz;
                }
            }
        }
        __syncthreads();   // F is rewritten by the next tile's gather
    }
}

// --------------------------------------------------------------------------------------------------------------------
// layer 2: s = PReLU2(zc + [mean_sta va | mean_src vb]); one warp per product node
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assoc_layer2_kernel(const GraphView gv, const float* __restrict__ packed,
                                                           const float* __restrict__ zc, const float* __restrict__ va,
                                                           const float* __restrict__ vb, float* __restrict__ s_out) {
    const float a2 = packed[as::SL + as::SL_A2];
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < gv.P; i += warps) {
        NbrRange rs, rg;
        node_ranges(gv, i, rs, rg);
        const float own = zc[i * LD_ZC + lane];
        const float mean = gather_mean16x2(va, vb, rs, rg, gv.sta_col, gv.src_col, lane);
        s_out[i * as::LD_S + lane] = (lane & 15) < 15 ? prelu(own + mean, a2) : 0.f;
    }
}

// Same, on plans with tiling tables: the source half was pre-averaged by the source pass (m2src [P][16]), so only the
// station half is gathered — by half-warps, two product nodes per warp pass.
__global__ void __launch_bounds__(256) assoc_layer2_pre_kernel(const GraphView gv, const float* __restrict__ packed,
                                                               const float* __restrict__ zc, const float* __restrict__ va,
                                                               const float* __restrict__ m2src, float* __restrict__ s_out) {
    const float a2 = packed[as::SL + as::SL_A2];
    const int lane = threadIdx.x & 31, half = lane >> 4, l = lane & 15;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n_pairs = (gv.P + 1) / 2;
    for (int64_t pr = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pr < n_pairs; pr += warps) {
        const int64_t i0 = 2 * pr, i1 = min(2 * pr + 1, gv.P - 1);        // an odd tail pair repeats its last node
        NbrRange r0, r1, unused;
        node_ranges(gv, i0, r0, unused);
        node_ranges(gv, i1, r1, unused);
        const float mean = gather_mean16x2(va, va, r0, r1, gv.sta_col, gv.sta_col, lane);
        const int64_t i = half ? i1 : i0;
        const float o1 = zc[i * LD_ZC + l] + mean;
        const float o2 = zc[i * LD_ZC + 16 + l] + m2src[i * LD_V + l];
        if (half == 0 || i1 != i0) {
            s_out[i * as::LD_S + l] = l < 15 ? prelu(o1, a2) : 0.f;
            s_out[i * as::LD_S + 16 + l] = l < 15 ? prelu(o2, a2) : 0.f;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// LocalSliceLgCollapse P and S: one warp per (pick, phase); lanes = hidden channels
// --------------------------------------------------------------------------------------------------------------------
constexpr int KC_WARPS = 4;

__global__ void __launch_bounds__(KC_WARPS * 32)
    assoc_collapse_kernel(const float* __restrict__ packed, const float* __restrict__ s_rows, int64_t P,
                          const int64_t* __restrict__ edges_p, const int64_t* __restrict__ edges_s,
                          const float* __restrict__ tlatent, const float* __restrict__ tpick,
                          const int64_t* __restrict__ ipick, const float* __restrict__ phase_label, int n_arv,
                          int l_dt, int k_infer, float dt0, float dt_step, float eps, float* __restrict__ arrival) {
    __shared__ __align__(16) float sW[2 * as::C_SIZE];
    for (int i = threadIdx.x; i < 2 * as::C_SIZE; i += KC_WARPS * 32) sW[i] = packed[as::CP_W1 + i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * KC_WARPS + (threadIdx.x >> 5);
    const int ph = blockIdx.y;                         // 0: P (LocalSliceLgCollapseP), 1: S
    if (a > n_arv) return;
    float* out = arrival + (int64_t)a * 30 + ph * 15;
    if (a == n_arv) {                                  // the null arrival of module.py:709-710
        if (lane < 15) out[lane] = 0.f;
        return;
    }
    const float* W1 = sW + ph * as::C_SIZE;
    const float* B1 = W1 + (as::CP_B1 - as::CP_W1);
    const float* W2 = W1 + (as::CP_W2 - as::CP_W1);
    const float* B2 = W1 + (as::CP_B2 - as::CP_W1);
    const float s1 = packed[as::SL + (ph ? as::SL_CS1 : as::SL_CP1)], s2 = packed[as::SL + (ph ? as::SL_CS2 : as::SL_CP2)];
    const int64_t* __restrict__ edges = ph ? edges_s : edges_p;
    const float tp = tpick[a];
    const float phl = phase_label[a];
    // t_index = floor((tpick - dt_partition[0]) / dt) in fp32, as torch evaluates it (module.py:630)
    const float tq = floorf(__fdiv_rn(__fsub_rn(tp, dt0), dt_step));
    const int64_t base = (ipick[a] * (int64_t)l_dt + (int64_t)tq) * k_infer;
    const bool in_table = tq >= 0.f && tq < (float)l_dt;
    float acc = 0.f;
    int cnt = 0;
    for (int e = 0; e < k_infer && in_table; ++e) {
        const int64_t j = edges[base + e];
        if (j < 0 || j >= P) continue;
        const float t_rel = __fsub_rn(tp, tlatent[j * 2 + ph]);                      // module.py:637
        if (!(fabsf(t_rel) < 2.0f * eps)) continue;
        ++cnt;
        const float row = s_rows[j * as::LD_S + lane];
        float h = B1[lane];
#pragma unroll
        for (int k = 0; k < 30; ++k) h = fmaf(__shfl_sync(FULL_MASK, row, k < 15 ? k : k + 1), W1[k * 32 + lane], h);
        h = fmaf(__fdiv_rn(t_rel, eps), W1[30 * 32 + lane], h);
        h = fmaf(phl, W1[31 * 32 + lane], h);
        acc += prelu(h, s1);
    }
    const float mean = cnt > 0 ? acc / (float)cnt : 0.f;
    float o = lane < 15 ? B2[lane] : 0.f;
#pragma unroll
    for (int k = 0; k < 30; ++k) {
        const float mk = __shfl_sync(FULL_MASK, mean, k);
        if (lane < 15) o = fmaf(mk, W2[k * 16 + lane], o);
    }
    if (lane < 15) out[lane] = prelu(o, s2);
}

}  // namespace

// --------------------------------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------------------------------
size_t assoc_packed_floats() { return (size_t)as::FLOATS; }

int assoc_layout(int32_t* out, int n) {
    static const int32_t k[] = {as::SD_W, as::SD_B, as::RO_WY, as::RO_B1, as::RO_WA, as::RO_W2, as::RO_B2, as::AI_W, as::AI_B,
                                as::M11_W, as::M11_B, as::M12_W, as::M12_B, as::W11, as::W12, as::B11, as::B12, as::W21A,
                                as::W22A, as::B21A, as::B22A, as::WVA, as::WVB, as::WCA, as::WCB, as::BCA, as::BCB,
                                as::CP_W1, as::CP_B1, as::CP_W2, as::CP_B2, as::CS_W1, as::CS_B1, as::CS_W2, as::CS_B2, as::SL};
    const int count = (int)(sizeof(k) / sizeof(k[0]));
    if (!out || n < count) {
        set_error("genie_assoc_layout: need room for " + std::to_string(count) + " offsets");
        return GENIE_ERR_INVALID;
    }
    for (int i = 0; i < count; ++i) out[i] = k[i];
    return GENIE_OK;
}

AssocWorkspace carve_assoc_workspace(const genie_plan* p, void* base) {
    AssocWorkspace w;
    const size_t P = (size_t)p->g.n_prod, G = (size_t)p->g.n_grid;
    size_t off = 0;
    auto take = [&](size_t floats) {
        float* ptr = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
        off += (floats * sizeof(float) + 255) / 256 * 256;
        return ptr;
    };
    w.tr = take(P * 32);
    w.a1 = take(P * 32);
    w.a2 = take(P * 32);
    w.zc = take(P * 32);
    w.va = take(P * 16);
    w.vb = take(P * 16);
    w.msrc = take(P * 32);
    w.yfc1 = take(G * 32);
    w.mask_out = take(G);
    w.t2 = take(T2_FLOATS + 96);
    w.bytes = off;
    return w;
}

int launch_assoc_product(const genie_plan* p, const float* packed, const float* x_spatial, int ld_x, const float* y, int T,
                         float thresh, const float* edge_attr, const float* x_latent, const float* mask,
                         const AssocWorkspace& w, float* s0_out, float* mask_out_copy, cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    const int G = p->g.n_grid;
    if (P == 0 || G == 0) return GENIE_OK;
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(assoc_layer1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM));
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(assoc_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K0_SMEM));
        attr_set.mark();
    }
    const GraphView gv = make_view(p);
    {
        TimedLaunch tl(KID_ASSOC_GRID_PRE, st);
        assoc_grid_pre_kernel<<<(G + 127) / 128, 128, 0, st>>>(packed, x_spatial, ld_x, y, T, G, thresh, w.yfc1, w.mask_out);
        GENIE_LAUNCH_CHECK();
    }
    if (mask_out_copy != nullptr)
        GENIE_CUDA_CHECK(cudaMemcpyAsync(mask_out_copy, w.mask_out, sizeof(float) * (size_t)G, cudaMemcpyDeviceToDevice, st));
    {
        TimedLaunch tl(KID_ASSOC_INIT, st);
        assoc_init_kernel<<<(unsigned)((P + K0_THREADS * K0_NPT - 1) / (K0_THREADS * K0_NPT)), K0_THREADS, K0_SMEM, st>>>(
            gv, packed, w.yfc1, w.mask_out, edge_attr, x_latent, mask, p->assoc_init_sta, p->assoc_init_src, s0_out, w.tr,
            w.a1, w.a2);
        GENIE_LAUNCH_CHECK();
    }
    // plans with tiling tables: the two means over source neighbours come from the source pass of the front end
    // (src_mean_kernels.cu: grouped grid nodes, every neighbour row read once from DRAM) instead of per-node L2 gathers
    const bool split = split_supported(p);
    int rc;
    if (split && (rc = launch_src_mean(p, 32, w.a2, w.msrc, nullptr, st))) return rc;
    if (split) {
        // layer 1 on the two-pipeline tcgen05 station pass (da_s1_kernel.cu, ASSOC); the blob is rebuilt per call (one block,
        // microseconds) so that the entry point stays stateless in the weights
        if ((rc = launch_assoc_pack_t2(packed, w.t2, st))) return rc;
        if ((rc = launch_assoc_layer1_s(p, w.t2, w.tr, w.a1, w.msrc, mask, w.mask_out, w.zc, w.va, w.vb, st))) return rc;
    } else {
        const int64_t n_tiles = (P + TM - 1) / TM;
        const int64_t grid = n_tiles < (int64_t)p->sm_count * 2 ? n_tiles : (int64_t)p->sm_count * 2;
        TimedLaunch tl(KID_ASSOC_LAYER1, st);
        assoc_layer1_kernel<<<(unsigned)grid, K2_THREADS, K2_SMEM, st>>>(gv, packed, w.tr, w.a1, w.a2,
                                                                        split ? w.msrc : nullptr, w.mask_out, mask,
                                                                        p->assoc_edge_sta, p->assoc_edge_src, w.zc, w.va,
                                                                        w.vb, n_tiles);
        GENIE_LAUNCH_CHECK();
    }
    float* m2src = split ? w.a1 : nullptr;             // a1 is dead after layer 1: [P][16] fits in its [P][32]
    if (split && (rc = launch_src_mean(p, 16, w.vb, m2src, nullptr, st))) return rc;
    if (split && p->g.n_sta_tiles <= 32) {
        // the station half on chip as well: the layer-2 station pass of the front end without its read-in (da_s2_kernel.cu, ASSOC)
        if ((rc = launch_assoc_layer2_s(p, packed + as::SL + as::SL_A2, w.zc, w.va, m2src, w.tr, st))) return rc;
    } else {
        const int64_t blocks = split ? (P / 2 + 8) / 8 : (P + 7) / 8;
        const int64_t cap = (int64_t)p->sm_count * 16;
        const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        TimedLaunch tl(KID_ASSOC_LAYER2, st);
        if (split)
            assoc_layer2_pre_kernel<<<grid, 256, 0, st>>>(gv, packed, w.zc, w.va, m2src, w.tr);
        else
            assoc_layer2_kernel<<<grid, 256, 0, st>>>(gv, packed, w.zc, w.va, w.vb, w.tr);
        GENIE_LAUNCH_CHECK();
    }
    return GENIE_OK;
}

int launch_assoc_collapse(const float* packed, const float* s_rows, int64_t P, const int64_t* edges_p, const int64_t* edges_s,
                          const float* tlatent, const float* tpick, const int64_t* ipick, const float* phase_label, int n_arv,
                          int l_dt, int k_infer, float dt0, float dt_step, float eps, float* arrival, cudaStream_t st) {
    const dim3 grid((unsigned)((n_arv + 1 + KC_WARPS - 1) / KC_WARPS), 2);
    TimedLaunch tl(KID_ASSOC_COLLAPSE, st);
    assoc_collapse_kernel<<<grid, KC_WARPS * 32, 0, st>>>(packed, s_rows, P, edges_p, edges_s, tlatent, tpick, ipick,
                                                         phase_label, n_arv, l_dt, k_infer, dt0, dt_step, eps, arrival);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
