// Read-out heads of forward_fixed_source (module.py:1015-1020): the step right after the product-graph front end.
//
//   y = TemporalAttention(SpatialDirect(x_spatial), t_query)                                  per grid node   (:251-260, :299-331)
//   x = TemporalAttention(SpatialAttention(x_spatial, x_query, x_context), t_query)           per query point (:262-297)
//
// Both are a few thousand FMAs per node on a [G,30] input — latency, not bandwidth: ONE WARP per node (lane = output
// channel, inputs broadcast with shuffles, weight rows read coalesced from shared memory; round 1 ran one THREAD per node:
// 0.40 ms of a 0.81 ms window at 100 x 5000, profiles/r6b_c2_bench.json).  Two algebraic folds keep the state small:
//   * TemporalAttention's query branch depends only on t_query, so  s[n,t,h] = <c[n,h,:], q[t,h,:]> / sqrt(L)  with
//     c = f_context_2(h1) is the linear map  (q f_context_2.weight / sqrt(L)) h1 + q f_context_2.bias / sqrt(L):  the host
//     folds it once per (weights, t_query) into A [T*H, 30], a0 [T*H] (genie_b200/ops.py HeadsWeights);
//   * SpatialAttention's segment softmax runs over exactly k consecutive edges per query (knn), in two passes over the
//     edges (scores, then values) so that no per-edge 75-vector is kept.
// Layout of the packed head weights: HD_* in layout.h (every matrix K-major [n_in][ld]).
#include "common.cuh"

using namespace gl;

namespace {

constexpr int HEADS_THREADS = 256;             // 8 warps; ONE WARP PER NODE (grid node or query point), persistent blocks
constexpr int HW = HEADS_THREADS / 32;
constexpr int NH = 5, NL = 15;                 // heads, latent width (module.py:262, 299: n_heads = 5, n_latent = 15)
constexpr int TH_NC = 4;                       // (t, h) columns per lane: T * NH <= 32 * TH_NC, i.e. up to 25 query times
constexpr int SCR = 75 + 75 + 16 * NH + 2;     // per-warp scratch: values (75), per-output products (75), attention weights

// Warp-cooperative dense layers: lane = output channel; input k is broadcast from lane k with a shuffle, the weight row k
// (K-major, `ld` floats) is one coalesced, conflict-free shared-memory read.  A node's few thousand FMAs become ~60
// instructions per 30 x 30 layer on 32 lanes instead of 900 serial ones on one thread; a warp per node also gives the small
// grids of C1 / C2 (5000 nodes) thousands of warps instead of a few dozen threads per SM.
__device__ __forceinline__ float mv30(float in, int n_in, const float* __restrict__ W, const float* __restrict__ b, int lane) {
    float acc = lane < 30 ? b[lane] : 0.f;
#pragma unroll 6
    for (int k = 0; k < n_in; ++k) acc = fmaf(__shfl_sync(FULL_MASK, in, k), W[k * 32 + lane], acc);
    return acc;                                  // lanes 30, 31: bias-free garbage of the zero padding columns = 0
}
// 75 outputs: lane holds o = lane, lane + 32, lane + 64 (the last only for lanes < 11)
__device__ __forceinline__ void mv75(float in, int n_in, const float* __restrict__ W, float (&acc)[3], int lane) {
#pragma unroll 5
    for (int k = 0; k < n_in; ++k) {
        const float x = __shfl_sync(FULL_MASK, in, k);
        acc[0] = fmaf(x, W[k * 76 + lane], acc[0]);
        acc[1] = fmaf(x, W[k * 76 + 32 + lane], acc[1]);
        if (lane < 11) acc[2] = fmaf(x, W[k * 76 + 64 + lane], acc[2]);
    }
}
__host__ __device__ __forceinline__ int th_ld(int T) { return (T * NH + 31) / 32 * 32; }   // row stride of the transposed fold table
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(FULL_MASK, v, s);
    return v;
}

// TemporalAttention.forward (module.py:315-331) of one node by one warp: `in` = channel `lane` of the node's 30-vector;
// out[t], t < T.  sAt: the folded query table transposed, [k][th_ld(T)] (column t * NH + h); sA0: [T * NH].
__device__ __forceinline__ void temporal_attention_warp(const float* __restrict__ sW, const float* __restrict__ sAt,
                                                        const float* __restrict__ sA0, int T, float in, float* scr, int lane,
                                                        float* __restrict__ out) {
    const float a1 = sW[HD_TA_SL], a2 = sW[HD_TA_SL + 1], a4 = sW[HD_TA_SL + 2], a5 = sW[HD_TA_SL + 3];
    const float h1 = prelu(mv30(in, 30, sW + HD_TA_WC1, sW + HD_TA_BC1, lane), a1);
    const float h2 = prelu(mv30(in, 30, sW + HD_TA_WV1, sW + HD_TA_BV1, lane), a2);
    float v[3] = {sW[HD_TA_BV2 + lane], sW[HD_TA_BV2 + 32 + lane], lane < 11 ? sW[HD_TA_BV2 + 64 + lane] : 0.f};
    mv75(h2, 30, sW + HD_TA_WV2, v, lane);
    float* sv = scr;                              // values of the node, [h * NL + l]
    sv[lane] = v[0];
    sv[32 + lane] = v[1];
    if (lane < 11) sv[64 + lane] = v[2];
    // scores s[t, h] for all T * NH <= 48 (t, h) pairs: lane owns columns lane and lane + 32
    const int TH = T * NH, ld = th_ld(T);
    float s[TH_NC];
#pragma unroll
    for (int c = 0; c < TH_NC; ++c) s[c] = 32 * c + lane < TH ? sA0[32 * c + lane] : 0.f;
#pragma unroll 2
    for (int k = 0; k < 30; ++k) {
        const float x = __shfl_sync(FULL_MASK, h1, k);
#pragma unroll
        for (int c = 0; c < TH_NC; ++c)
            if (32 * c < TH) s[c] = fmaf(x, sAt[k * ld + 32 * c + lane], s[c]);     // warp-uniform guard; padding columns are 0
    }
    __syncwarp();
    for (int t = 0; t < T; ++t) {
        // z[l] = mean_h s[t, h] v[h, l]  (lanes l < NL), then proj_1 (15 -> 30), proj_2 (30 -> 1)
        float z = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int col = t * NH + h;
            const float sc = col < 32 ? s[0] : col < 64 ? s[1] : col < 96 ? s[2] : s[3];
            const float sh = __shfl_sync(FULL_MASK, sc, col & 31);
            z = fmaf(sh, lane < NL ? sv[h * NL + lane] : 0.f, z);
        }
        z = prelu(z / (float)NH, a4);
        const float p1 = prelu(mv30(z, NL, sW + HD_TA_WP1, sW + HD_TA_BP1, lane), a5);
        const float o = warp_sum(lane < 30 ? p1 * sW[HD_TA_WP2 + lane] : 0.f);
        if (lane == 0) out[t] = o + sW[HD_TA_BP2];
    }
    __syncwarp();
}

__device__ __forceinline__ void load_heads(float* sW, float* sAt, float* sA0, const float* __restrict__ packed,
                                           const float* __restrict__ fold, int T) {
    for (int i = threadIdx.x; i < HD_FLOATS / 4; i += HEADS_THREADS)
        reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(packed)[i];
    const int TH = T * NH, ld = th_ld(T);
    for (int i = threadIdx.x; i < 30 * ld; i += HEADS_THREADS) {
        const int k = i / ld, c = i - k * ld;
        sAt[i] = c < TH ? fold[c * 32 + k] : 0.f;           // the host's table is [T * NH][32]
    }
    for (int i = threadIdx.x; i < TH; i += HEADS_THREADS) sA0[i] = fold[TH * 32 + i];
    __syncthreads();
}

constexpr int PROJ_LD = 160;                   // per grid node: f_context(x_j) part at 0-74, f_values(x_j) part at 80-154

// y[g, t] for every grid node: SpatialDirect -> TemporalAttention.
// proj != NULL: also the x_j parts of SpatialAttention's per-edge layers, f_context[:, :30] x_g and f_values[:, :30] x_g
// (module.py:288-290: both are linear in [x_j | edge attr], and the x_j part does not depend on the query) — one 30 -> 75
// product per grid node here instead of one per (query, neighbour) pair in heads_query_kernel.
__global__ void __launch_bounds__(HEADS_THREADS)
    heads_grid_kernel(const float* __restrict__ packed, const float* __restrict__ fold, int T,
                      const float* __restrict__ x_spatial, int ld_x, int G, float* __restrict__ y, float* __restrict__ proj) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sAt = smem + HD_FLOATS;
    float* sA0 = sAt + 30 * th_ld(T);
    float* scr = sA0 + th_ld(T) + (threadIdx.x >> 5) * SCR;
    load_heads(sW, sAt, sA0, packed, fold, T);
    const int lane = threadIdx.x & 31;
    for (int g = blockIdx.x * HW + (threadIdx.x >> 5); g < G; g += gridDim.x * HW) {
        const float x = lane < 30 ? __ldg(x_spatial + (int64_t)g * ld_x + lane) : 0.f;
        if (proj != nullptr) {
            float pc[3] = {0.f, 0.f, 0.f}, pv[3] = {0.f, 0.f, 0.f};
            mv75(x, 30, sW + HD_SA_WC, pc, lane);
            mv75(x, 30, sW + HD_SA_WV, pv, lane);
            float* pr = proj + (int64_t)g * PROJ_LD;
            pr[lane] = pc[0];
            pr[32 + lane] = pc[1];
            if (lane < 11) pr[64 + lane] = pc[2];
            pr[80 + lane] = pv[0];
            pr[112 + lane] = pv[1];
            if (lane < 11) pr[144 + lane] = pv[2];
        }
        const float yl = prelu(mv30(x, 30, sW + HD_SD_W, sW + HD_SD_B, lane), sW[HD_SD_SL]);        // module.py:258-260
        temporal_attention_warp(sW, sAt, sA0, T, yl, scr, lane, y + (int64_t)g * T);
    }
}

// x[q, t] for every query point: SpatialAttention over its k nearest context nodes -> TemporalAttention.
// PROJ: the x_j parts of f_context / f_values come from heads_grid_kernel's table (one row per context node).
template <bool PROJ>
__global__ void __launch_bounds__(HEADS_THREADS)
    heads_query_kernel(const float* __restrict__ packed, const float* __restrict__ fold, int T,
                       const float* __restrict__ x_spatial, int ld_x, const float* __restrict__ x_context,
                       const float* __restrict__ x_query, const int64_t* __restrict__ nbr, int k_nbr, int Q,
                       float scale_rel, float* __restrict__ x_out, const float* __restrict__ proj) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sAt = smem + HD_FLOATS;
    float* sA0 = sAt + 30 * th_ld(T);
    float* scr = sA0 + th_ld(T) + (threadIdx.x >> 5) * SCR;
    float* sprod = scr + 75;                       // per-output products q * c of one edge
    float* salpha = scr + 150;                     // [edge][head]
    load_heads(sW, sAt, sA0, packed, fold, T);
    const int lane = threadIdx.x & 31;
    const float sqrt_l = sqrtf((float)NL);
    const float a1 = sW[HD_SA_SL], a2 = sW[HD_SA_SL + 1];
    for (int qi = blockIdx.x * HW + (threadIdx.x >> 5); qi < Q; qi += gridDim.x * HW) {
        const float qx = __ldg(x_query + (int64_t)qi * 3), qy = __ldg(x_query + (int64_t)qi * 3 + 1),
                    qz = __ldg(x_query + (int64_t)qi * 3 + 2);
        // lane e < k fetches neighbour e and its edge attribute once (one coalesced id load, the coordinate loads of all edges in
        // flight together); the edge loops below broadcast them with shuffles instead of chaining id -> coordinates -> row loads
        int64_t jl = 0;
        float eal[3] = {0.f, 0.f, 0.f};
        if (lane < k_nbr) {
            jl = __ldg(nbr + (int64_t)qi * k_nbr + lane);
            eal[0] = (qx - __ldg(x_context + jl * 3)) / scale_rel;
            eal[1] = (qy - __ldg(x_context + jl * 3 + 1)) / scale_rel;
            eal[2] = (qz - __ldg(x_context + jl * 3 + 2)) / scale_rel;
        }
        // ---- pass 1: attention scores of the k edges (module.py:288-292) --------------------------------------------------
        for (int e = 0; e < k_nbr; ++e) {
            const int64_t j = __shfl_sync(FULL_MASK, jl, e);
            const float ea[3] = {__shfl_sync(FULL_MASK, eal[0], e), __shfl_sync(FULL_MASK, eal[1], e),
                                 __shfl_sync(FULL_MASK, eal[2], e)};
            const float xj = (!PROJ && lane < 30) ? __ldg(x_spatial + j * ld_x + lane) : 0.f;
            float c[3], qv[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int o = lane + 32 * r;
                const bool ok = o < 75;
                c[r] = ok ? sW[HD_SA_BC + o] : 0.f;
                qv[r] = ok ? sW[HD_SA_BQ + o] : 0.f;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    c[r] = fmaf(ea[d], ok ? sW[HD_SA_WC + (30 + d) * 76 + o] : 0.f, c[r]);
                    qv[r] = fmaf(ea[d], ok ? sW[HD_SA_WQ + d * 76 + o] : 0.f, qv[r]);
                }
            }
            if (PROJ) {
                const float* pr = proj + j * PROJ_LD;
                c[0] += __ldg(pr + lane);
                c[1] += __ldg(pr + 32 + lane);
                if (lane < 11) c[2] += __ldg(pr + 64 + lane);
            } else {
                mv75(xj, 30, sW + HD_SA_WC, c, lane);
            }
            sprod[lane] = qv[0] * c[0];
            sprod[32 + lane] = qv[1] * c[1];
            if (lane < 11) sprod[64 + lane] = qv[2] * c[2];
            __syncwarp();
            if (lane < NH) {
                float d = 0.f;
#pragma unroll
                for (int l = 0; l < NL; ++l) d += sprod[lane * NL + l];
                salpha[e * NH + lane] = prelu(d / sqrt_l, a1);
            }
            __syncwarp();
        }
        // ---- softmax over the query's edges, per head (PyG softmax: max-shifted) -----------------------------------------
        if (lane < NH) {
            float m = -INFINITY;
            for (int e = 0; e < k_nbr; ++e) m = fmaxf(m, salpha[e * NH + lane]);
            float s = 0.f;
            for (int e = 0; e < k_nbr; ++e) {
                const float w = expf(salpha[e * NH + lane] - m);
                salpha[e * NH + lane] = w;
                s += w;
            }
            for (int e = 0; e < k_nbr; ++e) salpha[e * NH + lane] /= (s + 1e-16f);
        }
        __syncwarp();
        // ---- pass 2: weighted values, mean over heads, projection (module.py:293-297) --------------------------------------
        float outl = 0.f;                            // lanes l < NL
        for (int e = 0; e < k_nbr; ++e) {
            const int64_t j = __shfl_sync(FULL_MASK, jl, e);
            const float ea[3] = {__shfl_sync(FULL_MASK, eal[0], e), __shfl_sync(FULL_MASK, eal[1], e),
                                 __shfl_sync(FULL_MASK, eal[2], e)};
            const float xj = (!PROJ && lane < 30) ? __ldg(x_spatial + j * ld_x + lane) : 0.f;
            float v[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int o = lane + 32 * r;
                const bool ok = o < 75;
                v[r] = ok ? sW[HD_SA_BV + o] : 0.f;
#pragma unroll
                for (int d = 0; d < 3; ++d) v[r] = fmaf(ea[d], ok ? sW[HD_SA_WV + (30 + d) * 76 + o] : 0.f, v[r]);
            }
            if (PROJ) {
                const float* pr = proj + j * PROJ_LD + 80;
                v[0] += __ldg(pr + lane);
                v[1] += __ldg(pr + 32 + lane);
                if (lane < 11) v[2] += __ldg(pr + 64 + lane);
            } else {
                mv75(xj, 30, sW + HD_SA_WV, v, lane);
            }
            scr[lane] = v[0];
            scr[32 + lane] = v[1];
            if (lane < 11) scr[64 + lane] = v[2];
            __syncwarp();
            if (lane < NL) {
#pragma unroll
                for (int h = 0; h < NH; ++h) outl = fmaf(salpha[e * NH + h], scr[h * NL + lane], outl);
            }
            __syncwarp();
        }
        const float xq = prelu(mv30(outl / (float)NH, NL, sW + HD_SA_WP, sW + HD_SA_BP, lane), a2);
        temporal_attention_warp(sW, sAt, sA0, T, xq, scr, lane, x_out + (int64_t)qi * T);
    }
}

size_t heads_smem_bytes(int T) { return sizeof(float) * (size_t)(HD_FLOATS + 31 * th_ld(T) + HW * SCR + 4); }

// persistent blocks: 8 nodes per block pass, at most three blocks per SM (62 KB of shared memory each)
int heads_blocks(int n) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int need = (n + HW - 1) / HW;
    return need < 3 * sms ? need : 3 * sms;
}

}  // namespace

int launch_heads_grid(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, int G, float* y,
                      float* proj, cudaStream_t st) {
    if (G == 0 || T == 0) return GENIE_OK;
    const size_t smem = heads_smem_bytes(T);
    if (T * NH > 32 * TH_NC) {
        set_error("heads: at most 25 query times (T * 5 <= 128 columns of the folded query table)");
        return GENIE_ERR_UNSUPPORTED;
    }
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(heads_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = heads_blocks(G);
    TimedLaunch tl(KID_HEADS_GRID, st);
    heads_grid_kernel<<<blocks, HEADS_THREADS, smem, st>>>(packed, fold, T, x_spatial, ld_x, G, y, proj);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_heads_query(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, const float* x_context,
                       const float* x_query, const int64_t* nbr, int k_nbr, int Q, float scale_rel, float* x_out,
                       const float* proj, cudaStream_t st) {
    if (Q == 0 || T == 0) return GENIE_OK;
    if (k_nbr < 1 || k_nbr > 16) {
        set_error("heads: the query read-out supports 1..16 context neighbours per query");
        return GENIE_ERR_UNSUPPORTED;
    }
    const size_t smem = heads_smem_bytes(T);
    if (T * NH > 32 * TH_NC) {
        set_error("heads: at most 25 query times (T * 5 <= 128 columns of the folded query table)");
        return GENIE_ERR_UNSUPPORTED;
    }
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(heads_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(heads_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = heads_blocks(Q);
    TimedLaunch tl(KID_HEADS_QUERY, st);
    if (proj != nullptr)
        heads_query_kernel<true><<<blocks, HEADS_THREADS, smem, st>>>(packed, fold, T, x_spatial, ld_x, x_context, x_query, nbr,
                                                                      k_nbr, Q, scale_rel, x_out, proj);
    else
        heads_query_kernel<false><<<blocks, HEADS_THREADS, smem, st>>>(packed, fold, T, x_spatial, ld_x, x_context, x_query, nbr,
                                                                       k_nbr, Q, scale_rel, x_out, nullptr);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
