// Read-out heads of forward_fixed_source (module.py:1015-1020): the step right after the product-graph front end.
//
//   y = TemporalAttention(SpatialDirect(x_spatial), t_query)                                  per grid node   (:251-260, :299-331)
//   x = TemporalAttention(SpatialAttention(x_spatial, x_query, x_context), t_query)           per query point (:262-297)
//
// Both are a few thousand FMAs per node on a [G,30] input — latency, not bandwidth: one thread per node, weights
// broadcast from shared memory, everything in registers.  Two algebraic folds keep the per-thread state small:
//   * TemporalAttention's query branch depends only on t_query, so  s[n,t,h] = <c[n,h,:], q[t,h,:]> / sqrt(L)  with
//     c = f_context_2(h1) is the linear map  (q f_context_2.weight / sqrt(L)) h1 + q f_context_2.bias / sqrt(L):  the host
//     folds it once per (weights, t_query) into A [T*H, 30], a0 [T*H] (genie_b200/ops.py HeadsWeights);
//   * SpatialAttention's segment softmax runs over exactly k consecutive edges per query (knn), in two passes over the
//     edges (scores, then values) so that no per-edge 75-vector is kept.
// Layout of the packed head weights: HD_* in layout.h (every matrix K-major [n_in][ld]).
#include "common.cuh"

using namespace gl;

namespace {

constexpr int HEADS_THREADS = 64;
constexpr int NH = 5, NL = 15;                 // heads, latent width (module.py:262, 299: n_heads = 5, n_latent = 15)

// acc[0..N) += a * w[0..N)   (w: shared memory, 16-byte aligned, same address for all lanes -> broadcast)
template <int N>
__device__ __forceinline__ void fma_row(float (&acc)[N], float a, const float* __restrict__ w) {
    static_assert(N % 4 == 0 || N == 30 || N == 15 || N == 75, "row width");
    constexpr int N4 = N / 4;
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int c = 0; c < N4; ++c) {
        const float4 v = w4[c];
        acc[4 * c] = fmaf(a, v.x, acc[4 * c]);
        acc[4 * c + 1] = fmaf(a, v.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(a, v.z, acc[4 * c + 2]);
        acc[4 * c + 3] = fmaf(a, v.w, acc[4 * c + 3]);
    }
#pragma unroll
    for (int i = 4 * N4; i < N; ++i) acc[i] = fmaf(a, w[i], acc[i]);
}

// TemporalAttention.forward (module.py:315-331) of one node: in[30] -> out[t], t < T, written with stride 1.
__device__ __forceinline__ void temporal_attention_node(const float* __restrict__ sW, const float* __restrict__ sA,
                                                        const float* __restrict__ sA0, int T, const float (&in)[30],
                                                        float* __restrict__ out) {
    const float a1 = sW[HD_TA_SL], a2 = sW[HD_TA_SL + 1], a4 = sW[HD_TA_SL + 2], a5 = sW[HD_TA_SL + 3];
    float h1[30], v[75];
    {
#pragma unroll
        for (int o = 0; o < 30; ++o) h1[o] = sW[HD_TA_BC1 + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) fma_row<30>(h1, in[k], sW + HD_TA_WC1 + k * 32);
#pragma unroll
        for (int o = 0; o < 30; ++o) h1[o] = prelu(h1[o], a1);
    }
    {
        float h2[30];
#pragma unroll
        for (int o = 0; o < 30; ++o) h2[o] = sW[HD_TA_BV1 + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) fma_row<30>(h2, in[k], sW + HD_TA_WV1 + k * 32);
#pragma unroll
        for (int o = 0; o < 75; ++o) v[o] = sW[HD_TA_BV2 + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) fma_row<75>(v, prelu(h2[k], a2), sW + HD_TA_WV2 + k * 76);
    }
    for (int t = 0; t < T; ++t) {
        float s[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) s[h] = sA0[t * NH + h];
#pragma unroll
        for (int k = 0; k < 30; ++k) {
#pragma unroll
            for (int h = 0; h < NH; ++h) s[h] = fmaf(h1[k], sA[(t * NH + h) * 32 + k], s[h]);
        }
        float p1[30];
#pragma unroll
        for (int o = 0; o < 30; ++o) p1[o] = sW[HD_TA_BP1 + o];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float z = 0.f;
#pragma unroll
            for (int h = 0; h < NH; ++h) z = fmaf(s[h], v[h * NL + l], z);
            fma_row<30>(p1, prelu(z / (float)NH, a4), sW + HD_TA_WP1 + l * 32);
        }
        float o = sW[HD_TA_BP2];
#pragma unroll
        for (int k = 0; k < 30; ++k) o = fmaf(prelu(p1[k], a5), sW[HD_TA_WP2 + k], o);
        out[t] = o;
    }
}

__device__ __forceinline__ void load_heads(float* sW, float* sA, const float* __restrict__ packed,
                                           const float* __restrict__ fold, int T) {
    for (int i = threadIdx.x; i < HD_FLOATS / 4; i += HEADS_THREADS)
        reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(packed)[i];
    const int nf = T * NH * 32 + T * NH;
    for (int i = threadIdx.x; i < nf; i += HEADS_THREADS) sA[i] = fold[i];
    __syncthreads();
}

// y[g, t] for every grid node: SpatialDirect -> TemporalAttention.
__global__ void __launch_bounds__(HEADS_THREADS)
    heads_grid_kernel(const float* __restrict__ packed, const float* __restrict__ fold, int T,
                      const float* __restrict__ x_spatial, int ld_x, int G, float* __restrict__ y) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sA = smem + HD_FLOATS;
    load_heads(sW, sA, packed, fold, T);
    const int g = blockIdx.x * HEADS_THREADS + threadIdx.x;
    if (g >= G) return;
    float yl[30];
    {
        float x[30];
#pragma unroll
        for (int k = 0; k < 30; ++k) x[k] = __ldg(x_spatial + (int64_t)g * ld_x + k);
#pragma unroll
        for (int o = 0; o < 30; ++o) yl[o] = sW[HD_SD_B + o];
#pragma unroll
        for (int k = 0; k < 30; ++k) fma_row<30>(yl, x[k], sW + HD_SD_W + k * 32);
        const float a = sW[HD_SD_SL];
#pragma unroll
        for (int o = 0; o < 30; ++o) yl[o] = prelu(yl[o], a);                     // module.py:258-260
    }
    temporal_attention_node(sW, sA, sA + T * NH * 32, T, yl, y + (int64_t)g * T);
}

// x[q, t] for every query point: SpatialAttention over its k nearest context nodes -> TemporalAttention.
__global__ void __launch_bounds__(HEADS_THREADS)
    heads_query_kernel(const float* __restrict__ packed, const float* __restrict__ fold, int T,
                       const float* __restrict__ x_spatial, int ld_x, const float* __restrict__ x_context,
                       const float* __restrict__ x_query, const int64_t* __restrict__ nbr, int k_nbr, int Q,
                       float scale_rel, float* __restrict__ x_out) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sA = smem + HD_FLOATS;
    load_heads(sW, sA, packed, fold, T);
    const int qi = blockIdx.x * HEADS_THREADS + threadIdx.x;
    if (qi >= Q) return;
    const float sqrt_l = sqrtf((float)NL);
    const float a1 = sW[HD_SA_SL], a2 = sW[HD_SA_SL + 1];
    const float qx = __ldg(x_query + (int64_t)qi * 3), qy = __ldg(x_query + (int64_t)qi * 3 + 1),
                qz = __ldg(x_query + (int64_t)qi * 3 + 2);
    constexpr int KMAX = 16;
    float alpha[KMAX][NH];
    // ---- pass 1: attention scores of the k edges (module.py:288-292) ------------------------------------------------------
#pragma unroll 1
    for (int e = 0; e < k_nbr; ++e) {
        const int64_t j = __ldg(nbr + (int64_t)qi * k_nbr + e);
        const float ea[3] = {(qx - __ldg(x_context + j * 3)) / scale_rel, (qy - __ldg(x_context + j * 3 + 1)) / scale_rel,
                             (qz - __ldg(x_context + j * 3 + 2)) / scale_rel};
        float xj[30];
#pragma unroll
        for (int c = 0; c < 30; ++c) xj[c] = __ldg(x_spatial + j * ld_x + c);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float qv[NL], cv[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                qv[l] = sW[HD_SA_BQ + h * NL + l];
                cv[l] = sW[HD_SA_BC + h * NL + l];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    qv[l] = fmaf(ea[c], sW[HD_SA_WQ + c * 76 + h * NL + l], qv[l]);
                    cv[l] = fmaf(ea[c], sW[HD_SA_WC + (30 + c) * 76 + h * NL + l], cv[l]);
                }
            }
#pragma unroll
            for (int c = 0; c < 30; ++c) {
#pragma unroll
                for (int l = 0; l < NL; ++l) cv[l] = fmaf(xj[c], sW[HD_SA_WC + c * 76 + h * NL + l], cv[l]);
            }
            float d = 0.f;
#pragma unroll
            for (int l = 0; l < NL; ++l) d = fmaf(qv[l], cv[l], d);
            if (e < KMAX) alpha[e][h] = prelu(d / sqrt_l, a1);
        }
    }
    // ---- softmax over the query's edges, per head (PyG softmax: max-shifted) ---------------------------------------------
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        float m = -INFINITY;
        for (int e = 0; e < k_nbr; ++e) m = fmaxf(m, alpha[e][h]);
        float s = 0.f;
        for (int e = 0; e < k_nbr; ++e) {
            alpha[e][h] = expf(alpha[e][h] - m);
            s += alpha[e][h];
        }
        for (int e = 0; e < k_nbr; ++e) alpha[e][h] /= (s + 1e-16f);
    }
    // ---- pass 2: weighted values, mean over heads, projection (module.py:293-297) ------------------------------------------
    float out[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) out[l] = 0.f;
#pragma unroll 1
    for (int e = 0; e < k_nbr; ++e) {
        const int64_t j = __ldg(nbr + (int64_t)qi * k_nbr + e);
        const float ea[3] = {(qx - __ldg(x_context + j * 3)) / scale_rel, (qy - __ldg(x_context + j * 3 + 1)) / scale_rel,
                             (qz - __ldg(x_context + j * 3 + 2)) / scale_rel};
        float xj[30];
#pragma unroll
        for (int c = 0; c < 30; ++c) xj[c] = __ldg(x_spatial + j * ld_x + c);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float vv[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) vv[l] = sW[HD_SA_BV + h * NL + l];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int l = 0; l < NL; ++l) vv[l] = fmaf(ea[c], sW[HD_SA_WV + (30 + c) * 76 + h * NL + l], vv[l]);
            }
#pragma unroll
            for (int c = 0; c < 30; ++c) {
#pragma unroll
                for (int l = 0; l < NL; ++l) vv[l] = fmaf(xj[c], sW[HD_SA_WV + c * 76 + h * NL + l], vv[l]);
            }
            const float w = alpha[e][h];
#pragma unroll
            for (int l = 0; l < NL; ++l) out[l] = fmaf(w, vv[l], out[l]);
        }
    }
    float xq[30];
#pragma unroll
    for (int o = 0; o < 30; ++o) xq[o] = sW[HD_SA_BP + o];
#pragma unroll
    for (int l = 0; l < NL; ++l) fma_row<30>(xq, out[l] / (float)NH, sW + HD_SA_WP + l * 32);
#pragma unroll
    for (int o = 0; o < 30; ++o) xq[o] = prelu(xq[o], a2);
    temporal_attention_node(sW, sA, sA + T * NH * 32, T, xq, x_out + (int64_t)qi * T);
}

size_t heads_smem_bytes(int T) { return sizeof(float) * (size_t)(HD_FLOATS + T * NH * 32 + T * NH + 4); }

}  // namespace

int launch_heads_grid(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, int G, float* y,
                      cudaStream_t st) {
    if (G == 0 || T == 0) return GENIE_OK;
    const size_t smem = heads_smem_bytes(T);
    if (smem > 200 * 1024) {
        set_error("heads: too many query times for the shared-memory fold table");
        return GENIE_ERR_UNSUPPORTED;
    }
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(heads_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch tl(KID_HEADS_GRID, st);
    heads_grid_kernel<<<(G + HEADS_THREADS - 1) / HEADS_THREADS, HEADS_THREADS, smem, st>>>(packed, fold, T, x_spatial, ld_x, G,
                                                                                           y);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_heads_query(const float* packed, const float* fold, int T, const float* x_spatial, int ld_x, const float* x_context,
                       const float* x_query, const int64_t* nbr, int k_nbr, int Q, float scale_rel, float* x_out,
                       cudaStream_t st) {
    if (Q == 0 || T == 0) return GENIE_OK;
    if (k_nbr < 1 || k_nbr > 16) {
        set_error("heads: the query read-out supports 1..16 context neighbours per query");
        return GENIE_ERR_UNSUPPORTED;
    }
    const size_t smem = heads_smem_bytes(T);
    if (smem > 200 * 1024) {
        set_error("heads: too many query times for the shared-memory fold table");
        return GENIE_ERR_UNSUPPORTED;
    }
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(heads_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch tl(KID_HEADS_QUERY, st);
    heads_query_kernel<<<(Q + HEADS_THREADS - 1) / HEADS_THREADS, HEADS_THREADS, smem, st>>>(
        packed, fold, T, x_spatial, ld_x, x_context, x_query, nbr, k_nbr, Q, scale_rel, x_out);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
