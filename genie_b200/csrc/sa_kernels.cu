// SpatialAggregation forward (module.py:231-249) on the grid-node kNN graph.
//
//   glob = mean over EDGES e of PReLU_3(fglobal x_{j(e)})                      (module.py:249, out-degree weighted)
//   m_e  = PReLU_1(fc1 [x_j ‖ (pos_i - pos_j)/scale_rel ‖ glob])
//   out_i = PReLU_2(fc2 [x_i ‖ mean_e m_e])
// fc1 is split by linearity into a per-node part W_x x_j (computed once per node instead of once per edge), a per-edge
// 3-term position part and a per-call constant W_g glob + b.  Two kernels per layer:
//   sa_pre_kernel   px_j = W_x x_j, and per-CTA partial sums of outdeg_j * PReLU_3(fglobal x_j)   (fixed order)
//   sa_main_kernel  every CTA re-reduces the partials in the same fixed order (deterministic, no atomics), then one warp
//                   per target node: lanes = the 30 hidden channels.
#include "common.cuh"

using namespace gl;

namespace {

constexpr int SA_THREADS = 256;
constexpr int SA_WARPS = SA_THREADS / 32;

__global__ void __launch_bounds__(SA_THREADS)
    sa_pre_kernel(const float* __restrict__ w, const float* __restrict__ x, int ld_x, int C,
                  const int32_t* __restrict__ outdeg, int G, float* __restrict__ px, float* __restrict__ partial) {
    __shared__ float sWx[30 * LD];
    __shared__ float sWgl[30 * 8 + 8];
    __shared__ float red[SA_WARPS][8];
    for (int i = threadIdx.x; i < 30 * LD; i += SA_THREADS) sWx[i] = w[SA_WX + i];
    for (int i = threadIdx.x; i < 30 * 8 + 8; i += SA_THREADS) sWgl[i] = w[SA_WGL + i];
    const float a3 = w[SA_SLOPES + 2];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gsum = 0.f;
    for (int g = blockIdx.x * SA_WARPS + warp; g < G; g += gridDim.x * SA_WARPS) {
        const float xv = lane < C ? x[(int64_t)g * ld_x + lane] : 0.f;
        float acc = 0.f;
        float t = sWgl[30 * 8 + (lane & 7)];
        for (int k = 0; k < C; ++k) {
            const float xk = __shfl_sync(FULL_MASK, xv, k);
            acc = fmaf(xk, sWx[k * LD + lane], acc);
            t = fmaf(xk, sWgl[k * 8 + (lane & 7)], t);
        }
        px[(int64_t)g * 32 + lane] = acc;
        gsum += prelu(t, a3) * (float)outdeg[g];
    }
    if (lane < 8) red[warp][lane] = gsum;
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int wdx = 0; wdx < SA_WARPS; ++wdx) s += red[wdx][threadIdx.x];
        partial[blockIdx.x * 8 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(SA_THREADS)
    sa_main_kernel(const float* __restrict__ w, const float* __restrict__ x, int ld_x, int C,
                   const float* __restrict__ px, const float* __restrict__ pos, float scale_rel,
                   const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int G,
                   const float* __restrict__ partial, int n_partial, float* __restrict__ out, int ld_out) {
    __shared__ float sWpg[8 * LD];
    __shared__ float sW2[60 * LD];
    __shared__ float sB2[LD];
    __shared__ float cg[LD];
    __shared__ float glob[8];
    for (int i = threadIdx.x; i < 8 * LD; i += SA_THREADS) sWpg[i] = w[SA_WPG + i];
    for (int i = threadIdx.x; i < 60 * LD; i += SA_THREADS) sW2[i] = w[SA_W2 + i];
    if (threadIdx.x < LD) sB2[threadIdx.x] = w[SA_B2 + threadIdx.x];
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int b = 0; b < n_partial; ++b) s += partial[b * 8 + threadIdx.x];
        const int64_t n_edges = rowptr[G];
        glob[threadIdx.x] = s / (float)n_edges;
    }
    __syncthreads();
    if (threadIdx.x < LD) {
        float c = w[SA_B1 + threadIdx.x];
#pragma unroll
        for (int q = 0; q < 5; ++q) c = fmaf(glob[q], sWpg[(3 + q) * LD + threadIdx.x], c);
        cg[threadIdx.x] = c;
    }
    __syncthreads();
    const float a1 = w[SA_SLOPES + 0], a2 = w[SA_SLOPES + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_scale = 1.f / scale_rel;
    const float wp0 = sWpg[0 * LD + lane] * inv_scale, wp1 = sWpg[1 * LD + lane] * inv_scale, wp2 = sWpg[2 * LD + lane] * inv_scale;
    const float cgl = cg[lane];
    for (int i = blockIdx.x * SA_WARPS + warp; i < G; i += gridDim.x * SA_WARPS) {
        const float pix = pos[(int64_t)i * 3 + 0];
        const float piy = pos[(int64_t)i * 3 + 1];
        const float piz = pos[(int64_t)i * 3 + 2];
        const int64_t beg = rowptr[i], end = rowptr[i + 1];
        float acc = 0.f;
        for (int64_t e0 = beg; e0 < end; e0 += 32) {
            const int cnt = (int)min((int64_t)32, end - e0);
            const int32_t cj = lane < cnt ? col[e0 + lane] : 0;
            for (int u = 0; u < cnt; ++u) {
                const int64_t j = __shfl_sync(FULL_MASK, cj, u);
                const float dx = pix - __ldg(pos + j * 3 + 0);      // (pos_i - pos_j) / scale_rel: the scale is folded into wp*
                const float dy = piy - __ldg(pos + j * 3 + 1);
                const float dz = piz - __ldg(pos + j * 3 + 2);
                float m = __ldg(px + j * 32 + lane) + cgl;
                m = fmaf(dx, wp0, m);
                m = fmaf(dy, wp1, m);
                m = fmaf(dz, wp2, m);
                acc += prelu(m, a1);
            }
        }
        const float agg = (end > beg) ? acc / (float)(end - beg) : 0.f;
        const float xv = lane < C ? x[(int64_t)i * ld_x + lane] : 0.f;
        float o = sB2[lane];
        for (int k = 0; k < C; ++k) o = fmaf(__shfl_sync(FULL_MASK, xv, k), sW2[k * LD + lane], o);
#pragma unroll 6
        for (int k = 0; k < 30; ++k) o = fmaf(__shfl_sync(FULL_MASK, agg, k), sW2[(30 + k) * LD + lane], o);
        if (lane < 30) out[(int64_t)i * ld_out + lane] = prelu(o, a2);
    }
}

}  // namespace

int launch_spatial_aggregation(const genie_plan* p, const float* packed, int layer, const float* x, int ld_x,
                               const float* pos, float scale_rel, float* px, float* partial, float* out, int ld_out,
                               cudaStream_t st) {
    const int G = p->g.n_grid;
    if (G == 0) return GENIE_OK;
    if (layer < 0 || layer > 2) {
        set_error("spatial aggregation: layer must be 0, 1 or 2");
        return GENIE_ERR_INVALID;
    }
    const int C = layer == 0 ? 15 : 30;
    const float* w = packed + SA_BASE + layer * SA_SIZE;
    int nb = (G + SA_WARPS - 1) / SA_WARPS;
    const int cap = p->sm_count * 2 < 1024 ? p->sm_count * 2 : 1024;
    if (nb > cap) nb = cap;
    {
        TimedLaunch tl(KID_SA_PRE, st);
        sa_pre_kernel<<<nb, SA_THREADS, 0, st>>>(w, x, ld_x, C, p->g.grid_outdeg, G, px, partial);
    }
    GENIE_LAUNCH_CHECK();
    // more, shorter CTAs for the per-node kernel: one warp per target node is latency bound (15 dependent-free row gathers
    // from L2 and ~75 shuffled FMAs per node), so occupancy is what counts
    int nb_main = (G + SA_WARPS - 1) / SA_WARPS;
    if (nb_main > p->sm_count * 8) nb_main = p->sm_count * 8;
    TimedLaunch tl(KID_SA_MAIN, st);
    sa_main_kernel<<<nb_main, SA_THREADS, 0, st>>>(w, x, ld_x, C, px, pos, scale_rel, p->g.grid_rowptr, p->g.grid_col, G,
                                               partial, nb, out, ld_out);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
