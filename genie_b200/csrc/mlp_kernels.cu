// Per-product-node dense layers of the TRAINING path (BASELINE.json configs[2]; train_GENIE_model.py:1786 `mz(*input_tensors)`
// with gradients): every `activate(Linear(torch.cat((a, b, c, ...), dim=1)))` of DataAggregation (module.py:87-96),
// DataAggregationAssociationPhase (:387-403), BipartiteGraphOperator.fc1 (:227) and BipartiteGraphReadOutOperator (:349-351),
// forward and backward, as ONE kernel each:
//
//   forward   y = PReLU_a(W [x_0 | x_1 | ... ] + b)                  the concatenation is never materialised
//   backward  g = gy * PReLU_a'(y);  gx_p = g W[:, cols of part p];  gW = g^T [x_0 | x_1 | ...];  gb = sum g;
//             ga = sum_{z < 0} gy * y / a                             (y = a z for z < 0; the SIGN of z is kept as one bit per
//                                                                      output in a uint32 per node — a slope may be negative)
//
// The reference runs these as cat + addmm + prelu (+ their autograd nodes): five passes over [P, 64..95] tensors per layer
// and, in the backward, a weight-gradient GEMM with a 30 x 95 result reduced over P = 10^5..10^7 rows that the library maps
// to one SM (profiles/r3j_train_profile.log: library GEMMs 50 % + cat / prelu / index 20 % of a training sample).  Here a
// CTA stages a tile of 128 nodes (inputs transposed in shared memory), every thread owns a 4 x 4 register tile (4 nodes x 4
// outputs: 16 FMAs per two 16-byte shared-memory loads, the weight load a warp-wide broadcast), and the weight gradient is
// accumulated in registers across all tiles of a persistent CTA (each thread owns a 4 x 4 block of gW) and written once as a
// per-CTA partial: the caller sums the partials (fixed order: bit-reproducible, no atomics).
#include "common.cuh"

namespace {

constexpr int TMN = 128;           // nodes per tile
constexpr int LDT = TMN + 4;       // row stride of the transposed tiles (floats): 16-byte aligned rows, and rows r, r + 1, ...
                                   // start 33 sixteen-byte units apart, so quarter-warps that read 8 consecutive rows at the same
                                   // column hit 8 different bank groups
constexpr int MLP_THREADS = 256;
constexpr int MAX_IN = GENIE_MLP_MAX_IN, MAX_OUT = GENIE_MLP_MAX_OUT;
constexpr int OB = 4, KB = 4;      // per-thread block of the weight gradient: 4 outputs x 4 (strided) input columns

struct MlpArgs {
    int64_t n_rows;
    int n_parts, n_in, n_out;
    int width[4], ld[4], off[4];   // off: first input column of a part
    const float* x[4];
    const float* weight;
    const float* bias;
    const float* slope;
};

// Global rows [n][w] (row stride ld) <-> transposed shared-memory tile T[col0 + k][n].  A warp instruction covers 8 consecutive
// columns of 4 consecutive rows: 32-byte global segments, and shared-memory addresses (k LDT + n) whose banks 4 k + n are all
// different (LDT = 132) — conflict free in both directions.  A warp owns 16 rows of the tile; four independent accesses per
// step, no integer division.
__device__ __forceinline__ void tile_in(const float* __restrict__ xp, int ld, int w, int col0, int64_t i0, int64_t n_rows,
                                        float* T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kl = lane & 7, nl = lane >> 3;
    for (int k0 = 0; k0 < w; k0 += 8) {
        const int k = k0 + kl;
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (warp * 4 + u) * 4 + nl;
            v[u] = (k < w && i < n_rows) ? __ldg(xp + i * ld + k) : 0.f;
        }
        if (k < w) {
#pragma unroll
            for (int u = 0; u < 4; ++u) T[(col0 + k) * LDT + (warp * 4 + u) * 4 + nl] = v[u];
        }
    }
}
__device__ __forceinline__ void tile_out(float* __restrict__ yp, int ld, int w, int col0, int64_t i0, int64_t n_rows,
                                         const float* T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kl = lane & 7, nl = lane >> 3;
    for (int k0 = 0; k0 < w; k0 += 8) {
        const int k = k0 + kl;
        if (k >= w) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = (warp * 4 + u) * 4 + nl;
            if (i0 + n < n_rows) yp[(i0 + n) * ld + k] = T[(col0 + k) * LDT + n];
        }
    }
}
__device__ __forceinline__ void load_x_tile(const MlpArgs& a, int64_t i0, float* Xs) {
    for (int p = 0; p < a.n_parts; ++p) tile_in(a.x[p], a.ld[p], a.width[p], a.off[p], i0, a.n_rows, Xs);
}

// Register tile of both kernels' node-side products: a thread owns 4 consecutive nodes (one LDS.128 of a transposed row)
// times 4 consecutive outputs / input columns (one LDS.128 of a weight row, the same address for the whole warp: broadcast):
// 16 FMAs per two 16-byte shared-memory loads.
__device__ __forceinline__ void fma4x4(float (&acc)[4][4], const float4& xv, const float4& wv) {
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ws[j], acc[i][j]);
}

__global__ void __launch_bounds__(MLP_THREADS) node_mlp_fwd_kernel(const MlpArgs a, float* __restrict__ y, int ld_y,
                                                                   uint32_t* __restrict__ neg_mask) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                                  // [n_in][32]  (K-major: row k = the weights that multiply input k)
    float* Xs = Ws + MAX_IN * 32;                    // [n_in][LDT]; re-used as the output tile [32][LDT]
    uint32_t* Ms = reinterpret_cast<uint32_t*>(Xs + MAX_IN * LDT);      // [TMN] bit o = pre-activation of output o is negative
    for (int idx = threadIdx.x; idx < a.n_in * 32; idx += MLP_THREADS) {
        const int k = idx >> 5, o = idx & 31;
        Ws[idx] = o < a.n_out ? a.weight[o * a.n_in + k] : 0.f;
    }
    const float slope = a.slope ? *a.slope : 1.f;
    const int64_t n_tiles = (a.n_rows + TMN - 1) / TMN;
    const int nq = threadIdx.x & 31, oq = threadIdx.x >> 5;        // nodes 4 nq .. 4 nq + 3, outputs 4 oq .. 4 oq + 3
    float bias4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias4[j] = (a.bias && 4 * oq + j < a.n_out) ? a.bias[4 * oq + j] : 0.f;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t i0 = t * TMN;
        __syncthreads();                              // previous tile's output reads are done
        load_x_tile(a, i0, Xs);
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = bias4[j];
#pragma unroll 4
        for (int k = 0; k < a.n_in; ++k)
            fma4x4(acc, *reinterpret_cast<const float4*>(Xs + k * LDT + 4 * nq), *reinterpret_cast<const float4*>(Ws + k * 32 + 4 * oq));
        if (neg_mask != nullptr && threadIdx.x < TMN) Ms[threadIdx.x] = 0u;
        __syncthreads();                              // all reads of the input tile are done: overwrite it with the output tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 v = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            if (a.slope) v = make_float4(prelu(v.x, slope), prelu(v.y, slope), prelu(v.z, slope), prelu(v.w, slope));
            *reinterpret_cast<float4*>(Xs + (4 * oq + j) * LDT + 4 * nq) = v;
        }
        if (neg_mask != nullptr) {
            // the backward pass needs the SIGN of the pre-activation (a PReLU slope may be negative: the output's sign is not it)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t bits = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) bits |= (acc[i][j] < 0.f ? 1u : 0u) << (4 * oq + j);
                if (bits) atomicOr(Ms + 4 * nq + i, bits);
            }
        }
        __syncthreads();
        tile_out(y, ld_y, a.n_out, 0, i0, a.n_rows, Xs);
        if (neg_mask != nullptr && threadIdx.x < TMN && i0 + threadIdx.x < a.n_rows) neg_mask[i0 + threadIdx.x] = Ms[threadIdx.x];
    }
}

// partial layout per CTA: gW [n_out][n_in] | gb [n_out] | ga [1]
__global__ void __launch_bounds__(MLP_THREADS)
    node_mlp_bwd_kernel(const MlpArgs a, const float* __restrict__ y, int ld_y, const uint32_t* __restrict__ neg_mask,
                        const float* __restrict__ gy, int ld_gy,
                        float* gx0, float* gx1, float* gx2, float* gx3, int ldg0, int ldg1, int ldg2, int ldg3,
                        float* __restrict__ partial, int partial_ld) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                                  // [32 rows][MAX_IN]  row o = W[o][:]  (for gx = g W: broadcast LDS.128 over k)
    float* Xs = Ws + 32 * MAX_IN;                    // [MAX_IN][LDT]; after the weight-gradient phase re-used for the gx tile
    float* Gs = Xs + MAX_IN * LDT;                   // [32][LDT]   g = gy * PReLU'(y), zero rows beyond n_out
    float* red = Gs + 32 * LDT;                      // [MLP_THREADS / 32] slope-gradient partials
    for (int idx = threadIdx.x; idx < 32 * MAX_IN; idx += MLP_THREADS) {
        const int o = idx / MAX_IN, k = idx - o * MAX_IN;
        Ws[idx] = (o < a.n_out && k < a.n_in) ? a.weight[o * a.n_in + k] : 0.f;
    }
    for (int idx = threadIdx.x; idx < 32 * LDT; idx += MLP_THREADS) Gs[idx] = 0.f;
    for (int idx = threadIdx.x; idx < MAX_IN * LDT; idx += MLP_THREADS) Xs[idx] = 0.f;
    const float slope = a.slope ? *a.slope : 1.f;
    const float inv_slope = (a.slope && slope != 0.f) ? 1.f / slope : 0.f;
    float* gxs[4] = {gx0, gx1, gx2, gx3};
    const int ldg[4] = {ldg0, ldg1, ldg2, ldg3};
    // weight gradient: thread = (output block ob: rows 4 ob .. 4 ob + 3) x (input columns kb, kb + KBLK, kb + 2 KBLK, kb + 3 KBLK —
    // strided, so that the lanes of a quarter-warp read 8 CONSECUTIVE rows of the input tile: conflict free)
    const int KBLK = (a.n_in + KB - 1) / KB;
    const int ob = threadIdx.x / KBLK, kb = threadIdx.x - ob * KBLK;
    const bool has_block = ob * OB < a.n_out;          // ceil(n_out / 4) * ceil(n_in / 4) <= 256 is checked by the launcher
    float gw[OB][KB];
#pragma unroll
    for (int i = 0; i < OB; ++i)
#pragma unroll
        for (int j = 0; j < KB; ++j) gw[i][j] = 0.f;
    float gb = 0.f, ga = 0.f;
    const int64_t n_tiles = (a.n_rows + TMN - 1) / TMN;
    const int nq = threadIdx.x & 31, kq0 = threadIdx.x >> 5;
    const int n_kq = (a.n_in + 3) / 4;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t i0 = t * TMN;
        __syncthreads();
        load_x_tile(a, i0, Xs);
        {   // g tile (and the slope gradient): the access pattern of tile_in
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int kl = lane & 7, nl = lane >> 3;
            for (int o0 = 0; o0 < a.n_out; o0 += 8) {
                const int o = o0 + kl;
                float g[4], yv[4];
                bool neg[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + (warp * 4 + u) * 4 + nl;
                    const bool ok = o < a.n_out && i < a.n_rows;
                    g[u] = ok ? __ldg(gy + i * ld_gy + o) : 0.f;
                    yv[u] = (ok && a.slope) ? __ldg(y + i * ld_y + o) : 0.f;
                    neg[u] = (ok && a.slope) ? ((__ldg(neg_mask + i) >> o) & 1u) != 0u : false;
                }
                if (o < a.n_out) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (neg[u]) {          // pre-activation z < 0: y = a z, dy/dz = a, dy/da = z = y / a
                            ga = fmaf(g[u], yv[u] * inv_slope, ga);
                            g[u] *= slope;
                        }
                        Gs[o * LDT + (warp * 4 + u) * 4 + nl] = g[u];
                    }
                }
            }
        }
        __syncthreads();
        // ---- gW += g^T x over the tile's nodes (four nodes per step: LDS.128 along the node axis); gb += sum g ---------------
        if (has_block) {
            const float* g0 = Gs + (ob * OB) * LDT;
#pragma unroll 2
            for (int r = 0; r < TMN; r += 4) {
                float4 gv[OB], xv[KB];
#pragma unroll
                for (int i = 0; i < OB; ++i) gv[i] = *reinterpret_cast<const float4*>(g0 + i * LDT + r);
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const int k = kb + j * KBLK;                   // rows >= n_in of the tile are zero
                    xv[j] = *reinterpret_cast<const float4*>(Xs + (k < MAX_IN ? k : 0) * LDT + r);
                }
#pragma unroll
                for (int i = 0; i < OB; ++i)
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        gw[i][j] += gv[i].x * xv[j].x + gv[i].y * xv[j].y + gv[i].z * xv[j].z + gv[i].w * xv[j].w;
            }
        }
        if (threadIdx.x < a.n_out) {
            const float4* gr = reinterpret_cast<const float4*>(Gs + threadIdx.x * LDT);
            float s = 0.f;
            for (int r = 0; r < TMN / 4; ++r) {
                const float4 v = gr[r];
                s += (v.x + v.y) + (v.z + v.w);
            }
            gb += s;
        }
        __syncthreads();                              // the input tile is dead: its area takes the gx tile
        // ---- gx = g W: thread = 4 nodes x 4 input columns (kq0, kq0 + 8, ...), one broadcast LDS.128 of W[o][4 kq ..] per o ------
        for (int kq = kq0; kq < n_kq; kq += MLP_THREADS / 32) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int o = 0; o < a.n_out; ++o)
                fma4x4(acc, *reinterpret_cast<const float4*>(Gs + o * LDT + 4 * nq), *reinterpret_cast<const float4*>(Ws + o * MAX_IN + 4 * kq));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(Xs + (4 * kq + j) * LDT + 4 * nq) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        }
        __syncthreads();
        for (int p = 0; p < a.n_parts; ++p)          // the rows of every wanted part
            if (gxs[p] != nullptr) tile_out(gxs[p], ldg[p], a.width[p], a.off[p], i0, a.n_rows, Xs);
    }
    // ---- per-CTA partials ---------------------------------------------------------------------------------------------------
    float* out = partial + (int64_t)blockIdx.x * partial_ld;
    if (has_block) {
#pragma unroll
        for (int i = 0; i < OB; ++i)
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int o = ob * OB + i, k = kb + j * KBLK;
                if (o < a.n_out && k < a.n_in) out[o * a.n_in + k] = gw[i][j];
            }
    }
    if (threadIdx.x < a.n_out) out[a.n_out * a.n_in + threadIdx.x] = gb;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) ga += __shfl_xor_sync(FULL_MASK, ga, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ga;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < MLP_THREADS / 32; ++w) s += red[w];
        out[a.n_out * a.n_in + a.n_out] = s;
    }
}

int fill_args(const genie_mlp_desc_t* d, MlpArgs& a) {
    if (!d || d->n_rows < 0 || d->n_parts < 1 || d->n_parts > 4 || d->n_out < 1 || d->n_out > MAX_OUT || !d->weight) return 1;
    a.n_rows = d->n_rows;
    a.n_parts = d->n_parts;
    a.n_out = d->n_out;
    int off = 0;
    for (int p = 0; p < 4; ++p) {
        a.width[p] = a.ld[p] = a.off[p] = 0;
        a.x[p] = nullptr;
        if (p < d->n_parts) {
            if (d->width[p] < 1 || d->ld[p] < d->width[p] || (d->n_rows > 0 && !d->x[p])) return 1;
            a.width[p] = d->width[p];
            a.ld[p] = d->ld[p];
            a.off[p] = off;
            a.x[p] = d->x[p];
            off += d->width[p];
        }
    }
    if (off > MAX_IN) return 1;
    a.n_in = off;
    a.weight = d->weight;
    a.bias = d->bias;
    a.slope = d->slope;
    return 0;
}

constexpr size_t FWD_SMEM = (size_t)(MAX_IN * 32 + MAX_IN * LDT + TMN) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(32 * MAX_IN + MAX_IN * LDT + 32 * LDT + MLP_THREADS / 32) * sizeof(float);

}  // namespace

int mlp_partial_rows(int sm_count) { return 2 * sm_count; }

int launch_node_mlp_fwd(const genie_mlp_desc_t* d, float* y, int ld_y, uint32_t* neg_mask, int sm_count, cudaStream_t st) {
    MlpArgs a;
    if (fill_args(d, a) || !y || ld_y < d->n_out) {
        set_error("genie_node_mlp_fwd: bad descriptor (1-4 parts, n_in <= 104, n_out <= 32)");
        return GENIE_ERR_INVALID;
    }
    if (a.n_rows == 0) return GENIE_OK;
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(node_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(node_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        attr_set.mark();
    }
    const int64_t n_tiles = (a.n_rows + TMN - 1) / TMN;
    const int64_t cap = (int64_t)sm_count * 3;
    TimedLaunch tl(KID_NODE_MLP_FWD, st);
    node_mlp_fwd_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), MLP_THREADS, FWD_SMEM, st>>>(a, y, ld_y,
                                                                                                  d->slope ? neg_mask : nullptr);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_node_mlp_bwd(const genie_mlp_desc_t* d, const float* y, int ld_y, const uint32_t* neg_mask, const float* gy,
                        int ld_gy, float* const* gx,
                        const int* ld_gx, float* partial, int sm_count, cudaStream_t st) {
    MlpArgs a;
    if (fill_args(d, a) || !gy || ld_gy < d->n_out || !partial || (d->slope && (!y || ld_y < d->n_out || !neg_mask))) {
        set_error("genie_node_mlp_bwd: bad argument");
        return GENIE_ERR_INVALID;
    }
    if (((a.n_out + OB - 1) / OB) * ((a.n_in + KB - 1) / KB) > MLP_THREADS) {
        set_error("genie_node_mlp_bwd: layer too large for the register-blocked weight gradient");
        return GENIE_ERR_UNSUPPORTED;
    }
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(node_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(node_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        attr_set.mark();
    }
    const int rows = mlp_partial_rows(sm_count);
    const int pld = a.n_out * a.n_in + a.n_out + 1;
    TimedLaunch tl(KID_NODE_MLP_BWD, st);
    node_mlp_bwd_kernel<<<rows, MLP_THREADS, BWD_SMEM, st>>>(a, y, ld_y, neg_mask, gy, ld_gy, gx ? gx[0] : nullptr, gx ? gx[1] : nullptr,
                                                             gx ? gx[2] : nullptr, gx ? gx[3] : nullptr, ld_gx ? ld_gx[0] : 0,
                                                             ld_gx ? ld_gx[1] : 0, ld_gx ? ld_gx[2] : 0, ld_gx ? ld_gx[3] : 0,
                                                             partial, pld);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
