// Per-product-node dense layers of the TRAINING path (BASELINE.json configs[2]; train_GENIE_model.py:1786 `mz(*input_tensors)`
// with gradients): every `activate(Linear(torch.cat((a, b, c, ...), dim=1)))` of DataAggregation (module.py:87-96),
// DataAggregationAssociationPhase (:387-403), BipartiteGraphOperator.fc1 (:227) and BipartiteGraphReadOutOperator (:349-351),
// forward and backward, as ONE kernel each:
//
//   forward   y = PReLU_a(W [x_0 | x_1 | ... ] + b)                  the concatenation is never materialised
//   backward  g = gy * PReLU_a'(y);  gx_p = g W[:, cols of part p];  gW = g^T [x_0 | x_1 | ...];  gb = sum g;
//             ga = sum gy * min(y, 0) / a                             (y = a z for z < 0, so dz/da-term z = y / a)
//
// The reference runs these as cat + addmm + prelu (+ their autograd nodes): five passes over [P, 64..95] tensors per layer
// and, in the backward, a weight-gradient GEMM with a 30 x 95 result reduced over P = 10^5..10^7 rows that the library maps
// to one SM (profiles/r3j_train_profile.log: library GEMMs 50 % + cat / prelu / index 20 % of a training sample).  Here a
// CTA stages a tile of 128 nodes (inputs transposed in shared memory: conflict-free for the thread-per-node matrix-vector
// products AND for the tile-level g^T x product), the weights are broadcast from shared memory, and the weight gradient is
// accumulated in registers across all tiles of a persistent CTA (each thread owns a 3 x 4 block of gW) and written once as a
// per-CTA partial: the caller sums the partials (fixed order: bit-reproducible, no atomics).
#include "common.cuh"

namespace {

constexpr int TMN = 128;           // nodes per tile
constexpr int LDT = TMN + 1;       // row stride of the transposed tiles (floats)
constexpr int MLP_THREADS = 256;
constexpr int MAX_IN = GENIE_MLP_MAX_IN, MAX_OUT = GENIE_MLP_MAX_OUT;
constexpr int OB = 3, KB = 4;      // per-thread block of the weight gradient

struct MlpArgs {
    int64_t n_rows;
    int n_parts, n_in, n_out;
    int width[4], ld[4], off[4];   // off: first input column of a part
    const float* x[4];
    const float* weight;
    const float* bias;
    const float* slope;
};

// X tile -> Xs[k][n] (transposed), coalesced reads of every part's rows
__device__ __forceinline__ void load_x_tile(const MlpArgs& a, int64_t i0, float* Xs) {
    for (int p = 0; p < a.n_parts; ++p) {
        const int w = a.width[p];
        const float* __restrict__ xp = a.x[p];
        const int total = TMN * w;
        for (int idx = threadIdx.x; idx < total; idx += MLP_THREADS) {
            const int n = idx / w, k = idx - n * w;
            const int64_t i = i0 + n;
            Xs[(a.off[p] + k) * LDT + n] = i < a.n_rows ? xp[i * a.ld[p] + k] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(MLP_THREADS) node_mlp_fwd_kernel(const MlpArgs a, float* __restrict__ y, int ld_y) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                                  // [n_in][32]  (K-major: row k = the weights that multiply input k)
    float* Xs = Ws + MAX_IN * 32;                    // [n_in][LDT]; re-used as the output tile [n_out][LDT]
    for (int idx = threadIdx.x; idx < a.n_in * 32; idx += MLP_THREADS) {
        const int k = idx >> 5, o = idx & 31;
        Ws[idx] = o < a.n_out ? a.weight[o * a.n_in + k] : 0.f;
    }
    const float slope = a.slope ? *a.slope : 1.f;
    const int64_t n_tiles = (a.n_rows + TMN - 1) / TMN;
    const int n = threadIdx.x & (TMN - 1), half = threadIdx.x >> 7;       // two threads per node: outputs [0,16) / [16,32)
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t i0 = t * TMN;
        __syncthreads();                              // previous tile's output reads are done
        load_x_tile(a, i0, Xs);
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[o] = (a.bias && half * 16 + o < a.n_out) ? a.bias[half * 16 + o] : 0.f;
        for (int k = 0; k < a.n_in; ++k) fma_row16(acc, Xs[k * LDT + n], Ws + k * 32 + half * 16);
        __syncthreads();                              // all reads of the input tile are done: overwrite it with the output tile
#pragma unroll
        for (int o = 0; o < 16; ++o) Xs[(half * 16 + o) * LDT + n] = a.slope ? prelu(acc[o], slope) : acc[o];
        __syncthreads();
        const int total = TMN * a.n_out;
        for (int idx = threadIdx.x; idx < total; idx += MLP_THREADS) {
            const int r = idx / a.n_out, o = idx - r * a.n_out;
            if (i0 + r < a.n_rows) y[(i0 + r) * ld_y + o] = Xs[o * LDT + r];
        }
    }
}

// partial layout per CTA: gW [n_out][n_in] | gb [n_out] | ga [1]
__global__ void __launch_bounds__(MLP_THREADS)
    node_mlp_bwd_kernel(const MlpArgs a, const float* __restrict__ y, int ld_y, const float* __restrict__ gy, int ld_gy,
                        float* gx0, float* gx1, float* gx2, float* gx3, int ldg0, int ldg1, int ldg2, int ldg3,
                        float* __restrict__ partial, int partial_ld) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                                  // [n_out = 32 rows][MAX_IN]  row o = W[o][:]  (N-major for gx = g W)
    float* Xs = Ws + 32 * MAX_IN;                    // [n_in][LDT]
    float* Gs = Xs + MAX_IN * LDT;                   // [32][LDT]   g = gy * PReLU'(y), zero rows beyond n_out
    float* red = Gs + 32 * LDT;                      // [MLP_THREADS / 32] slope-gradient partials
    for (int idx = threadIdx.x; idx < 32 * MAX_IN; idx += MLP_THREADS) {
        const int o = idx / MAX_IN, k = idx - o * MAX_IN;
        Ws[idx] = (o < a.n_out && k < a.n_in) ? a.weight[o * a.n_in + k] : 0.f;
    }
    for (int idx = threadIdx.x; idx < 32 * LDT; idx += MLP_THREADS) Gs[idx] = 0.f;
    const float slope = a.slope ? *a.slope : 1.f;
    const float inv_slope = (a.slope && slope != 0.f) ? 1.f / slope : 0.f;
    float* gxs[4] = {gx0, gx1, gx2, gx3};
    const int ldg[4] = {ldg0, ldg1, ldg2, ldg3};
    // this thread's block of the weight gradient: rows o0 .. o0+OB-1, columns k0 .. k0+KB-1 (KBLK blocks along k)
    const int KBLK = (a.n_in + KB - 1) / KB;
    const int ob = threadIdx.x / KBLK, kb = threadIdx.x - ob * KBLK;
    const bool has_block = ob * OB < a.n_out;          // (n_out / 3) * ceil(n_in / 4) <= 256 is checked by the launcher
    const int o0 = ob * OB, k0 = kb * KB;
    float gw[OB][KB];
#pragma unroll
    for (int i = 0; i < OB; ++i)
#pragma unroll
        for (int j = 0; j < KB; ++j) gw[i][j] = 0.f;
    float gb = 0.f, ga = 0.f;
    const int64_t n_tiles = (a.n_rows + TMN - 1) / TMN;
    const int n = threadIdx.x & (TMN - 1), half = threadIdx.x >> 7;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t i0 = t * TMN;
        __syncthreads();
        load_x_tile(a, i0, Xs);
        {   // g tile (and the slope gradient): coalesced reads of gy / y rows
            const int total = TMN * a.n_out;
            for (int idx = threadIdx.x; idx < total; idx += MLP_THREADS) {
                const int r = idx / a.n_out, o = idx - r * a.n_out;
                float g = 0.f;
                if (i0 + r < a.n_rows) {
                    g = gy[(i0 + r) * ld_gy + o];
                    if (a.slope) {
                        const float yv = y[(i0 + r) * ld_y + o];
                        if (yv < 0.f) {
                            ga = fmaf(g, yv * inv_slope, ga);
                            g *= slope;
                        }
                    }
                }
                Gs[o * LDT + r] = g;
            }
        }
        __syncthreads();
        // ---- gx = g W, thread per (node, half of the input columns of every part) -------------------------------------------
        for (int p = 0; p < a.n_parts; ++p) {
            if (gxs[p] == nullptr) continue;
            const int w = a.width[p];
            const int64_t i = i0 + n;
            for (int k = half; k < w; k += 2) {
                const float* wk = Ws + a.off[p] + k;
                float s = 0.f;
                for (int o = 0; o < a.n_out; ++o) s = fmaf(Gs[o * LDT + n], wk[o * MAX_IN], s);
                if (i < a.n_rows) gxs[p][i * ldg[p] + k] = s;
            }
        }
        // ---- gW += g^T x over the tile's nodes; gb += sum g ------------------------------------------------------------------
        if (has_block) {
            const float* g0 = Gs + o0 * LDT;
            const float* x0 = Xs + k0 * LDT;
#pragma unroll 4
            for (int r = 0; r < TMN; ++r) {
                float gv[OB], xv[KB];
#pragma unroll
                for (int i = 0; i < OB; ++i) gv[i] = g0[i * LDT + r];
#pragma unroll
                for (int j = 0; j < KB; ++j) xv[j] = (k0 + j < a.n_in) ? x0[j * LDT + r] : 0.f;
#pragma unroll
                for (int i = 0; i < OB; ++i)
#pragma unroll
                    for (int j = 0; j < KB; ++j) gw[i][j] = fmaf(gv[i], xv[j], gw[i][j]);
            }
        }
        if (threadIdx.x < a.n_out) {
            const float* gr = Gs + threadIdx.x * LDT;
            float s = 0.f;
            for (int r = 0; r < TMN; ++r) s += gr[r];
            gb += s;
        }
    }
    // ---- per-CTA partials ---------------------------------------------------------------------------------------------------
    float* out = partial + (int64_t)blockIdx.x * partial_ld;
    if (has_block) {
#pragma unroll
        for (int i = 0; i < OB; ++i)
#pragma unroll
            for (int j = 0; j < KB; ++j)
                if (o0 + i < a.n_out && k0 + j < a.n_in) out[(o0 + i) * a.n_in + k0 + j] = gw[i][j];
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

constexpr size_t FWD_SMEM = (size_t)(MAX_IN * 32 + MAX_IN * LDT) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(32 * MAX_IN + MAX_IN * LDT + 32 * LDT + MLP_THREADS / 32) * sizeof(float);

}  // namespace

int mlp_partial_rows(int sm_count) { return 2 * sm_count; }

int launch_node_mlp_fwd(const genie_mlp_desc_t* d, float* y, int ld_y, int sm_count, cudaStream_t st) {
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
    node_mlp_fwd_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), MLP_THREADS, FWD_SMEM, st>>>(a, y, ld_y);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_node_mlp_bwd(const genie_mlp_desc_t* d, const float* y, int ld_y, const float* gy, int ld_gy, float* const* gx,
                        const int* ld_gx, float* partial, int sm_count, cudaStream_t st) {
    MlpArgs a;
    if (fill_args(d, a) || !gy || ld_gy < d->n_out || !partial || (d->slope && (!y || ld_y < d->n_out))) {
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
    node_mlp_bwd_kernel<<<rows, MLP_THREADS, BWD_SMEM, st>>>(a, y, ld_y, gy, ld_gy, gx ? gx[0] : nullptr, gx ? gx[1] : nullptr,
                                                             gx ? gx[2] : nullptr, gx ? gx[3] : nullptr, ld_gx ? ld_gx[0] : 0,
                                                             ld_gx ? ld_gx[1] : 0, ld_gx ? ld_gx[2] : 0, ld_gx ? ld_gx[3] : 0,
                                                             partial, pld);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
