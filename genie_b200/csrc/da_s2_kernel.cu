// DataAggregation layer 2, station pass, fused with BipartiteGraphOperator — CARTESIAN graphs with tiling tables.
//
// Reference: module.py:94-98 (second aggregation of DataAggregation) and module.py:224-229 (Bipartite_ReadIn):
//   x_latent = PReLU2([c_a + mean_sta v_a | c_b + mean_src v_b])            (the 15-wide halves of l2_t*_2, split by linearity
//                                                                           in layer 1: zc = [c_a 0 | c_b 0], va, vb)
//   h        = max_c(mask) * PReLU(fc1 [x_latent | edge attr])              33 -> 30, per product node (max_c(mask) was
//                                                                           left in padding channel 15 of zc by layer 1)
//   out[g]   = PReLU(fc2 sum_{s} h[g,s])                                    sum over the stations of grid node g, 30 -> 15
// mean_src v_b comes from the source pass (src_mean_kernels.cu, 16-float rows).  One tile = (grid node g, compact set of
// <= 128 stations): producer warps stage the v_a rows of the tile and of its station halo (64 B each), the tile's zc rows
// and its mean_src rows with cp.async into one of three shared-memory buffers; three compute warpgroups (thread per row,
// one buffer each) gather the <= 16 station neighbours out of shared memory, run the 33 -> 30 linear layer with the
// weights broadcast from shared memory, and reduce over the tile's rows with a register-transposing butterfly.  A CTA owns
// whole grid nodes: the per-tile sums of one grid node meet in shared memory in a fixed order (no atomics: results are
// bit-reproducible), and the last tile to finish applies fc2 and writes the read-in row.
#include "bf16.cuh"
#include "common.cuh"

using namespace gl;

namespace {

constexpr int S2_THREADS = 640;
constexpr int N_WG = 4;                              // compute warpgroups = buffers
constexpr int ROWS = GENIE_TILE_ROWS_MAX;
constexpr int NT_MAX = 32;                           // station tiles per grid node
constexpr int G_SLOTS = 8;                           // grid nodes in flight inside one CTA (a tile can start while the
                                                     // tiles 1, 2, 4, 5 places before it are unfinished: <= 6 nodes)

constexpr int SB_VA = 0;                             // [ROWS + 1][64 B]  v_a rows (tile stations first, then halo)
constexpr int SB_ZC = SB_VA + (ROWS + 1) * 64;       // [128][128 B]      zc rows, 16-byte chunks XOR-swizzled by row
constexpr int SB_M2 = SB_ZC + 128 * 128;             // [128][64 B]       mean_src(v_b) rows, chunks swizzled by (row >> 1)
constexpr int SB_SIZE = (SB_M2 + 128 * 64 + 127) / 128 * 128;
constexpr int W_FLOATS = RI_END - RI_WFC1;
constexpr int SM_W = 0;
constexpr int SM_BUF = (W_FLOATS * 4 + 127) / 128 * 128;
constexpr int SM_PART = SM_BUF + N_WG * SB_SIZE;     // [G_SLOTS][NT_MAX][32] per-tile column sums
constexpr int SM_WPART = SM_PART + G_SLOTS * NT_MAX * 32 * 4;   // [N_WG][2][4][32] per-warp column sums
constexpr int SM_BAR = SM_WPART + N_WG * 2 * 4 * 32 * 4;
constexpr int SM_TOTAL = SM_BAR + 128;
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// fc1 of the read-in (33 x 30 + bias) in the constant bank: FFMA reads its weight operand straight from c[][] (uniform, no
// register-file write-back), which takes the weight broadcasts off the shared-memory pipe (the bound of this kernel in
// profiles/r1k).  Refreshed from the caller's packed weights by a stream-ordered device-to-device copy before every launch.
__constant__ float c_ri_slots[GENIE_CSLOTS][W_FLOATS];      // one slot per plan (genie_plan::cslot)

struct Bars {
    uint64_t full[N_WG], empty[N_WG];
    int count[G_SLOTS];
};
static_assert(sizeof(Bars) <= 128, "barrier block");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();     // protocol bug: fail the launch instead of hanging the GPU
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Column sums over the 32 lanes of a warp: on return lane l holds sum_lanes v[l].  31 shuffles (recursive halving with a
// register transpose) instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = up ? v[i] : v[i + step];
            const float keep = up ? v[i + step] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL_MASK, send, step);
        }
    }
    return v[0];
}

// BF16: the mean_src(v_b) rows `m2` are bf16 (32 bytes; genie_plan_set_storage), staged as they are and widened on read.
// ASSOC: layer 2 of DataAggregationAssociationPhase (module.py:400-402) — the same gather and activation with the phase's slope
// `*a2_assoc`, the rows stored as [o1(15) 0 | o2(15) 0] into `latent_out` (ld 32); no read-in follows.
template <bool STORE_LATENT, int CSLOT, bool BF16, bool ASSOC = false>
__global__ void __launch_bounds__(S2_THREADS, 1)
    da_layer2_s_kernel(const float* __restrict__ packed, const float* __restrict__ zc, const float* __restrict__ va,
                       const float* __restrict__ m2, const float* __restrict__ mask, const float* __restrict__ edge_attr,
                       float* __restrict__ latent_out, float* __restrict__ out, int ld_out, int S, int G, int NT,
                       const int32_t* __restrict__ tile_rows, const int32_t* __restrict__ tile_meta,
                       const uint16_t* __restrict__ tile_nbr, const float* __restrict__ tile_invdeg, const float* __restrict__ a2_assoc) {
    extern __shared__ __align__(128) unsigned char smem[];
    const float* c_ri = c_ri_slots[CSLOT];     // compile-time slot: fc1 stays immediate constant-bank operands
    float* sW = reinterpret_cast<float*>(smem + SM_W);
    float* part = reinterpret_cast<float*>(smem + SM_PART);
    float* wpart = reinterpret_cast<float*>(smem + SM_WPART);
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (!ASSOC)
        for (int i = threadIdx.x; i < W_FLOATS; i += S2_THREADS) sW[i] = packed[RI_WFC1 + i];
    if (threadIdx.x < N_WG * 4) {       // zero rows of the v_a areas
        const int b = threadIdx.x >> 2, c = threadIdx.x & 3;
        *reinterpret_cast<float4*>(smem + SM_BUF + b * SB_SIZE + SB_VA + ROWS * 64 + c * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < N_WG; ++b) {
            mbar_init(&bars->full[b], 128);
            mbar_init(&bars->empty[b], 128);
        }
        for (int s = 0; s < G_SLOTS; ++s) bars->count[s] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const float a2 = ASSOC ? __ldg(a2_assoc) : packed[DA_SLOPES + SL_A2];
    __syncthreads();
    const float ri_a1 = ASSOC ? 0.f : sW[RI_SLOPES - RI_WFC1], ri_a2 = ASSOC ? 0.f : sW[RI_SLOPES - RI_WFC1 + 1];

    // number of grid nodes of this CTA and of its tiles
    const int n_g = blockIdx.x < G ? (G - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_q = (int64_t)n_g * NT;

    if (warp < 4) {
        // ================================ producers: cp.async row gather ==============================================
        // 64-byte rows: thread `tid` owns chunk c4 = tid & 3 of rows r4 + 32 j; 128-byte rows: chunk c8 = tid & 7 of rows
        // r8 + 16 j.  The station ids of all its rows are fetched in one batch before the copies are issued.
        const int tid = threadIdx.x;
        const int r4 = tid >> 2, c4 = tid & 3, r8 = tid >> 3, c8 = tid & 7;
        constexpr int J4 = (ROWS + 31) / 32;
        for (int64_t q = 0; q < n_q; ++q) {
            const int k = (int)(q / NT), T = (int)(q - (int64_t)k * NT);
            const int g = blockIdx.x + k * gridDim.x;
            const int b = (int)(q % N_WG);
            const int64_t n = q / N_WG;
            const int n_own = __ldg(tile_meta + 2 * T), n_rows = __ldg(tile_meta + 2 * T + 1);
            const int32_t* rows = tile_rows + (int64_t)T * ROWS;
            int id4[J4], id8[8];
#pragma unroll
            for (int j = 0; j < J4; ++j) id4[j] = (r4 + 32 * j) < n_rows ? __ldg(rows + r4 + 32 * j) : -1;
#pragma unroll
            for (int j = 0; j < 8; ++j) id8[j] = (r8 + 16 * j) < n_own ? __ldg(rows + r8 + 16 * j) : -1;
            if (n > 0) mbar_wait(&bars->empty[b], (uint32_t)((n - 1) & 1));
            const uint32_t sb = smem_u32(smem + SM_BUF + b * SB_SIZE);
            const int64_t node0 = (int64_t)g * S;
#pragma unroll
            for (int j = 0; j < J4; ++j)
                if (id4[j] >= 0) cp_async16(sb + SB_VA + (r4 + 32 * j) * 64 + c4 * 16, va + (node0 + id4[j]) * LD_V + c4 * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r8 + 16 * j;
                if (id8[j] >= 0)
                    cp_async16(sb + SB_ZC + r * 128 + ((c8 ^ (r & 7)) << 4), zc + (node0 + id8[j]) * LD_ZC + c8 * 4);
            }
            if (BF16) {      // 32-byte rows: chunk c2 = tid & 1 of rows (tid >> 1) + 64 j, swizzled by (row >> 2) & 1
                const int c2 = tid & 1;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = (tid >> 1) + 64 * j;
                    if (r < n_own)
                        cp_async16(sb + SB_M2 + r * 32 + ((c2 ^ ((r >> 2) & 1)) << 4),
                                   reinterpret_cast<const unsigned char*>(m2) + (node0 + __ldg(rows + r)) * 32 + c2 * 16);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r4 + 32 * j;
                    if (r < n_own)
                        cp_async16(sb + SB_M2 + r * 64 + ((c4 ^ ((r >> 1) & 3)) << 4), m2 + (node0 + id4[j]) * LD_V + c4 * 4);
                }
            }
            cp_async_arrive_noinc(&bars->full[b]);
        }
    } else {
        // ================================ compute warpgroups (thread per row) ==========================================
        const int wg = (warp - 4) >> 2, wiw = (warp - 4) & 3;
        const int r = wiw * 32 + lane;
        const int key = lane & 3;
        unsigned char* sb = smem + SM_BUF + wg * SB_SIZE;
        int64_t n = 0;
        for (int64_t q = wg; q < n_q; q += N_WG, ++n) {
            const int k = (int)(q / NT), T = (int)(q - (int64_t)k * NT);
            const int g = blockIdx.x + k * gridDim.x;
            const int slot = k & (G_SLOTS - 1);
            const int n_own = __ldg(tile_meta + 2 * T);
            const bool valid = r < n_own;
            const int64_t node = (int64_t)g * S + (valid ? __ldg(tile_rows + (int64_t)T * ROWS + r) : 0);
            const uint4* nb = reinterpret_cast<const uint4*>(tile_nbr + ((int64_t)T * 128 + r) * 16);
            const uint4 n0 = __ldg(nb), n1 = __ldg(nb + 1);
            const float invdeg = __ldg(tile_invdeg + T * 128 + r);
            float e0 = 0.f, e1 = 0.f, e2 = 0.f;
            if (!ASSOC && valid) {
                e0 = __ldg(edge_attr + node * 3);
                e1 = __ldg(edge_attr + node * 3 + 1);
                e2 = __ldg(edge_attr + node * 3 + 2);
            }
            mbar_wait(&bars->full[wg], (uint32_t)(n & 1));
            // ---- sum of the station neighbours' v_a rows (chunk k ^ key of every row) ----------------------------------------
            float4 acc[4];
            {
                f32x4_t a2[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) a2[c].lo = a2[c].hi = 0ull;
                const uint32_t w[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t idx = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
                    const unsigned char* ra = sb + SB_VA + idx * 64;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        fadd4(a2[c], *reinterpret_cast<const float4*>(ra + ((c ^ key) << 4)));
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = to_float4(a2[c]);
            }
            {   // un-rotate: acc[c] <- chunk c
                const bool s0 = key & 1, s1 = key & 2;
                float4 t0 = s0 ? acc[1] : acc[0], t1 = s0 ? acc[0] : acc[1], t2 = s0 ? acc[3] : acc[2], t3 = s0 ? acc[2] : acc[3];
                acc[0] = s1 ? t2 : t0; acc[2] = s1 ? t0 : t2; acc[1] = s1 ? t3 : t1; acc[3] = s1 ? t1 : t3;
            }
            // ---- x_latent ----------------------------------------------------------------------------------------------------
            float x[32];
            float zmask = 0.f;       // max_c(mask) of the node: layer 1 left it in padding channel 15 of the zc row
            if (valid) {
                float m2v[16];
                if (BF16) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float f[8];
                        bf16_unpack8(*reinterpret_cast<const uint4*>(sb + SB_M2 + r * 32 + ((c ^ ((r >> 2) & 1)) << 4)), f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) m2v[8 * c + e] = f[e];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 t4 = *reinterpret_cast<const float4*>(sb + SB_M2 + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
                        m2v[4 * c] = t4.x; m2v[4 * c + 1] = t4.y; m2v[4 * c + 2] = t4.z; m2v[4 * c + 3] = t4.w;
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 za = *reinterpret_cast<const float4*>(sb + SB_ZC + r * 128 + ((c ^ (r & 7)) << 4));
                    if (c == 3) zmask = za.w;
                    const float4 zb = *reinterpret_cast<const float4*>(sb + SB_ZC + r * 128 + (((c + 4) ^ (r & 7)) << 4));
                    const float4 mb = make_float4(m2v[4 * c], m2v[4 * c + 1], m2v[4 * c + 2], m2v[4 * c + 3]);
                    x[4 * c + 0] = prelu(za.x + acc[c].x * invdeg, a2);
                    x[4 * c + 1] = prelu(za.y + acc[c].y * invdeg, a2);
                    x[4 * c + 2] = prelu(za.z + acc[c].z * invdeg, a2);
                    x[4 * c + 3] = prelu(za.w + acc[c].w * invdeg, a2);
                    x[16 + 4 * c + 0] = prelu(zb.x + mb.x, a2);
                    x[16 + 4 * c + 1] = prelu(zb.y + mb.y, a2);
                    x[16 + 4 * c + 2] = prelu(zb.z + mb.z, a2);
                    x[16 + 4 * c + 3] = prelu(zb.w + mb.w, a2);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = 0.f;
            }
            mbar_arrive(&bars->empty[wg]);           // release: every shared-memory read of this tile has completed
            // x[0..14] = first half, x[16..30] = second half (x[15], x[31] are padding)
            if (ASSOC) {
                if (valid) {
                    x[15] = 0.f;
                    x[31] = 0.f;
                    float4* dst = reinterpret_cast<float4*>(latent_out + node * 32);
#pragma unroll
                    for (int c = 0; c < 8; ++c) __stcs(dst + c, make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
                }
                continue;
            }
            if (STORE_LATENT && valid) {
                float2* dst = reinterpret_cast<float2*>(latent_out + node * 30);       // 120-byte rows: 8-byte aligned
#pragma unroll
                for (int i = 0; i < 7; ++i) dst[i] = make_float2(x[2 * i], x[2 * i + 1]);
                dst[7] = make_float2(x[14], x[16]);
#pragma unroll
                for (int i = 0; i < 7; ++i) dst[8 + i] = make_float2(x[17 + 2 * i], x[18 + 2 * i]);
            }
            // ---- h = max(mask) * PReLU(fc1 [x_latent | attr]) ------------------------------------------------------------------
            float h[32];
            {
                f32x2_t acc2[15];
#pragma unroll
                for (int o = 0; o < 15; ++o) acc2[o] = pack2(c_ri[(RI_BFC1 - RI_WFC1) + 2 * o], c_ri[(RI_BFC1 - RI_WFC1) + 2 * o + 1]);
#pragma unroll
                for (int i = 0; i < 15; ++i) {
#pragma unroll
                    for (int o = 0; o < 15; ++o) ffma2(acc2[o], x[i], c_ri[i * LD + 2 * o], c_ri[i * LD + 2 * o + 1]);
                }
#pragma unroll
                for (int i = 0; i < 15; ++i) {
#pragma unroll
                    for (int o = 0; o < 15; ++o)
                        ffma2(acc2[o], x[16 + i], c_ri[(15 + i) * LD + 2 * o], c_ri[(15 + i) * LD + 2 * o + 1]);
                }
#pragma unroll
                for (int o = 0; o < 15; ++o) {
                    ffma2(acc2[o], e0, c_ri[30 * LD + 2 * o], c_ri[30 * LD + 2 * o + 1]);
                    ffma2(acc2[o], e1, c_ri[31 * LD + 2 * o], c_ri[31 * LD + 2 * o + 1]);
                    ffma2(acc2[o], e2, c_ri[32 * LD + 2 * o], c_ri[32 * LD + 2 * o + 1]);
                }
                float acc30[30];
#pragma unroll
                for (int o = 0; o < 15; ++o) unpack2(acc2[o], acc30[2 * o], acc30[2 * o + 1]);
                const float mmax = valid ? zmask : 0.f;
#pragma unroll
                for (int o = 0; o < 30; ++o) h[o] = valid ? mmax * prelu(acc30[o], ri_a1) : 0.f;
                h[30] = 0.f;
                h[31] = 0.f;
            }
            // ---- sum over the tile's rows; per-grid-node combination in a fixed order ---------------------------------------------
            const float colsum = warp_colsum32(h, lane);
            float* wp = wpart + ((wg * 2 + (int)(n & 1)) * 4) * 32;
            wp[wiw * 32 + lane] = colsum;
            named_bar_sync(1 + wg, 128);
            if (wiw == 0) {
                const float tot = (wp[lane] + wp[32 + lane]) + (wp[64 + lane] + wp[96 + lane]);
                float* pg = part + (slot * NT_MAX) * 32;
                pg[T * 32 + lane] = tot;
                __threadfence_block();
                __syncwarp();
                int old = 0;
                if (lane == 0) old = atomicAdd(&bars->count[slot], 1);
                old = __shfl_sync(FULL_MASK, old, 0);
                if (old == NT - 1) {     // last tile of grid node g: fc2
                    __threadfence_block();
                    if (lane == 0) bars->count[slot] = 0;
                    float xg = 0.f;
                    for (int t2 = 0; t2 < NT; ++t2) xg += pg[t2 * 32 + lane];
                    const float* W2 = sW + (RI_WFC2 - RI_WFC1);
                    float o = lane < 16 ? sW[(RI_BFC2 - RI_WFC1) + (lane & 15)] : 0.f;
#pragma unroll
                    for (int c = 0; c < 30; ++c) {
                        const float xc = __shfl_sync(FULL_MASK, xg, c);
                        o = fmaf(xc, W2[c * LD16 + (lane & 15)], o);
                    }
                    if (lane < 15) out[(int64_t)g * ld_out + lane] = prelu(o, ri_a2);
                }
            }
        }
    }
}

}  // namespace

int launch_da_layer2_s(const genie_plan* p, const float* packed, const float* zc, const float* va, const float* m2,
                       const float* mask, const float* edge_attr, float* latent_out, float* out, int ld_out,
                       cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    if (g.n_sta_tiles > NT_MAX) {
        set_error("launch_da_layer2_s: too many station tiles");
        return GENIE_ERR_UNSUPPORTED;
    }
    const int n_own = g.n_grid_owned > 0 ? g.n_grid_owned : g.n_grid;
    const unsigned grid = (unsigned)(n_own < p->sm_count ? n_own : p->sm_count);
    GENIE_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_ri_slots, packed + RI_WFC1, sizeof(float) * W_FLOATS, sizeof(float) * W_FLOATS * p->cslot,
                                             cudaMemcpyDeviceToDevice, st));
    static PerDeviceOnce attr_set;
    const bool set_attr = attr_set.need();
    TimedLaunch tl(KID_DA_LAYER2_S, st);
    const bool bf = p->storage == GENIE_STORAGE_BF16;
#define GENIE_S2_LAUNCH_B(L, C, B)                                                                                             \
    do {                                                                                                                        \
        if (set_attr)                                                                                                           \
            GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer2_s_kernel<L, C, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)); \
        if (L == (latent_out != nullptr) && C == p->cslot && B == bf)                                                           \
            da_layer2_s_kernel<L, C, B><<<grid, S2_THREADS, SM_TOTAL, st>>>(packed, zc, va, m2, mask, edge_attr, latent_out, out, ld_out, \
                                                                            g.n_sta, n_own, g.n_sta_tiles, g.sta_tile_rows,    \
                                                                            g.sta_tile_meta, g.sta_tile_nbr, g.sta_tile_invdeg, nullptr); \
    } while (0)
#define GENIE_S2_LAUNCH(L, C) GENIE_S2_LAUNCH_B(L, C, false); GENIE_S2_LAUNCH_B(L, C, true)
    GENIE_S2_LAUNCH(false, 0); GENIE_S2_LAUNCH(false, 1); GENIE_S2_LAUNCH(false, 2); GENIE_S2_LAUNCH(false, 3);
    GENIE_S2_LAUNCH(true, 0); GENIE_S2_LAUNCH(true, 1); GENIE_S2_LAUNCH(true, 2); GENIE_S2_LAUNCH(true, 3);
#undef GENIE_S2_LAUNCH
#undef GENIE_S2_LAUNCH_B
    if (set_attr) attr_set.mark();
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

// Layer 2 of the association phase on a plan with tiling tables: s rows [P][32] = PReLU(zc + [mean_sta v_a | m2]) with padding
// channels 15 and 31 zero.  m2: mean over source neighbours of v_b, [P][16] (src_mean_kernels.cu).
int launch_assoc_layer2_s(const genie_plan* p, const float* slope_dev, const float* zc, const float* va, const float* m2, float* s_out,
                          cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    if (g.n_sta_tiles > NT_MAX) {
        set_error("launch_assoc_layer2_s: too many station tiles");
        return GENIE_ERR_UNSUPPORTED;
    }
    const int n_own = g.n_grid_owned > 0 ? g.n_grid_owned : g.n_grid;
    const unsigned grid = (unsigned)(n_own < p->sm_count ? n_own : p->sm_count);
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer2_s_kernel<false, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set.mark();
    }
    TimedLaunch tl(KID_ASSOC_LAYER2, st);
    da_layer2_s_kernel<false, 0, false, true><<<grid, S2_THREADS, SM_TOTAL, st>>>(
        nullptr, zc, va, m2, nullptr, nullptr, s_out, nullptr, 0, g.n_sta, n_own, g.n_sta_tiles, g.sta_tile_rows, g.sta_tile_meta,
        g.sta_tile_nbr, g.sta_tile_invdeg, slope_dev);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
