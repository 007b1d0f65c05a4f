// DataAggregation layer 1 (+ the node-local half of layer 2) on the 5th-generation tensor cores — CARTESIAN graphs.
//
// Same mathematics as da_layer1_kernel (da_kernels.cu; reference module.py:88-96), restructured for sm_100a:
//   * a tile is <= 128 consecutive stations of ONE grid node g, so every source-neighbour tile (g', same stations) is one
//     contiguous 16 KB block of the layer-0 output: it is fetched by TMA (cp.async.bulk.tensor, 128B swizzle) into a
//     shared-memory ring and summed thread-per-row, conflict-free, with no address arithmetic and no L1 traffic;
//   * layer 0 stores p = PReLU12(tr0), so that sum needs no per-edge activation; tr0 and PReLU11(tr0) are recovered
//     from p (PReLU with a positive slope is invertible);
//   * the three dense stages  [tr0 | mean_sta | mean_src | mask | 1] -> tr (60)  -> [h_a | h_b | c_a | c_b] (90)
//     -> [v_a | v_b] (30)  run as tcgen05.mma kind::tf32 with the 3xTF32 split (hi*hi + lo*hi + hi*lo, error ~1e-6),
//     accumulators and the thread-per-row A operands in tensor memory, weights resident in shared memory in the
//     canonical K-major UMMA layout;
//   * warp-specialised, one persistent CTA per SM: TMA producer | MMA issuer | source-gather warpgroup |
//     epilogue warpgroup | station-gather warps, with the stage-B operands double-buffered so that the gathers of tile
//     i+1 overlap the MMAs / epilogues of tile i.  All hand-offs are mbarriers (bounded spins: a bug traps, never hangs).
#include "common.cuh"
#include "tc_common.cuh"

using namespace gl;
using namespace tc;

namespace {

constexpr int TC_THREADS = 512;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_ALLOC = 2;
constexpr int WG_S0 = 4, WG_E0 = 8, WG_T0 = 12, N_T_WARPS = 4;
constexpr int NSTAGE = 4;
constexpr int TILE_BYTES = 128 * 128;
constexpr int CS_A = 129 * 16;               // chunk stride of the station-mean A operand (padded: conflict-free stores)
constexpr int A_HALF = 8 * CS_A;             // hi or lo part of one buffer
constexpr int A_BUF = 2 * A_HALF;

// shared memory map (bytes)
constexpr int SM_RING = 0;
constexpr int SM_W = SM_RING + NSTAGE * TILE_BYTES;
constexpr int SM_A = SM_W + TC_FLOATS * 4;
constexpr int SM_BAR = SM_A + 2 * A_BUF;
constexpr int SM_TOTAL = SM_BAR + 256;
static_assert(SM_W % 1024 == 0 && SM_A % 16 == 0 && SM_BAR % 8 == 0, "alignment");
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// tensor memory map (columns)
constexpr int TM_R1 = 0;        // + buf * 144: [0,40) tr0-row hi, [40,80) lo, [80,112) mean_src hi, [112,144) lo
constexpr int TM_R1_STRIDE = 144;
constexpr int TM_R2 = 288;      // [0,64) A hi, [64,128) A lo   (tr, later PReLU(h))
constexpr int TM_D = 416;       // 96 accumulator columns
constexpr int TM_COLS = 512;

struct Bars {
    uint64_t ring_full[NSTAGE], ring_empty[NSTAGE];
    uint64_t opS_full[2], opT_full[2], opB_free[2];
    uint64_t d_full, aE_full, d_free;
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

struct TileInfo {
    int g, s0, cnt;
};

// tile t -> (station tile st = t / G, grid node = the (t % G)-th node of the locality order)
__device__ __forceinline__ TileInfo tile_of(int64_t t, int G, int S, int R, const int32_t* __restrict__ order) {
    TileInfo ti;
    const int st = (int)(t / G);
    const int gi = (int)(t - (int64_t)st * G);
    ti.g = order ? __ldg(order + gi) : gi;
    ti.s0 = st * R;
    ti.cnt = min(R, S - ti.s0);
    return ti;
}

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

// mbarrier arrive that cannot be issued before the registers `a`, `b` are available (i.e. before the shared-memory loads
// that produced them have completed): rz is zero at run time, but the compiler cannot know it, so the address depends on
// the loaded data.  Used to hand a TMA ring stage back to the producer only after it has really been read.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t rz, uint32_t dep) {
    const uint32_t addr = smem_u32(bar) + (dep & rz);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// 16 consecutive fp32 values -> 3xTF32 parts -> TMEM columns [col, col+16) (hi) and [col + lo_off, ...) (lo)
__device__ __forceinline__ void st_split16(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[16]) {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        h[i] = tf32_hi(v[i]);
        l[i] = v[i] - h[i];
    }
    tmem_st16(taddr_hi, h);
    tmem_st16(taddr_lo, l);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    da_layer1_tc_kernel(const __grid_constant__ CUtensorMap tmap_p, const GraphView gv, const float* __restrict__ packed,
                        const float* __restrict__ p, const float* __restrict__ mask, float* __restrict__ zc,
                        float* __restrict__ va, float* __restrict__ vb, int R, int64_t n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const float* tcw = packed + TC_BASE;
    if (tcw[TC_SCAL + TCS_OK] == 0.f) return;   // slopes not eligible: the generic kernels run instead (uniform exit)

    float* sW = reinterpret_cast<float*>(smem + SM_W);
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = gv.S, G = gv.G;

    // ---- one-time set-up ------------------------------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(tcw);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < TC_FLOATS / 4; i += TC_THREADS) dst[i] = src[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&bars->ring_full[s], 1);
            mbar_init(&bars->ring_empty[s], 128);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->opS_full[b], 128);
            mbar_init(&bars->opT_full[b], N_T_WARPS * 32);
            mbar_init(&bars->opB_free[b], 1);
        }
        mbar_init(&bars->d_full, 1);
        mbar_init(&bars->aE_full, 128);
        mbar_init(&bars->d_free, 128);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_p);
    }
    if (warp == WARP_ALLOC) {
        tmem_alloc(&bars->tmem_base, TM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();     // weights written with generic stores, read by tcgen05.mma
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = bars->tmem_base;
    const float* sc = sW + TC_SCAL;

    if (warp == WARP_TMA) {
        // ================================ TMA producer ================================================================
        int stage = 0, phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const TileInfo ti = tile_of(t, G, S, R, gv.grid_order);
            const int64_t beg = gv.src_rowptr[ti.g];
            const int deg = (int)(gv.src_rowptr[ti.g + 1] - beg);
            for (int j0 = -1; j0 < deg; j0 += 32) {
                const int j = j0 + lane;
                int gj = -1;
                if (j < deg) gj = j < 0 ? ti.g : gv.src_col[beg + j];
                const int n = min(32, deg - j0);
                for (int u = 0; u < n; ++u) {
                    const int gu = __shfl_sync(FULL_MASK, gj, u);
                    if (lane == 0) {
                        mbar_wait(&bars->ring_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bars->ring_full[stage], TILE_BYTES);
                        tma_load_2d(smem + SM_RING + stage * TILE_BYTES, &tmap_p, 0, gu * S + ti.s0,
                                    &bars->ring_full[stage]);
                    }
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ================================ MMA issuer ==================================================================
        if (elect_one()) {
            const uint32_t wbase = smem_u32(sW);
            const uint32_t i64 = umma_idesc_tf32(128, 64), i32 = umma_idesc_tf32(128, 32);
            const uint32_t i96 = umma_idesc_tf32(128, 96), i16 = umma_idesc_tf32(128, 16);
            uint32_t ph_a = 0;
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int buf = (int)(it & 1);
                const uint32_t use_par = (uint32_t)((it >> 1) & 1);
                mbar_wait(&bars->opS_full[buf], use_par);
                mbar_wait(&bars->opT_full[buf], use_par);
                if (it > 0) mbar_wait(&bars->d_free, (uint32_t)((it - 1) & 1));
                tc_fence_after_sync();
                const uint32_t r1 = tm + TM_R1 + buf * TM_R1_STRIDE;
                const uint32_t r2 = tm + TM_R2, d = tm + TM_D;
                const uint32_t a_sta = smem_u32(smem + SM_A + buf * A_BUF);
                // ---- stage B: D[0,64) = [tr1 | tr2] pre-activation --------------------------------------------------
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_lo = pass == 1;            // A operand: lo part on pass 1
                    const uint32_t b_lo = pass == 2;            // B operand: lo part on pass 2
                    const uint32_t b1a = wbase + 4 * (b_lo ? TC_B1A_LO : TC_B1A_HI);
                    const uint32_t b1b = wbase + 4 * (b_lo ? TC_B1B_LO : TC_B1B_HI);
                    const uint32_t b1c = wbase + 4 * (b_lo ? TC_B1C_LO : TC_B1C_HI);
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks)
                        umma_tf32_ts(d, r1 + (a_lo ? 40 : 0) + ks * 8, umma_desc_kmajor(b1a + ks * 2 * 64 * 16, 64 * 16, 128),
                                     i64, (pass | ks) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ss(d, umma_desc_kmajor(a_sta + (a_lo ? A_HALF : 0) + ks * 2 * CS_A, CS_A, 128),
                                     umma_desc_kmajor(b1b + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(d + 32, r1 + (a_lo ? 112 : 80) + ks * 8,
                                     umma_desc_kmajor(b1c + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
                }
                umma_commit(&bars->opB_free[buf]);
                umma_commit(&bars->d_full);
                // ---- stage C: D[0,96) (bias preloaded by the epilogue) += tr-row * B2 ----------------------------------
                mbar_wait(&bars->aE_full, ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t b2 = wbase + 4 * (pass == 2 ? TC_B2_LO : TC_B2_HI);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(d, r2 + (pass == 1 ? 64 : 0) + ks * 8,
                                     umma_desc_kmajor(b2 + ks * 2 * 96 * 16, 96 * 16, 128), i96, 1u);
                }
                umma_commit(&bars->d_full);
                // ---- stage D: D[0,16) = v_a, D[16,32) = v_b --------------------------------------------------------------
                mbar_wait(&bars->aE_full, ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t b3a = wbase + 4 * (pass == 2 ? TC_B3A_LO : TC_B3A_HI);
                    const uint32_t b3b = wbase + 4 * (pass == 2 ? TC_B3B_LO : TC_B3B_HI);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_tf32_ts(d, r2 + (pass == 1 ? 64 : 0) + ks * 8,
                                     umma_desc_kmajor(b3a + ks * 2 * 16 * 16, 16 * 16, 128), i16, (pass | ks) ? 1u : 0u);
                        umma_tf32_ts(d + 16, r2 + (pass == 1 ? 64 : 0) + 32 + ks * 8,
                                     umma_desc_kmajor(b3b + ks * 2 * 16 * 16, 16 * 16, 128), i16, (pass | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&bars->d_full);
            }
        }
    } else if (warp >= WG_S0 && warp < WG_S0 + 4) {
        // ================================ source-gather warpgroup (thread per row) =====================================
        const int r = (warp - WG_S0) * 32 + lane;
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        const float inv12 = sc[TCS_INV12];
        const int swz = r & 7;
        const uint32_t rz = (uint32_t)((uint64_t)n_tiles >> 62);   // 0, opaque to the compiler
        int stage = 0, phase = 0;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const TileInfo ti = tile_of(t, G, S, R, gv.grid_order);
            const int buf = (int)(it & 1);
            const int deg = (int)(gv.src_rowptr[ti.g + 1] - gv.src_rowptr[ti.g]);
            float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < ti.cnt) mk = reinterpret_cast<const float4*>(mask)[(int64_t)ti.g * S + ti.s0 + r];
            if (it >= 2) mbar_wait(&bars->opB_free[buf], (uint32_t)(((it >> 1) - 1) & 1));
            const uint32_t r1 = lane_base + TM_R1 + buf * TM_R1_STRIDE;
            // ---- own row: tr0 = PReLU12^-1(p) ------------------------------------------------------------------------------
            {
                mbar_wait(&bars->ring_full[stage], phase);
                const unsigned char* row = smem + SM_RING + stage * TILE_BYTES + r * 128;
                float x[40];
                uint32_t dep = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ swz) << 4));
                    dep |= __float_as_uint(v.x);      // one word of every 16-byte load
                    x[4 * c + 0] = v.x >= 0.f ? v.x : v.x * inv12;
                    x[4 * c + 1] = v.y >= 0.f ? v.y : v.y * inv12;
                    x[4 * c + 2] = v.z >= 0.f ? v.z : v.z * inv12;
                    x[4 * c + 3] = v.w >= 0.f ? v.w : v.w * inv12;
                }
                mbar_arrive_after(&bars->ring_empty[stage], rz, dep);
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
                x[30] = mk.x; x[31] = mk.y; x[32] = mk.z; x[33] = mk.w; x[34] = 1.f;
#pragma unroll
                for (int i = 35; i < 40; ++i) x[i] = 0.f;
                {
                    float a[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = x[i];
                    st_split16(r1 + 0, r1 + 40, a);
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = x[16 + i];
                    st_split16(r1 + 16, r1 + 56, a);
                    float h8[8], l8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        h8[i] = tf32_hi(x[32 + i]);
                        l8[i] = x[32 + i] - h8[i];
                    }
                    tmem_st8(r1 + 32, h8);
                    tmem_st8(r1 + 72, l8);
                }
            }
            // ---- sum of the source neighbours' p rows (already activated) --------------------------------------------------
            float2 acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
            for (int j = 0; j < deg; ++j) {
                mbar_wait(&bars->ring_full[stage], phase);
                const unsigned char* row = smem + SM_RING + stage * TILE_BYTES + r * 128;
                uint32_t dep = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ swz) << 4));
                    dep |= __float_as_uint(v.x);
                    acc[2 * c] = __fadd2_rn(acc[2 * c], make_float2(v.x, v.y));
                    acc[2 * c + 1] = __fadd2_rn(acc[2 * c + 1], make_float2(v.z, v.w));
                }
                mbar_arrive_after(&bars->ring_empty[stage], rz, dep);
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            {
                const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
                float a[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[2 * i] = acc[i].x * inv;
                    a[2 * i + 1] = acc[i].y * inv;
                }
                st_split16(r1 + 80, r1 + 112, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[2 * i] = acc[8 + i].x * inv;
                    a[2 * i + 1] = acc[8 + i].y * inv;
                }
                st_split16(r1 + 96, r1 + 128, a);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->opS_full[buf]);
        }
    } else if (warp >= WG_E0 && warp < WG_E0 + 4) {
        // ================================ epilogue warpgroup (thread per row) ==========================================
        const int r = (warp - WG_E0) * 32 + lane;
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t r2 = lane_base + TM_R2, d = lane_base + TM_D;
        const float a1 = sc[TCS_A1], a21 = sc[TCS_A21], a22 = sc[TCS_A22];
        const float* bias2 = sW + TC_BIAS2;
        uint32_t ph_d = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const TileInfo ti = tile_of(t, G, S, R, gv.grid_order);
            const bool valid = r < ti.cnt;
            const int64_t node = (int64_t)ti.g * S + ti.s0 + r;
            float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) mk = reinterpret_cast<const float4*>(mask)[node];
            // ---- stage B epilogue: tr = PReLU1(D) -> A operand of stage C (mask in the four spare columns) -----------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                float v[16];
                tmem_ld16(d + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = prelu_f(v[i], a1);
                if (c == 16) {
                    v[14] = mk.x;
                    v[15] = mk.y;
                }
                if (c == 48) {
                    v[14] = mk.z;
                    v[15] = mk.w;
                }
                st_split16(r2 + c, r2 + 64 + c, v);
            }
#pragma unroll
            for (int c = 0; c < 96; c += 16) {     // accumulator <- bias of stage C
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b = *reinterpret_cast<const float4*>(bias2 + c + 4 * q);
                    v[4 * q] = b.x; v[4 * q + 1] = b.y; v[4 * q + 2] = b.z; v[4 * q + 3] = b.w;
                }
                tmem_st16(d + c, v);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->aE_full);
            // ---- stage C epilogue: PReLU(h) -> A operand of stage D; c -> global ------------------------------------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                float v[16];
                tmem_ld16(d + c, v);
                tmem_ld_wait();
                const float a = c < 32 ? a21 : a22;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = prelu_f(v[i], a);
                st_split16(r2 + c, r2 + 64 + c, v);
            }
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                float v[16];
                tmem_ld16(d + 64 + c, v);
                tmem_ld_wait();
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(zc + node * LD_ZC + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q)      // streaming: written once, read by the next kernel - keep L2 for the gathers
                        __stcs(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->aE_full);
            // ---- stage D epilogue: v_a, v_b -> global -------------------------------------------------------------------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                float v[16];
                tmem_ld16(d + c, v);
                tmem_ld_wait();
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>((c ? vb : va) + node * LD_V);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        __stcs(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bars->d_free);
        }
    } else if (warp >= WG_T0 && warp < WG_T0 + N_T_WARPS) {
        // ================================ station-gather warps (8 lanes x float4 per row) ================================
        // Row slots of this quarter-warp: rr = (warp - WG_T0) * 4 + q + 16 j, j = 0..7.  The station graph is the same for
        // every grid node, so the (<= 16) neighbour ids of the 8 slots stay in registers while the station tile is unchanged
        // (lane `sub` keeps entries sub and sub + 8) and are broadcast with shuffles.
        const int q = lane >> 3, sub = lane & 7;
        const float r11 = sc[TCS_R11];
        constexpr int ROWS_T = 128 / (N_T_WARPS * 4);
        int idx_lo[ROWS_T], idx_hi[ROWS_T], degs[ROWS_T];
        int cur_s0 = -1;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const TileInfo ti = tile_of(t, G, S, R, gv.grid_order);
            const int buf = (int)(it & 1);
            if (ti.s0 != cur_s0) {
                cur_s0 = ti.s0;
#pragma unroll
                for (int j = 0; j < ROWS_T; ++j) {
                    const int rr = (warp - WG_T0) * 4 + q + 16 * j;
                    int64_t beg = 0;
                    int deg = 0;
                    if (rr < ti.cnt) {
                        beg = gv.sta_rowptr[ti.s0 + rr];
                        deg = (int)(gv.sta_rowptr[ti.s0 + rr + 1] - beg);
                    }
                    degs[j] = deg;
                    idx_lo[j] = sub < deg ? gv.sta_col[beg + sub] : -1;
                    idx_hi[j] = sub + 8 < deg ? gv.sta_col[beg + sub + 8] : -1;
                }
            }
            if (it >= 2) mbar_wait(&bars->opB_free[buf], (uint32_t)(((it >> 1) - 1) & 1));
            unsigned char* a_hi = smem + SM_A + buf * A_BUF;
            const float* pg = p + ((int64_t)ti.g * S) * 32 + 4 * sub;
#pragma unroll
            for (int j = 0; j < ROWS_T; ++j) {
                const int rr = (warp - WG_T0) * 4 + q + 16 * j;
                const int deg = degs[j];
                // all 16 loads are issued unconditionally (absent edges re-read row 0 of the block and are masked in the
                // sum): predicated loads would be serialised through one temporary register
                float4 v[16];
                unsigned okmask = 0;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int jj = __shfl_sync(FULL_MASK, u < 8 ? idx_lo[j] : idx_hi[j], (lane & 24) | (u & 7));
                    okmask |= (jj >= 0 ? 1u : 0u) << u;
                    v[u] = __ldg(reinterpret_cast<const float4*>(pg + (int64_t)max(jj, 0) * 32));
                }
                float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const float w = (okmask >> u) & 1u ? 1.f : 0.f;
                    s01 = __ffma2_rn(make_float2(prelu_f(v[u].x, r11), prelu_f(v[u].y, r11)), make_float2(w, w), s01);
                    s23 = __ffma2_rn(make_float2(prelu_f(v[u].z, r11), prelu_f(v[u].w, r11)), make_float2(w, w), s23);
                }
                const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
                const float m0 = s01.x * inv, m1 = s01.y * inv, m2 = s23.x * inv, m3 = s23.y * inv;
                const float4 h = make_float4(tf32_hi(m0), tf32_hi(m1), tf32_hi(m2), tf32_hi(m3));
                const float4 l = make_float4(m0 - h.x, m1 - h.y, m2 - h.z, m3 - h.w);
                *reinterpret_cast<float4*>(a_hi + sub * CS_A + rr * 16) = h;
                *reinterpret_cast<float4*>(a_hi + A_HALF + sub * CS_A + rr * 16) = l;
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars->opT_full[buf]);
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_ALLOC) tmem_dealloc(tm, TM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

}  // namespace

bool da_tc_supported(const genie_plan* p) {
    const genie_graph_desc_t& g = p->g;
    return g.mode == GENIE_GRAPH_CARTESIAN && g.n_sta >= 32 && g.n_prod > 0 && g.n_prod < (int64_t)0x7fffff00 &&
           g.sta_max_deg >= 1 && g.sta_max_deg <= 16 && encode_fn() != nullptr;
}

int launch_da_layer1_tc(const genie_plan* p, const float* packed, const float* pfeat, const float* mask, float* zc,
                        float* va, float* vb, cudaStream_t st) {
    const int S = p->g.n_sta, G = p->g.n_grid;
    const int n_st = (S + 127) / 128;
    const int R = (S + n_st - 1) / n_st;
    const int64_t n_tiles = (int64_t)n_st * G;
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {32, (cuuint64_t)p->g.n_prod};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pfeat), gdim, gstr, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return GENIE_ERR_CUDA;
    }
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set.mark();
    }
    const int64_t grid = n_tiles < p->sm_count ? n_tiles : p->sm_count;
    TimedLaunch tl(KID_DA_LAYER1_TC, st);
    da_layer1_tc_kernel<<<(unsigned)grid, TC_THREADS, SM_TOTAL, st>>>(tmap, make_view(p), packed, pfeat, mask, zc, va, vb,
                                                                      R, n_tiles);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
