// DataAggregation layer 1, station pass (+ the node-local half of layer 2) — CARTESIAN graphs with tiling tables.
//
// Mathematics as da_layer1_kernel / da_layer1_tc_kernel (reference module.py:88-96), split in two passes so that every
// gather runs on chip:
//   source pass  (src_mean_kernels.cu)   msrc[g,s,:] = mean_{g' in N_src(g)} p[g',s,:],     p = PReLU12(tr0)
//   station pass (this kernel)           one tile = (grid node g, compact set of <= 128 stations):
//     * producer warps stage, with cp.async, the p rows of the tile's stations AND of the halo of their station-graph
//       in-neighbours (<= 288 rows x 128 B, table genie_graph_desc_t.sta_tile_rows), the tile's msrc rows and its mask
//       rows into one of two shared-memory buffers;
//     * the gather warpgroup (thread per row) converts the staged rows in place to PReLU11(tr0), sums the <= 16 neighbour
//       rows of its station out of shared memory (16-byte chunks are visited in a per-lane rotated order, so arbitrary
//       rows are bank-conflict free), and writes the three A operands  [tr0 | mask | 1], mean_sta, mean_src  as 3xTF32
//       hi/lo parts straight into tensor memory;
//     * the MMA warp runs  stage B [.. ] -> tr (60),  stage C tr -> [h_a | h_b | c_a | c_b] (90),  stage D PReLU(h) ->
//       [v_a | v_b] (30)  as tcgen05.mma kind::tf32 (hi*hi + lo*hi + hi*lo), weights resident in shared memory in the
//       canonical K-major UMMA layout, A operands and accumulators in tensor memory;
//     * the epilogue warpgroup applies the activations between the stages and stores c (zc) and v_a / v_b.
// DRAM sees p, msrc and mask once (the tiles of one grid node are consecutive, its 128 KB block stays in L2); nothing is
// gathered from L2.  All hand-offs are mbarriers with bounded spins (a protocol bug traps, it never hangs).
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

using namespace gl;
using namespace tc;

namespace {

constexpr int S1_THREADS = 512;
constexpr int WARP_MMA = 0, WARP_ALLOC = 1;
constexpr int WG_G0 = 4, WG_E0 = 8, WG_P0 = 12;      // gather / epilogue / producer warpgroups (first warp)
constexpr int ROWS = GENIE_TILE_ROWS_MAX;            // staged p rows per tile; row ROWS is the zero row
constexpr int NBUF = 2;

// shared memory map (bytes)
constexpr int SB_P = 0;                              // [ROWS + 1][128 B]   p rows (tile stations first, then halo)
constexpr int SB_MS = SB_P + (ROWS + 1) * 128;       // [128][128 B]        msrc rows, 16-byte chunks XOR-swizzled by row
constexpr int SB_MK = SB_MS + 128 * 128;             // [128][16 B]         mask rows
constexpr int SB_SIZE = SB_MK + 128 * 16;
constexpr int SM_W = 0;                              // tensor-core weight blob (layout.h TC_*)
constexpr int SM_BUF = (TC_FLOATS * 4 + 1023) / 1024 * 1024;
constexpr int SM_BAR = SM_BUF + NBUF * SB_SIZE;
constexpr int SM_TOTAL = SM_BAR + 256;
static_assert(SB_SIZE % 16 == 0 && SM_BUF % 1024 == 0 && SM_BAR % 8 == 0, "alignment");
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// tensor memory map (columns)
constexpr int TM_OWN_HI = 0, TM_OWN_LO = 40;         // [tr0(30) | mask(4) | 1 | 0 x 5]
constexpr int TM_STA_HI = 80, TM_STA_LO = 112;       // mean over station neighbours of PReLU11(tr0)
constexpr int TM_SRC_HI = 144, TM_SRC_LO = 176;      // mean over source neighbours of PReLU12(tr0)
constexpr int TM_R2 = 208;                           // [0,64) A hi, [64,128) A lo   (tr, later PReLU(h))
constexpr int TM_D = 336;                            // 96 accumulator columns
constexpr int TM_COLS = 512;

struct Bars {
    uint64_t full[NBUF], empty[NBUF];
    uint64_t opA_full, opA_free;
    uint64_t d_full, aE_full, d_free;
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

// Development aid (genie_debug_trace): CTA 0 stamps clock64() at the hand-off points of its first tiles, 24 slots per tile.
#define S1_TRACE(slot)                                                                      \
    do {                                                                                    \
        if (trace != nullptr && blockIdx.x == 0 && it >= trace_start && it < trace_start + trace_tiles)                  \
            trace[(it - trace_start) * 24 + (slot)] = clock64();                                \
    } while (0)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// The mbarrier receives one arrival from this thread when all of its earlier cp.async have landed.
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// 16 consecutive fp32 values -> 3xTF32 parts -> TMEM columns [hi, hi+16) and [lo, lo+16)
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

// v[k] holds the 16-byte chunk (k ^ key) of a row; afterwards v[k] holds chunk k.
__device__ __forceinline__ void unrotate8(float4 (&v)[8], int key) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const bool sw = (key >> b) & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((k >> b) & 1) continue;
            const float4 a = v[k], c = v[k | (1 << b)];
            v[k] = sw ? c : a;
            v[k | (1 << b)] = sw ? a : c;
        }
    }
}

__global__ void __launch_bounds__(S1_THREADS, 1)
    da_layer1_s_kernel(const float* __restrict__ packed, const float* __restrict__ p, const float* __restrict__ msrc,
                       const float* __restrict__ mask, float* __restrict__ zc, float* __restrict__ va,
                       float* __restrict__ vb, int S, int NT, const int32_t* __restrict__ tile_rows,
                       const int32_t* __restrict__ tile_meta, const uint16_t* __restrict__ tile_nbr,
                       const float* __restrict__ tile_invdeg, int64_t n_tiles, long long* __restrict__ trace, int trace_tiles, int trace_start) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const float* tcw = packed + TC_BASE;
    if (tcw[TC_SCAL + TCS_OK] == 0.f) return;   // slopes not eligible: the generic kernels run instead (uniform exit)

    float* sW = reinterpret_cast<float*>(smem + SM_W);
    Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time set-up ------------------------------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(tcw);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < TC_FLOATS / 4; i += S1_THREADS) dst[i] = src[i];
        // the zero row of both buffers (padding target of the neighbour table)
        if (threadIdx.x < NBUF * 8) {
            const int b = threadIdx.x >> 3, c = threadIdx.x & 7;
            *reinterpret_cast<float4*>(smem + SM_BUF + b * SB_SIZE + SB_P + ROWS * 128 + c * 16) =
                make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&bars->full[b], 128);
            mbar_init(&bars->empty[b], 128);
        }
        mbar_init(&bars->opA_full, 128);
        mbar_init(&bars->opA_free, 1);
        mbar_init(&bars->d_full, 1);
        mbar_init(&bars->aE_full, 128);
        mbar_init(&bars->d_free, 128);
        fence_barrier_init();
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

    if (warp >= WG_P0) {
        // ================================ producers: cp.async row gather ==============================================
        // Thread `tid` owns the 16-byte chunk c = tid & 7 of the staged rows rr + 16 j (rr = tid >> 3).  The halo rows are
        // only ever read as PReLU11(tr0): the thread converts the chunks it copied itself (element-wise, so no other
        // thread is involved) before it arrives on the buffer's barrier; the tile's own rows stay p (the gather warpgroup
        // needs tr0 of its own row first and converts them itself).
        const int tid = threadIdx.x - WG_P0 * 32;
        const int rr = tid >> 3, c = tid & 7;
        const float r11 = sc[TCS_R11];
        constexpr int JMAX = (ROWS + 15) / 16;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int g = (int)(t / NT), T = (int)(t - (int64_t)g * NT);
            const int buf = (int)(it & 1);
            const int n_own = __ldg(tile_meta + 2 * T), n_rows = __ldg(tile_meta + 2 * T + 1);
            const int32_t* rows = tile_rows + (int64_t)T * ROWS;
            int ids[JMAX];
#pragma unroll
            for (int j = 0; j < JMAX; ++j) ids[j] = (rr + 16 * j) < n_rows ? __ldg(rows + rr + 16 * j) : -1;
            const int id_m = tid < n_own ? __ldg(rows + tid) : -1;
            if (it >= NBUF) mbar_wait(&bars->empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
            if (tid == 0) S1_TRACE(17);
            unsigned char* sbp = smem + SM_BUF + buf * SB_SIZE;
            const uint32_t sb = smem_u32(sbp);
            const int64_t node0 = (int64_t)g * S;
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (ids[j] >= 0) cp_async16(sb + SB_P + (rr + 16 * j) * 128 + c * 16, p + (node0 + ids[j]) * 32 + c * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = rr + 16 * j;
                if (r < n_own) cp_async16(sb + SB_MS + r * 128 + ((c ^ (r & 7)) << 4), msrc + (node0 + ids[j]) * 32 + c * 4);
            }
            if (id_m >= 0) cp_async16(sb + SB_MK + tid * 16, mask + (node0 + id_m) * 4);
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int j = 0; j < JMAX; ++j) {
                const int r = rr + 16 * j;
                if (r >= n_own && ids[j] >= 0) {
                    float4* a = reinterpret_cast<float4*>(sbp + SB_P + r * 128 + c * 16);
                    const float4 v = *a;
                    *a = make_float4(prelu_f(v.x, r11), prelu_f(v.y, r11), prelu_f(v.z, r11), prelu_f(v.w, r11));
                }
            }
            mbar_arrive(&bars->full[buf]);
            if (tid == 0) S1_TRACE(18);
        }
    } else if (warp == WARP_MMA) {
        // ================================ MMA issuer ==================================================================
        if (elect_one()) {
            const uint32_t wbase = smem_u32(sW);
            const uint32_t i64 = umma_idesc_tf32(128, 64), i32 = umma_idesc_tf32(128, 32);
            const uint32_t i96 = umma_idesc_tf32(128, 96), i16 = umma_idesc_tf32(128, 16);
            const uint32_t r2 = tm + TM_R2, d = tm + TM_D;
            uint32_t ph_a = 0;
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                mbar_wait(&bars->opA_full, (uint32_t)(it & 1));
                if (it > 0) mbar_wait(&bars->d_free, (uint32_t)((it - 1) & 1));
                tc_fence_after_sync();
                S1_TRACE(0);
                // ---- stage B: D[0,64) = [tr1 | tr2] pre-activation --------------------------------------------------
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const bool a_lo = pass == 1;                // A operand: lo part on pass 1
                    const bool b_lo = pass == 2;                // B operand: lo part on pass 2
                    const uint32_t b1a = wbase + 4 * (b_lo ? TC_B1A_LO : TC_B1A_HI);
                    const uint32_t b1b = wbase + 4 * (b_lo ? TC_B1B_LO : TC_B1B_HI);
                    const uint32_t b1c = wbase + 4 * (b_lo ? TC_B1C_LO : TC_B1C_HI);
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks)
                        umma_tf32_ts(d, tm + (a_lo ? TM_OWN_LO : TM_OWN_HI) + ks * 8,
                                     umma_desc_kmajor(b1a + ks * 2 * 64 * 16, 64 * 16, 128), i64, (pass | ks) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(d, tm + (a_lo ? TM_STA_LO : TM_STA_HI) + ks * 8,
                                     umma_desc_kmajor(b1b + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(d + 32, tm + (a_lo ? TM_SRC_LO : TM_SRC_HI) + ks * 8,
                                     umma_desc_kmajor(b1c + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
                }
                umma_commit(&bars->opA_free);
                umma_commit(&bars->d_full);
                S1_TRACE(1);
                // ---- stage C: D[0,96) (bias preloaded by the epilogue) += tr-row * B2 ----------------------------------
                mbar_wait(&bars->aE_full, ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
                S1_TRACE(2);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t b2 = wbase + 4 * (pass == 2 ? TC_B2_LO : TC_B2_HI);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(d, r2 + (pass == 1 ? 64 : 0) + ks * 8,
                                     umma_desc_kmajor(b2 + ks * 2 * 96 * 16, 96 * 16, 128), i96, 1u);
                }
                umma_commit(&bars->d_full);
                S1_TRACE(3);
                // ---- stage D: D[0,16) = v_a, D[16,32) = v_b --------------------------------------------------------------
                mbar_wait(&bars->aE_full, ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
                S1_TRACE(4);
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
                S1_TRACE(5);
            }
        }
    } else if (warp >= WG_G0 && warp < WG_G0 + 4) {
        // ================================ gather warpgroup (thread per row) ===========================================
        const int r = (warp - WG_G0) * 32 + lane;
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        const float inv12 = sc[TCS_INV12], r11 = sc[TCS_R11];
        const int key = lane & 7;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int T = (int)(t % NT);
            const int buf = (int)(it & 1);
            const int n_own = __ldg(tile_meta + 2 * T);
            // neighbour table of this row (staged-row indices; padding = the zero row) and 1 / degree
            const uint4* nb = reinterpret_cast<const uint4*>(tile_nbr + ((int64_t)T * 128 + r) * 16);
            const uint4 n0 = __ldg(nb), n1 = __ldg(nb + 1);
            const float invdeg = __ldg(tile_invdeg + T * 128 + r);
            unsigned char* sb = smem + SM_BUF + buf * SB_SIZE;
            mbar_wait(&bars->full[buf], (uint32_t)((it >> 1) & 1));
            if (r == 0) S1_TRACE(12);
            // ---- own row: keep tr0 = PReLU12^-1(p), leave PReLU11(tr0) in place (the halo rows were converted by the producers)
            float4 own[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) own[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < n_own) {
                unsigned char* ra = sb + SB_P + r * 128;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4* a = reinterpret_cast<float4*>(ra + ((k ^ key) << 4));
                    const float4 v = *a;
                    own[k] = make_float4(prelu_f(v.x, inv12), prelu_f(v.y, inv12), prelu_f(v.z, inv12), prelu_f(v.w, inv12));
                    *a = make_float4(prelu_f(v.x, r11), prelu_f(v.y, r11), prelu_f(v.z, r11), prelu_f(v.w, r11));
                }
            }
            unrotate8(own, key);
            named_bar_sync(1, 128);
            if (r == 0) S1_TRACE(13);
            // ---- sum of the station neighbours' rows (16-byte chunk k ^ key of every row: conflict free) ------------------
            float4 acc[8];
            {
                f32x4_t a2[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a2[k].lo = a2[k].hi = 0ull;
                const uint32_t w[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t idx = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
                    const unsigned char* ra = sb + SB_P + idx * 128;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        fadd4(a2[k], *reinterpret_cast<const float4*>(ra + ((k ^ key) << 4)));
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = to_float4(a2[k]);
            }
            unrotate8(acc, key);
            const bool valid = r < n_own;
            float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) mk = *reinterpret_cast<const float4*>(sb + SB_MK + r * 16);
            // ---- A operands -> tensor memory (free once stage B of the previous tile has completed) ------------------------
            if (r == 0) S1_TRACE(14);
            if (it > 0) mbar_wait(&bars->opA_free, (uint32_t)((it - 1) & 1));
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(15);
            {
                float a[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[4 * q] = own[q].x; a[4 * q + 1] = own[q].y; a[4 * q + 2] = own[q].z; a[4 * q + 3] = own[q].w;
                }
                st_split16(lane_base + TM_OWN_HI, lane_base + TM_OWN_LO, a);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[4 * q] = own[4 + q].x; a[4 * q + 1] = own[4 + q].y; a[4 * q + 2] = own[4 + q].z;
                    a[4 * q + 3] = own[4 + q].w;
                }
                a[14] = mk.x; a[15] = mk.y;          // channels 30, 31 of a feature row are padding
                st_split16(lane_base + TM_OWN_HI + 16, lane_base + TM_OWN_LO + 16, a);
                const float h8[8] = {mk.z, mk.w, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const float l8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                tmem_st8(lane_base + TM_OWN_HI + 32, h8);      // mask values and 1 are exact in tf32
                tmem_st8(lane_base + TM_OWN_LO + 32, l8);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = acc[4 * half + q];
                        a[4 * q] = v.x * invdeg; a[4 * q + 1] = v.y * invdeg; a[4 * q + 2] = v.z * invdeg;
                        a[4 * q + 3] = v.w * invdeg;
                    }
                    st_split16(lane_base + TM_STA_HI + 16 * half, lane_base + TM_STA_LO + 16 * half, a);
                    // own msrc row (16-byte chunks swizzled by the producer: conflict free, static registers)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid)
                            v = *reinterpret_cast<const float4*>(sb + SB_MS + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
                        a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
                    }
                    st_split16(lane_base + TM_SRC_HI + 16 * half, lane_base + TM_SRC_LO + 16 * half, a);
                }
            }
            mbar_arrive(&bars->empty[buf]);          // release: every shared-memory read of this tile has completed
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->opA_full);
            if (r == 0) S1_TRACE(16);
        }
    } else if (warp >= WG_E0 && warp < WG_E0 + 4) {
        // ================================ epilogue warpgroup (thread per row) ==========================================
        const int r = (warp - WG_E0) * 32 + lane;
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t r2 = lane_base + TM_R2, d = lane_base + TM_D;
        const float a1 = sc[TCS_A1], a21 = sc[TCS_A21], a22 = sc[TCS_A22];
        const float* bias2 = sW + TC_BIAS2;
        uint32_t ph_d = 0;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int g = (int)(t / NT), T = (int)(t - (int64_t)g * NT);
            const bool valid = r < __ldg(tile_meta + 2 * T);
            const int64_t node = (int64_t)g * S + (valid ? __ldg(tile_rows + (int64_t)T * ROWS + r) : 0);
            float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) mk = reinterpret_cast<const float4*>(mask)[node];
            // ---- stage B epilogue: tr = PReLU1(D) -> A operand of stage C (mask in the four spare columns) -----------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(6);
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
            if (r == 0) S1_TRACE(7);
            // ---- stage C epilogue: PReLU(h) -> A operand of stage D; c -> global ------------------------------------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(8);
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
                    for (int q = 0; q < 4; ++q)
                        __stcs(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->aE_full);
            if (r == 0) S1_TRACE(9);
            // ---- stage D epilogue: v_a, v_b -> global -------------------------------------------------------------------------
            mbar_wait(&bars->d_full, ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(10);
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
            if (r == 0) S1_TRACE(11);
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_ALLOC) tmem_dealloc(tm, TM_COLS);
}

}  // namespace

static long long* g_s1_trace = nullptr;
static int g_s1_trace_tiles = 0, g_s1_trace_start = 0;
void set_s1_trace(long long* buf, int tiles) {
    g_s1_trace = buf;
    g_s1_trace_tiles = tiles;
    const char* e = getenv("GENIE_TRACE_START");
    g_s1_trace_start = e ? atoi(e) : 0;
}

int launch_da_layer1_s(const genie_plan* p, const float* packed, const float* pfeat, const float* msrc,
                       const float* mask, float* zc, float* va, float* vb, cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    const int64_t n_tiles = (int64_t)g.n_sta_tiles * (g.n_grid_owned > 0 ? g.n_grid_owned : g.n_grid);
    static bool attr_set = false;
    if (!attr_set) {
        GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer1_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set = true;
    }
    const int64_t grid = n_tiles < p->sm_count ? n_tiles : p->sm_count;
    TimedLaunch tl(KID_DA_LAYER1_S, st);
    da_layer1_s_kernel<<<(unsigned)grid, S1_THREADS, SM_TOTAL, st>>>(packed, pfeat, msrc, mask, zc, va, vb, g.n_sta,
                                                                     g.n_sta_tiles, g.sta_tile_rows, g.sta_tile_meta,
                                                                     g.sta_tile_nbr, g.sta_tile_invdeg, n_tiles,
                                                                     g_s1_trace, g_s1_trace_tiles, g_s1_trace_start);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
