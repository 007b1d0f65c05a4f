// DataAggregation layer 1, station pass (+ the node-local half of layer 2) — CARTESIAN graphs with tiling tables.
//
// Mathematics as da_layer1_kernel / da_layer1_tc_kernel (reference module.py:88-96), split in two passes so that every
// gather runs on chip:
//   source pass  (src_mean_kernels.cu)   msrc[g,s,:] = mean_{g' in N_src(g)} p[g',s,:],     p = PReLU12(tr0)
//   station pass (this kernel)           one tile = (grid node g, compact set of <= 128 stations).
// The CTA runs TWO tile pipelines (even / odd tiles) that share the producers and the weights in shared memory; a pipeline
// owns one shared-memory buffer, 256 tensor-memory columns, a gather warpgroup, an epilogue warpgroup and an MMA-issuing
// warp, so the gather of one tile, the tensor-core chain of another and the epilogues overlap:
//     * producer warps stage, with cp.async, the p rows of the tile's stations AND of the halo of their station-graph
//       in-neighbours (<= 288 rows x 128 B, table genie_graph_desc_t.sta_tile_rows), the tile's msrc rows and its mask
//       rows (`mask` == NULL: the four mask values ride bit-packed in channel 30 of the p rows — a1 fused into layer 0,
//       genie_window_fwd — and a compact copy of that chunk is staged instead);
//       the (otherwise waiting) gather warpgroup converts the staged p rows in place to PReLU11(tr0), the only form they
//       are read in.  (Measured and rejected in round 2, profiles/r4_s1_experiments.md: the conversion folded into the
//       gather as max(p, r p), the rows passed through the producers' registers, conversion by the producers per commit
//       group, conversion by the gather warpgroup per commit group, and a start-up phase offset between the pipelines.)
//     * the gather warpgroup (thread per row) sums the <= 16 neighbour rows of its station out of shared memory (16-byte
//       chunks are visited in a per-lane rotated order, so arbitrary rows are bank-conflict free), recovers tr0 of its
//       own row, and writes the three 32-column A operands  [tr0 | mask0,1], [mean_sta | mask2,3], [mean_src | mask2,3]
//       as 3xTF32 hi/lo parts straight into tensor memory;
//     * the MMA warp runs  stage B [.. ] -> tr (60),  stage C tr -> [h_a | h_b | c_a | c_b] (90),  stage D PReLU(h) ->
//       [v_a | v_b] (30)  as tcgen05.mma kind::tf32 (hi*hi + lo*hi + hi*lo), weights resident in shared memory in the
//       canonical K-major UMMA layout (layout.h T2_*: biases ride on the lo pass), A operands and accumulators in
//       tensor memory; the operands of stages C and D and the accumulators alias the stage-B operands (dead by then):
//           columns   0-191  stage-B operands  ->  0-127 A operand of stage C, then of stage D
//           columns 192-255  stage-B accumulator (tr)
//           columns 128-223  stage-C accumulator ([h | c]),  224-255 stage-D accumulator ([v_a | v_b])
//     * the epilogue warpgroup applies the activations between the stages and stores c (zc) and v_a / v_b.
// ASSOC = true: layer 1 of DataAggregationAssociationPhase (module.py:387-398; assoc_kernels.cu) on the same pipeline.  The
// association phase differs in three places: its messages are PReLU(l1_t*_1 tr) (two more row tensors, written by
// assoc_init_kernel: the staged / gathered rows are those, never converted, and the node's own tr row is read from global
// memory by the epilogue warpgroup after an L2 prefetch by the producers), its mask has a fifth channel — the source mask
// of the row's grid node, constant over a tile, so its weight columns are added in the epilogues (blob floats
// T2_FLOATS ..+96) — and its weights (pack.cu assoc_pack_t2_kernel builds the blob in the same T2_* layout).
// DRAM sees p, msrc and mask once (the tiles of one grid node are consecutive, its 128 KB block stays in L2); nothing is
// gathered from L2.  All hand-offs are mbarriers with bounded spins (a protocol bug traps, it never hangs).
#include "bf16.cuh"
#include "common.cuh"
#include "input.cuh"
#include "tc_common.cuh"
#include <cstdlib>

using namespace gl;
using namespace tc;

namespace {

constexpr int S1_THREADS = 768;
constexpr int WARP_ALLOC = 2;                         // warps 0, 1: MMA issuers of pipelines 0, 1
constexpr int WG_G0 = 4, WG_E0 = 12, WG_P0 = 20;     // gather (2 x 4 warps) / epilogue (2 x 4 warps) / producer warps
constexpr int ROWS = GENIE_TILE_ROWS_MAX;            // staged p rows per tile; row ROWS is the zero row
constexpr int NPIPE = 2;
constexpr int P_THREADS = 64;                        // producer threads per pipeline (two warps)
constexpr int S1_ASSOC_TERMS = 96;                   // association blob: weight columns of the source mask, [tr1 | tr2 | c_a | c_b]

// shared memory map (bytes).  Row format of the staged tensors: fp32 rows (128 B; ONE staging buffer per pipeline, shared
// memory is full) or the bf16 rows of the fast storage mode (genie_plan_set_storage; 64 B, TWO staging buffers per pipeline:
// the fill of a pipeline's next tile overlaps the gather of its current one — the producer -> convert -> gather -> release
// loop of a single buffer is what bounds the fp32 kernel, profiles/r4_s1_experiments.md).
constexpr int SM_W = 0;                              // tensor-core weight blob (layout.h T2_*)
constexpr int SM_BUF = (T2_FLOATS * 4 + 1023) / 1024 * 1024;
template <bool BF16>
struct Fmt {
    static constexpr int RB = BF16 ? 64 : 128;       // bytes of a staged feature row
    static constexpr int CPR = RB / 16;              // 16-byte chunks per row
    static constexpr int NBUF = BF16 ? 2 : 1;        // staging buffers per pipeline
    static constexpr int SB_P = 0;                   // [ROWS + 1][RB]   staged rows (tile stations first, then halo)
    static constexpr int SB_MS = SB_P + (ROWS + 1) * RB;   // [128][RB]    msrc rows, 16-byte chunks XOR-swizzled by row
    static constexpr int SB_MK = SB_MS + 128 * RB;   // [128][16 B]      mask rows
    static constexpr int SB_SIZE = SB_MK + 128 * 16;
    static constexpr int SM_BAR = SM_BUF + NPIPE * NBUF * SB_SIZE;
    static constexpr int SM_SCR = SM_BAR + 256;      // [8 epilogue warps][32 rows][64 B] store-transposition scratch
    static constexpr int SM_TOTAL = SM_SCR + 8 * 2048;
    static_assert(SB_SIZE % 16 == 0 && SM_BUF % 1024 == 0 && SM_BAR % 8 == 0, "alignment");
    static_assert(SM_TOTAL <= 232448, "shared memory budget");
};

// tensor memory map (columns, relative to the pipeline's 256-column block)
constexpr int TM_OWN_HI = 0, TM_OWN_LO = 32;         // [tr0(30) | mask0 mask1], lo part [.. | 1 1] (bias of stage B)
constexpr int TM_STA_HI = 64, TM_STA_LO = 96;        // [mean over station neighbours of PReLU11(tr0) | mask2 mask3]
constexpr int TM_SRC_HI = 128, TM_SRC_LO = 160;      // [mean over source neighbours of PReLU12(tr0) | mask2 mask3]
constexpr int TM_X = 192;                            // stage-B accumulator, 64 columns
constexpr int TM_R2_HI = 0, TM_R2_LO = 64;           // A operand of stages C and D (aliases OWN / STA)
constexpr int TM_DC = 128;                           // stage-C accumulator, 96 columns (aliases SRC and X)
constexpr int TM_DD = 224;                           // stage-D accumulator, 32 columns
constexpr int TM_PIPE = 256;
constexpr int TM_COLS = 512;

struct Bars {
    uint64_t full[NPIPE][2], empty[NPIPE][2];      // per staging buffer
    uint64_t opA_full[NPIPE], opA_free[NPIPE];
    uint64_t d_full[NPIPE], aE_full[NPIPE], d_free[NPIPE];
    uint64_t raw[NPIPE][2];        // copies landed (producers -> gather warpgroup, which converts the rows in place)
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

// Development aid (genie_debug_trace): CTA 0 stamps clock64() at the hand-off points of its tiles, 24 slots per tile and
// pipeline (trace[tile][pipeline][slot]).
#define S1_TRACE(slot)                                                                                     \
    do {                                                                                                   \
        if (trace != nullptr && blockIdx.x == 0 && k >= trace_start && k < trace_start + trace_tiles)      \
            trace[(k - trace_start) * 48 + q * 24 + (slot)] = clock64();                                   \
    } while (0)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// 16 consecutive fp32 values -> 3xTF32 parts -> TMEM columns [hi, hi+16) and [lo, lo+16).
// FAST (the bf16-storage mode): the tensor-core stages run single-pass TF32 — the operand goes in as it is (the tensor core
// reads the top 19 bits) and no lo part is written; the one lo block a bias needs is written by st_split16_bias.
template <bool FAST>
__device__ __forceinline__ void st_split16(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[16]) {
    if (FAST) {
        tmem_st16(taddr_hi, v);
        return;
    }
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        h[i] = tf32_hi(v[i]);
        l[i] = v[i] - h[i];
    }
    tmem_st16(taddr_hi, h);
    tmem_st16(taddr_lo, l);
}
// as st_split16, with lo columns 14 and 15 forced to 1 (the bias rows of the lo pass, layout.h T2_*_BIAS)
template <bool FAST>
__device__ __forceinline__ void st_split16_bias(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[16]) {
    if (FAST) {      // lo columns 8-15 of this half = k 24-31 of the operand: zeros, and ones in the two bias columns
        tmem_st16(taddr_hi, v);
        const float ones[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1.f};
        tmem_st8(taddr_lo + 8, ones);
        return;
    }
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        h[i] = tf32_hi(v[i]);
        l[i] = i >= 14 ? 1.f : v[i] - h[i];
    }
    tmem_st16(taddr_hi, h);
    tmem_st16(taddr_lo, l);
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

// Coalescing store of 16 consecutive floats per row (thread = row of the warp's 32-row block): the chunk goes through a
// warp-private 2 KB scratch (16-byte pieces XOR-swizzled: conflict free both ways) and leaves as 64-byte runs, four lanes
// per row — 8 lines per store instruction instead of 32 (a thread-per-row STG.128 costs one L1 wavefront per lane).
//   sid[j] = station id of row (lane >> 2) + 8 j of the warp's block, or -1; dst + (node0 + sid) * ld is the row's address.
__device__ __forceinline__ void store16_rows(const float (&v)[16], unsigned char* scr, int lane, const int (&sid)[4],
                                             float* __restrict__ dst, int64_t node0, int ld) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(scr + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
            make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    const int chunk = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int row = (lane >> 2) + 8 * j;
        const float4 x = *reinterpret_cast<const float4*>(scr + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
        if (sid[j] >= 0) __stcs(reinterpret_cast<float4*>(dst + (node0 + sid[j]) * ld) + chunk, x);
    }
    __syncwarp();
}

// bf16 variant: 16 floats -> 32-byte row pieces of [P][16] bf16 rows, two lanes per row, 16 rows per store instruction.
//   sid2[j] = station id of row (lane >> 1) + 16 j of the warp's block, or -1.
__device__ __forceinline__ void store16_rows_bf16(const float (&v)[16], unsigned char* scr, int lane, const int (&sid2)[2],
                                                  float* __restrict__ dst, int64_t node0) {
    const float lo8[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
    const float hi8[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
    const int sw = (lane >> 2) & 1;
    *reinterpret_cast<uint4*>(scr + lane * 32 + ((0 ^ sw) << 4)) = bf16_pack8(lo8);
    *reinterpret_cast<uint4*>(scr + lane * 32 + ((1 ^ sw) << 4)) = bf16_pack8(hi8);
    __syncwarp();
    const int chunk = lane & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (lane >> 1) + 16 * j;
        const uint4 x = *reinterpret_cast<const uint4*>(scr + row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4));
        if (sid2[j] >= 0) __stcs(reinterpret_cast<uint4*>(dst) + (node0 + sid2[j]) * 2 + chunk, x);
    }
    __syncwarp();
}

// v[k][0..7] holds the 8 channels of 16-byte chunk (k ^ key) of a bf16 row; afterwards v[k] holds chunk k.
__device__ __forceinline__ void unrotate4x8(float (&v)[4][8], int key) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const bool sw = (key >> b) & 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((k >> b) & 1) continue;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float a = v[k][e], c = v[k | (1 << b)][e];
                v[k][e] = sw ? c : a;
                v[k | (1 << b)][e] = sw ? a : c;
            }
        }
    }
}

// The row's four mask values: the caller's Mask row, or (packed) channel 30 of the own p row, of which the producers stage a
// compact copy of the last 16-byte chunk (fp32: channels 28-31, bf16: channels 24-31).
template <bool BF16>
__device__ __forceinline__ float4 s1_load_mask(const unsigned char* sb, int r, bool valid, bool packed) {
    float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        mk = *reinterpret_cast<const float4*>(sb + Fmt<BF16>::SB_MK + r * 16);
        if (packed) mk = unpack_mask(BF16 ? bf16_lo(__float_as_uint(mk.w)) : mk.z);
    }
    return mk;
}

// Epilogue-warpgroup half of the stage-B operands of one tile: [tr0 | mask0,1] (tr0 recovered from the staged PReLU11(tr0)
// of the thread's own row) and [mean_src | mask2,3] (the thread's msrc row) -> tensor memory.  Returns the row's mask.

template <bool BF16, bool ASSOC>
__device__ __forceinline__ float4 s1_own_operands(const unsigned char* sb, int r, bool valid, int key, float inv11,
                                                  uint32_t lane_base, bool packed, const float* __restrict__ own_row) {
    using F = Fmt<BF16>;
    const float4 mk = s1_load_mask<BF16>(sb, r, valid, packed);
    float a[16];
    if (ASSOC) {
        // the node's own tr row: 128 contiguous bytes from global memory (L2: prefetched by the producers with the tile)
        float4 own[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) own[c] = valid ? __ldg(reinterpret_cast<const float4*>(own_row) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = own[4 * half + u];
                a[4 * u] = v.x; a[4 * u + 1] = v.y; a[4 * u + 2] = v.z; a[4 * u + 3] = v.w;
            }
            if (half) {
                a[14] = mk.x;
                a[15] = mk.y;
                st_split16_bias<BF16>(lane_base + TM_OWN_HI + 16, lane_base + TM_OWN_LO + 16, a);
            } else {
                st_split16<BF16>(lane_base + TM_OWN_HI, lane_base + TM_OWN_LO, a);
            }
        }
    }
    if (BF16) {
        // 64-byte rows: four chunks of 8 channels; own row in the per-lane rotated order (key = lane & 3), msrc row swizzled
        // by (row >> 1) & 3 — conflict free, static registers
        float own[4][8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (valid) u = *reinterpret_cast<const uint4*>(sb + F::SB_P + r * 64 + ((c ^ key) << 4));
            bf16_unpack8(u, own[c]);
        }
        unrotate4x8(own, key);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = prelu_f(own[2 * half + (i >> 3)][i & 7], inv11);
            if (half) {
                a[14] = mk.x;
                a[15] = mk.y;
                st_split16_bias<BF16>(lane_base + TM_OWN_HI + 16, lane_base + TM_OWN_LO + 16, a);      // bias of stage B
            } else {
                st_split16<BF16>(lane_base + TM_OWN_HI, lane_base + TM_OWN_LO, a);
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int u2 = 0; u2 < 2; ++u2) {
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (valid) u = *reinterpret_cast<const uint4*>(sb + F::SB_MS + r * 64 + (((2 * half + u2) ^ ((r >> 1) & 3)) << 4));
                float f[8];
                bf16_unpack8(u, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[8 * u2 + e] = f[e];
            }
            if (half) {
                a[14] = mk.z;
                a[15] = mk.w;
            }
            st_split16<BF16>(lane_base + TM_SRC_HI + 16 * half, lane_base + TM_SRC_LO + 16 * half, a);
        }
        return mk;
    }
    if (!ASSOC) {
        float4 own[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) own[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const unsigned char* ra = sb + F::SB_P + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) own[c] = *reinterpret_cast<const float4*>(ra + ((c ^ key) << 4));
        }
        unrotate8(own, key);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = own[4 * half + u];
                a[4 * u] = prelu_f(v.x, inv11); a[4 * u + 1] = prelu_f(v.y, inv11);
                a[4 * u + 2] = prelu_f(v.z, inv11); a[4 * u + 3] = prelu_f(v.w, inv11);
            }
            if (half) {
                a[14] = mk.x;       // channels 30, 31 of a feature row are padding: the mask rides there
                a[15] = mk.y;
                st_split16_bias<BF16>(lane_base + TM_OWN_HI + 16, lane_base + TM_OWN_LO + 16, a);      // bias of stage B
            } else {
                st_split16<BF16>(lane_base + TM_OWN_HI, lane_base + TM_OWN_LO, a);
            }
        }
    }
    // own msrc row (16-byte chunks swizzled by the producer: conflict free, static registers)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) v = *reinterpret_cast<const float4*>(sb + F::SB_MS + r * 128 + (((4 * half + u) ^ (r & 7)) << 4));
            a[4 * u] = v.x; a[4 * u + 1] = v.y; a[4 * u + 2] = v.z; a[4 * u + 3] = v.w;
        }
        if (half) {
            a[14] = mk.z;
            a[15] = mk.w;
        }
        st_split16<BF16>(lane_base + TM_SRC_HI + 16 * half, lane_base + TM_SRC_LO + 16 * half, a);
    }
    return mk;
}

template <bool EDGE, bool BF16, bool ASSOC, bool EXPORT>
__global__ void __launch_bounds__(S1_THREADS, 1)
    da_layer1_s_kernel(const float* __restrict__ tcw, const float* __restrict__ p, const float* __restrict__ msrc,
                       const float* __restrict__ mask, float* __restrict__ zc, float* __restrict__ va,
                       float* __restrict__ vb, int S, int NT, const int32_t* __restrict__ tile_rows,
                       const int32_t* __restrict__ tile_meta, const uint16_t* __restrict__ tile_nbr,
                       const float* __restrict__ tile_invdeg, int64_t n_tiles, const float* __restrict__ edge_sta,
                       const float* __restrict__ edge_src, const float* __restrict__ tr_own,
                       const float* __restrict__ mask_out, const int32_t* __restrict__ exp_ptr,
                       const int32_t* __restrict__ exp_peer, const int32_t* __restrict__ exp_row, float* const* __restrict__ peer_base,
                       long long* __restrict__ trace, int trace_tiles, int trace_start) {
    static_assert(!(ASSOC && BF16), "the association rows are fp32");
    extern __shared__ __align__(1024) unsigned char smem[];
    using F = Fmt<BF16>;
    constexpr int NBUF = F::NBUF;
    // Who writes the OWN / SRC stage-B operands: the epilogue warpgroup (after its last epilogue of the previous tile), so that
    // the gather warpgroup can hand the staging buffer back before it touches tensor memory.  Moving them to the gather
    // warpgroup in the bf16 kernel (where the epilogue warpgroup's serial chain is the bound, profiles/r5c_trace_bf16.log)
    // was measured and lost: 8.6 -> 10.0 ms at C4, the extra shared-memory reads land in the gather phase (r9a).
    constexpr bool OWN_BY_GATHER = false;
    const bool packed_mask = mask == nullptr;
    if (tcw[T2_SCAL + TCS_OK] == 0.f) return;   // slopes not eligible: the generic kernels run instead (uniform exit)

    float* sW = reinterpret_cast<float*>(smem + SM_W);
    Bars* bars = reinterpret_cast<Bars*>(smem + F::SM_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time set-up ------------------------------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(tcw);
        float4* dst = reinterpret_cast<float4*>(sW);
        constexpr int NW4 = (T2_FLOATS + (ASSOC ? S1_ASSOC_TERMS : 0)) / 4;
        static_assert(NW4 * 16 <= SM_BUF, "weight blob");
        for (int i = threadIdx.x; i < NW4; i += S1_THREADS) dst[i] = src[i];
        // the zero row of every staging buffer (padding target of the neighbour table)
        if (threadIdx.x < NPIPE * NBUF * F::CPR) {
            const int b = threadIdx.x / F::CPR, c = threadIdx.x % F::CPR;
            *reinterpret_cast<float4*>(smem + SM_BUF + b * F::SB_SIZE + F::SB_P + ROWS * F::RB + c * 16) =
                make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < NPIPE; ++b) {
            for (int u = 0; u < 2; ++u) {
                mbar_init(&bars->full[b][u], 128);    // gather warpgroup, after the in-place conversion
                mbar_init(&bars->raw[b][u], P_THREADS);
                mbar_init(&bars->empty[b][u], 256);   // gather + epilogue warpgroups
            }
            mbar_init(&bars->opA_full[b], OWN_BY_GATHER ? 128 : 256);
            mbar_init(&bars->opA_free[b], 1);
            mbar_init(&bars->d_full[b], 1);
            mbar_init(&bars->aE_full[b], 128);
            mbar_init(&bars->d_free[b], 128);
        }
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
    const float* sc = sW + T2_SCAL;

    if (warp >= WG_P0) {
        // ================================ producers of pipeline q: row gather ==========================================
        // Thread `tid` owns the 16-byte chunk c of the staged rows rr + RPP j (fp32 rows: c = tid & 7, rr = tid >> 3; bf16
        // rows: c = tid & 3, rr = tid >> 2); everything goes by cp.async.
        const int q = (warp - WG_P0) >> 1;
        const int tid = threadIdx.x - (WG_P0 + 2 * q) * 32;
        constexpr int CPR = F::CPR, RPP = P_THREADS / CPR;         // chunks per row; rows per pass of the 64 threads (8 / 16)
        const int rr = tid / CPR, c = tid % CPR;
        constexpr int JMAX = (ROWS + RPP - 1) / RPP, JOWN = 128 / RPP;
        const unsigned char* pb = reinterpret_cast<const unsigned char*>(p);
        const unsigned char* mb = reinterpret_cast<const unsigned char*>(msrc);
        int64_t k = 0;
        for (int64_t t = blockIdx.x + (int64_t)q * gridDim.x; t < n_tiles; t += 2 * (int64_t)gridDim.x, ++k) {
            const int bi = (int)(k % NBUF);
            const uint32_t n = (uint32_t)(k / NBUF);
            const uint32_t sb = smem_u32(smem + SM_BUF + (q * NBUF + bi) * F::SB_SIZE);
            const int g = (int)(t / NT), T = (int)(t - (int64_t)g * NT);
            const int n_own = __ldg(tile_meta + 2 * T), n_rows = __ldg(tile_meta + 2 * T + 1);
            const int32_t* rows = tile_rows + (int64_t)T * ROWS;
            int ids[JMAX];
#pragma unroll
            for (int j = 0; j < JMAX; ++j) ids[j] = (rr + RPP * j) < n_rows ? __ldg(rows + rr + RPP * j) : -1;
            const int id_m0 = tid < n_own ? __ldg(rows + tid) : -1;
            const int id_m1 = tid + 64 < n_own ? __ldg(rows + tid + 64) : -1;
            if (n > 0) mbar_wait(&bars->empty[q][bi], (n - 1) & 1);
            if (tid == 0) S1_TRACE(17);
            const int64_t node0 = (int64_t)g * S;
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (ids[j] >= 0) cp_async16(sb + F::SB_P + (rr + RPP * j) * F::RB + c * 16, pb + (node0 + ids[j]) * F::RB + c * 16);
            if (ASSOC && c == 0) {
#pragma unroll
                for (int j = 0; j < JOWN; ++j)
                    if (rr + RPP * j < n_own)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(tr_own + (node0 + ids[j]) * 32));
            }
#pragma unroll
            for (int j = 0; j < JOWN; ++j) {
                const int r = rr + RPP * j;
                const int sw = BF16 ? ((r >> 1) & 3) : (r & 7);
                if (r < n_own) cp_async16(sb + F::SB_MS + r * F::RB + ((c ^ sw) << 4), mb + (node0 + ids[j]) * F::RB + c * 16);
            }
            {   // mask rows: from the caller's Mask [P,4], or (packed) a compact copy of the last chunk of the own p rows, whose
                // channel 30 carries the four mask bits — thread-per-row readers get it without bank conflicts
                const unsigned char* mrow = packed_mask ? pb + (F::RB - 16) : reinterpret_cast<const unsigned char*>(mask);
                const int mld = packed_mask ? F::RB : 16;
                if (id_m0 >= 0) cp_async16(sb + F::SB_MK + tid * 16, mrow + (node0 + id_m0) * mld);
                if (id_m1 >= 0) cp_async16(sb + F::SB_MK + (tid + 64) * 16, mrow + (node0 + id_m1) * mld);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            mbar_arrive(&bars->raw[q][bi]);
            if (tid == 0) S1_TRACE(18);

        }
    } else if (warp < NPIPE) {
        // ================================ MMA issuer of pipeline q ====================================================
        const int q = warp;
        if (elect_one()) {
            const uint32_t wbase = smem_u32(sW);
            const uint32_t i64 = umma_idesc_tf32(128, 64), i32 = umma_idesc_tf32(128, 32);
            const uint32_t i96 = umma_idesc_tf32(128, 96), i16 = umma_idesc_tf32(128, 16);
            const uint32_t base = tm + q * TM_PIPE;
            uint32_t ph_a = 0;
            int64_t k = 0;
            for (int64_t t = blockIdx.x + (int64_t)q * gridDim.x; t < n_tiles; t += 2 * (int64_t)gridDim.x, ++k) {
                mbar_wait(&bars->opA_full[q], (uint32_t)(k & 1));
                if (k > 0) mbar_wait(&bars->d_free[q], (uint32_t)((k - 1) & 1));
                tc_fence_after_sync();
                S1_TRACE(0);
                // ---- stage B: X[0,64) = [tr1 | tr2] pre-activation (bias on the lo pass) ------------------------------
                if (BF16) {      // fast mode: single-pass TF32 + the one lo k-step that carries the bias
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(base + TM_X, base + TM_OWN_HI + ks * 8,
                                     umma_desc_kmajor(wbase + 4 * T2_S1A_HI + ks * 2 * 64 * 16, 64 * 16, 128), i64, ks ? 1u : 0u);
                    umma_tf32_ts(base + TM_X, base + TM_OWN_LO + 3 * 8, umma_desc_kmajor(wbase + 4 * T2_S1A_BIAS, 64 * 16, 128), i64, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(base + TM_X, base + TM_STA_HI + ks * 8,
                                     umma_desc_kmajor(wbase + 4 * T2_S1B_HI + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(base + TM_X + 32, base + TM_SRC_HI + ks * 8,
                                     umma_desc_kmajor(wbase + 4 * T2_S1C_HI + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
                }
#pragma unroll
                for (int pass = 0; pass < (BF16 ? 0 : 3); ++pass) {
                    const bool a_lo = pass == 1;                // A operand: lo part on pass 1
                    const bool b_lo = pass == 2;                // B operand: lo part on pass 2
                    const uint32_t s1a = wbase + 4 * (b_lo ? T2_S1A_LO : T2_S1A_HI);
                    const uint32_t s1b = wbase + 4 * (b_lo ? T2_S1B_LO : T2_S1B_HI);
                    const uint32_t s1c = wbase + 4 * (b_lo ? T2_S1C_LO : T2_S1C_HI);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t baddr = (a_lo && ks == 3) ? wbase + 4 * T2_S1A_BIAS : s1a + ks * 2 * 64 * 16;
                        umma_tf32_ts(base + TM_X, base + (a_lo ? TM_OWN_LO : TM_OWN_HI) + ks * 8,
                                     umma_desc_kmajor(baddr, 64 * 16, 128), i64, (pass | ks) ? 1u : 0u);
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(base + TM_X, base + (a_lo ? TM_STA_LO : TM_STA_HI) + ks * 8,
                                     umma_desc_kmajor(s1b + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(base + TM_X + 32, base + (a_lo ? TM_SRC_LO : TM_SRC_HI) + ks * 8,
                                     umma_desc_kmajor(s1c + ks * 2 * 32 * 16, 32 * 16, 128), i32, 1u);
                }
                umma_commit(&bars->d_full[q]);
                S1_TRACE(1);
                // ---- stage C: DC[0,96) = tr-row * S2 (+ bias on the lo pass) ----------------------------------------------
                mbar_wait(&bars->aE_full[q], ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
                S1_TRACE(2);
                if (BF16) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(base + TM_DC, base + TM_R2_HI + ks * 8,
                                     umma_desc_kmajor(wbase + 4 * T2_S2_HI + ks * 2 * 96 * 16, 96 * 16, 128), i96, ks ? 1u : 0u);
                    umma_tf32_ts(base + TM_DC, base + TM_R2_LO + 3 * 8, umma_desc_kmajor(wbase + 4 * T2_S2_BIAS, 96 * 16, 128), i96, 1u);
                }
#pragma unroll
                for (int pass = 0; pass < (BF16 ? 0 : 3); ++pass) {
                    const uint32_t s2 = wbase + 4 * (pass == 2 ? T2_S2_LO : T2_S2_HI);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t baddr = (pass == 1 && ks == 3) ? wbase + 4 * T2_S2_BIAS : s2 + ks * 2 * 96 * 16;
                        umma_tf32_ts(base + TM_DC, base + (pass == 1 ? TM_R2_LO : TM_R2_HI) + ks * 8,
                                     umma_desc_kmajor(baddr, 96 * 16, 128), i96, (pass | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&bars->d_full[q]);
                S1_TRACE(3);
                // ---- stage D: DD[0,16) = v_a, DD[16,32) = v_b -------------------------------------------------------------
                mbar_wait(&bars->aE_full[q], ph_a);
                ph_a ^= 1;
                tc_fence_after_sync();
                S1_TRACE(4);
#pragma unroll
                for (int pass = 0; pass < (BF16 ? 1 : 3); ++pass) {
                    const uint32_t s3a = wbase + 4 * (pass == 2 ? T2_S3A_LO : T2_S3A_HI);
                    const uint32_t s3b = wbase + 4 * (pass == 2 ? T2_S3B_LO : T2_S3B_HI);
                    const uint32_t a = base + (pass == 1 ? TM_R2_LO : TM_R2_HI);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_tf32_ts(base + TM_DD, a + ks * 8, umma_desc_kmajor(s3a + ks * 2 * 16 * 16, 16 * 16, 128), i16,
                                     (pass | ks) ? 1u : 0u);
                        umma_tf32_ts(base + TM_DD + 16, a + 32 + ks * 8,
                                     umma_desc_kmajor(s3b + ks * 2 * 16 * 16, 16 * 16, 128), i16, (pass | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&bars->opA_free[q]);      // columns 0-191 may be overwritten by the next tile's operands
                umma_commit(&bars->d_full[q]);
                S1_TRACE(5);
            }
        }
    } else if (warp >= WG_G0 && warp < WG_G0 + 4 * NPIPE) {
        // ================================ gather warpgroup of pipeline q (thread per row) ==============================
        const int q = (warp - WG_G0) >> 2;
        const int r = ((warp - WG_G0) & 3) * 32 + lane;
        const uint32_t lane_base = tm + q * TM_PIPE + ((uint32_t)((warp & 3) * 32) << 16);
        const int key = lane & (F::CPR - 1);
        const float r11 = sc[TCS_R11];
        int64_t k = 0;
        for (int64_t t = blockIdx.x + (int64_t)q * gridDim.x; t < n_tiles; t += 2 * (int64_t)gridDim.x, ++k) {
            const int bi = (int)(k % NBUF);
            const uint32_t n = (uint32_t)(k / NBUF);
            unsigned char* sb = smem + SM_BUF + (q * NBUF + bi) * F::SB_SIZE;
            const int T = (int)(t % NT);
            const int n_own = __ldg(tile_meta + 2 * T);
            // neighbour table of this row (staged-row indices; padding = the zero row) and 1 / degree
            const uint4* nb = reinterpret_cast<const uint4*>(tile_nbr + ((int64_t)T * 128 + r) * 16);
            const uint4 n0 = __ldg(nb), n1 = __ldg(nb + 1);
            const float invdeg = __ldg(tile_invdeg + T * 128 + r);
            // ---- staged p rows -> PReLU11(tr0), in place: 16-byte chunk r & 7 of the rows (r >> 3) + 16 j, in batches whose
            //      loads are all in flight before the first store ----------------------------------------------------------
            mbar_wait(&bars->raw[q][bi], n & 1);
            const int n_rows = __ldg(tile_meta + 2 * T + 1);
            if (ASSOC) {
                // the staged rows are the messages themselves
            } else if (BF16) {
                // 64-byte rows: chunk r & 3 of the rows (r >> 2) + 32 u
                unsigned char* cb = sb + F::SB_P + (r >> 2) * 64 + (r & 3) * 16;
                constexpr int CB = (ROWS + 31) / 32;
                uint4 v[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u) v[u] = *reinterpret_cast<const uint4*>(cb + u * 32 * 64);
#pragma unroll
                for (int u = 0; u < CB; ++u)
                    if ((r >> 2) + 32 * u < n_rows) {
                        float f[8];
                        bf16_unpack8(v[u], f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = prelu_f(f[e], r11);
                        *reinterpret_cast<uint4*>(cb + u * 32 * 64) = bf16_pack8(f);
                    }
            } else {
                unsigned char* cb = sb + F::SB_P + (r >> 3) * 128 + (r & 7) * 16;
                constexpr int CB = 9;
                static_assert((ROWS + 15) / 16 == 2 * CB, "conversion batches");
#pragma unroll
                for (int j0 = 0; j0 < 2 * CB; j0 += CB) {
                    if ((r >> 3) + 16 * j0 >= n_rows) break;
                    float4 v[CB];
#pragma unroll
                    for (int u = 0; u < CB; ++u) v[u] = *reinterpret_cast<const float4*>(cb + (j0 + u) * 16 * 128);
#pragma unroll
                    for (int u = 0; u < CB; ++u)
                        if ((r >> 3) + 16 * (j0 + u) < n_rows)
                            *reinterpret_cast<float4*>(cb + (j0 + u) * 16 * 128) = make_float4(
                                prelu_f(v[u].x, r11), prelu_f(v[u].y, r11), prelu_f(v[u].z, r11), prelu_f(v[u].w, r11));
                }
            }
            mbar_arrive(&bars->full[q][bi]);
            mbar_wait(&bars->full[q][bi], n & 1);
            if (r == 0) S1_TRACE(12);
            // ---- sum of the station neighbours' rows (16-byte chunk k ^ key of every row: conflict free) ------------------
            float mean_sta[32];                      // channels 0-31 of the neighbour SUM (scaled by 1 / degree below)
            const uint32_t w[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
            if (BF16) {
                // (64-byte rows cover half of the banks: lanes l and l + 4 read the same chunk position and collide when their
                // rows have the same parity — plan._pair_neighbour_order orders the neighbour lists against that)
                f32x2_t a2[4][4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 4; ++e) a2[c][e] = 0ull;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t idx = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
                    const unsigned char* ra = sb + F::SB_P + idx * 64;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 u = *reinterpret_cast<const uint4*>(ra + ((c ^ key) << 4));
                        fadd2(a2[c][0], pack2(bf16_lo(u.x), bf16_hi(u.x)));
                        fadd2(a2[c][1], pack2(bf16_lo(u.y), bf16_hi(u.y)));
                        fadd2(a2[c][2], pack2(bf16_lo(u.z), bf16_hi(u.z)));
                        fadd2(a2[c][3], pack2(bf16_lo(u.w), bf16_hi(u.w)));
                    }
                }
                float acc8[4][8];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 4; ++e) unpack2(a2[c][e], acc8[c][2 * e], acc8[c][2 * e + 1]);
                unrotate4x8(acc8, key);
#pragma unroll
                for (int i = 0; i < 32; ++i) mean_sta[i] = acc8[i >> 3][i & 7];
            } else {
                float4 acc[8];
                {
                    f32x4_t a2[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) a2[c].lo = a2[c].hi = 0ull;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t idx = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
                        const unsigned char* ra = sb + F::SB_P + idx * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) fadd4(a2[c], *reinterpret_cast<const float4*>(ra + ((c ^ key) << 4)));
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = to_float4(a2[c]);
                }
                unrotate8(acc, key);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    mean_sta[4 * c] = acc[c].x; mean_sta[4 * c + 1] = acc[c].y;
                    mean_sta[4 * c + 2] = acc[c].z; mean_sta[4 * c + 3] = acc[c].w;
                }
            }
            const bool valid = r < n_own;
            const float4 mk = s1_load_mask<BF16>(sb, r, valid, packed_mask);
            if (!OWN_BY_GATHER) mbar_arrive(&bars->empty[q][bi]);     // release (gather half): every shared-memory read is done
            // ---- A operand -> tensor memory (free once stage D of the pipeline's previous tile has completed) ---------------
            if (r == 0) S1_TRACE(14);
            if (k > 0) mbar_wait(&bars->opA_free[q], (uint32_t)((k - 1) & 1));
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(15);
            {
                float a[16];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) a[i] = mean_sta[16 * half + i] * invdeg;
                    if (half) {
                        a[14] = mk.z;       // channels 30, 31 of a feature row are padding: the mask rides there
                        a[15] = mk.w;
                    }
                    st_split16<BF16>(lane_base + TM_STA_HI + 16 * half, lane_base + TM_STA_LO + 16 * half, a);
                }
            }
            if (OWN_BY_GATHER) {
                s1_own_operands<BF16, false>(sb, r, valid, key, sc[TCS_INV11], lane_base, packed_mask, nullptr);
                mbar_arrive(&bars->empty[q][bi]);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->opA_full[q]);
            if (r == 0) S1_TRACE(16);
        }
    } else if (warp >= WG_E0 && warp < WG_E0 + 4 * NPIPE) {
        // ================================ epilogue warpgroup of pipeline q (thread per row) ============================
        const int q = (warp - WG_E0) >> 2;
        const int r = ((warp - WG_E0) & 3) * 32 + lane;
        const uint32_t lane_base = tm + q * TM_PIPE + ((uint32_t)((warp & 3) * 32) << 16);
        const float a1 = sc[TCS_A1], a21 = sc[TCS_A21], a22 = sc[TCS_A22], inv11 = sc[TCS_INV11];
        const int key = lane & (F::CPR - 1);
        unsigned char* scr = smem + F::SM_SCR + (warp - WG_E0) * 2048;
        const int row0 = ((warp - WG_E0) & 3) * 32;                 // first tile row of this warp
        uint32_t ph_d = 0;
        int64_t k = 0;
        // The warpgroup also writes the OWN / SRC halves of the stage-B operands (the gather warpgroup writes STA): for the
        // first tile up front, for every later tile right after the previous tile's last epilogue.
        bool valid = false;
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const int64_t t = blockIdx.x + (int64_t)q * gridDim.x;
            if (t < n_tiles) {
                const int g = (int)(t / NT), T = (int)(t - (int64_t)g * NT);
                valid = r < __ldg(tile_meta + 2 * T);
                mbar_wait(&bars->full[q][0], 0u);
                if (OWN_BY_GATHER) {          // only the row's mask is needed here (for the epilogues)
                    mk = s1_load_mask<BF16>(smem + SM_BUF + (q * NBUF) * F::SB_SIZE, r, valid, packed_mask);
                    mbar_arrive(&bars->empty[q][0]);
                } else {
                    const float* own_row = (ASSOC && valid) ? tr_own + ((int64_t)g * S + __ldg(tile_rows + (int64_t)T * ROWS + r)) * 32 : nullptr;
                    mk = s1_own_operands<BF16, ASSOC>(smem + SM_BUF + (q * NBUF) * F::SB_SIZE, r, valid, key, inv11, lane_base, packed_mask,
                                                      own_row);
                    mbar_arrive(&bars->empty[q][0]);
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bars->opA_full[q]);
                }
            }
        }
        for (int64_t t = blockIdx.x + (int64_t)q * gridDim.x; t < n_tiles; t += 2 * (int64_t)gridDim.x, ++k) {
            const int g = (int)(t / NT), T = (int)(t - (int64_t)g * NT);
            const int64_t node0 = (int64_t)g * S;
            int sid[4], sid2[2];
            {
                const int n_own = __ldg(tile_meta + 2 * T);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rw = row0 + (lane >> 2) + 8 * j;
                    sid[j] = rw < n_own ? __ldg(tile_rows + (int64_t)T * ROWS + rw) : -1;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int rw = row0 + (lane >> 1) + 16 * j;
                    sid2[j] = (BF16 && rw < n_own) ? __ldg(tile_rows + (int64_t)T * ROWS + rw) : -1;
                }
            }
            // grid-sharded plan: the export range of this tile's grid node (fetched here, long before the stage-D epilogue uses it)
            int exp_b = 0, exp_e = 0;
            if (EXPORT) {
                exp_b = __ldg(exp_ptr + g);
                exp_e = __ldg(exp_ptr + g + 1);
            }
            // edge-feature model (genie_plan_set_edge_terms): rows of the additive terms of this thread's station / grid node
            const float* et_sta = nullptr;
            const float* et_src = nullptr;
            if (EDGE) {
                const int srow = valid ? __ldg(tile_rows + (int64_t)T * ROWS + r) : 0;
                et_sta = edge_sta + (int64_t)srow * GENIE_EDGE_TERM_LD;
                et_src = edge_src + (int64_t)g * GENIE_EDGE_TERM_LD;
            }
            const float mo = ASSOC ? __ldg(mask_out + g) : 0.f;      // source mask of the tile's grid node (fifth mask channel)
            const float* term = sW + T2_FLOATS;
            // ---- stage B epilogue: tr = PReLU1(X) -> A operand of stage C (mask in the four spare columns) -----------------
            mbar_wait(&bars->d_full[q], ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(6);
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                float v[16];
                tmem_ld16(lane_base + TM_X + c, v);
                tmem_ld_wait();
                if (EDGE) {       // edge-feature model: tr1 += station term, tr2 += grid-node term
                    const float4* e4 = reinterpret_cast<const float4*>((c < 32 ? et_sta : et_src) + (c & 16));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 e = __ldg(e4 + u);
                        v[4 * u] += e.x; v[4 * u + 1] += e.y; v[4 * u + 2] += e.z; v[4 * u + 3] += e.w;
                    }
                }
                if (ASSOC) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaf(mo, term[c + i], v[i]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = prelu_f(v[i], a1);
                if (c == 16) {
                    v[14] = mk.x;
                    v[15] = mk.y;
                    st_split16_bias<BF16>(lane_base + TM_R2_HI + c, lane_base + TM_R2_LO + c, v);     // bias of stage C
                } else {
                    if (c == 48) {
                        v[14] = mk.z;
                        v[15] = mk.w;
                    }
                    st_split16<BF16>(lane_base + TM_R2_HI + c, lane_base + TM_R2_LO + c, v);
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->aE_full[q]);
            if (r == 0) S1_TRACE(7);
            // ---- stage C epilogue: PReLU(h) -> A operand of stage D; c -> global ------------------------------------------
            mbar_wait(&bars->d_full[q], ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(8);
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                float v[16];
                tmem_ld16(lane_base + TM_DC + c, v);
                tmem_ld_wait();
                const float a = c < 32 ? a21 : a22;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = prelu_f(v[i], a);
                st_split16<BF16>(lane_base + TM_R2_HI + c, lane_base + TM_R2_LO + c, v);
            }
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                float v[16];
                tmem_ld16(lane_base + TM_DC + 64 + c, v);
                tmem_ld_wait();
                if (EDGE) {       // edge-feature model: c_a += station term, c_b += grid-node term
                    const float4* e4 = reinterpret_cast<const float4*>((c ? et_src : et_sta) + 32);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 e = __ldg(e4 + u);
                        v[4 * u] += e.x; v[4 * u + 1] += e.y; v[4 * u + 2] += e.z; v[4 * u + 3] += e.w;
                    }
                }
                if (ASSOC) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaf(mo, term[64 + c + i], v[i]);
                }
                if (c == 0) v[15] = fmaxf(fmaxf(mk.x, mk.y), fmaxf(mk.z, mk.w));     // padding channel 15: max_c(mask) for layer 2's read-in
                store16_rows(v, scr, lane, sid, zc + c, node0, LD_ZC);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&bars->aE_full[q]);
            if (r == 0) S1_TRACE(9);
            // ---- stage D epilogue: v_a, v_b -> global -------------------------------------------------------------------------
            mbar_wait(&bars->d_full[q], ph_d);
            ph_d ^= 1;
            tc_fence_after_sync();
            if (r == 0) S1_TRACE(10);
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
                float v[16];
                tmem_ld16(lane_base + TM_DD + c, v);
                tmem_ld_wait();
                if (BF16 && c) store16_rows_bf16(v, scr, lane, sid2, vb, node0);      // v_b is a gathered tensor: bf16 rows
                else store16_rows(v, scr, lane, sid, c ? vb : va, node0, LD_V);
                if (EXPORT && c) {
                    // grid-sharded plan: the v_b rows of a grid node that peers hold as halo also go straight into the peers'
                    // landing buffers (peer stores over NVLink; the transfer rides on this kernel instead of an all-to-all)
                    for (int e = exp_b; e < exp_e; ++e) {
                        float* pb = peer_base[__ldg(exp_peer + e)];
                        const int64_t rnode0 = (int64_t)__ldg(exp_row + e) * S;
                        if (BF16) store16_rows_bf16(v, scr, lane, sid2, pb, rnode0);
                        else store16_rows(v, scr, lane, sid, pb, rnode0, LD_V);
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bars->d_free[q]);
            if (r == 0) S1_TRACE(11);
            // ---- OWN / SRC operands of the pipeline's next tile (columns 0-191 are free: stage D has completed) --------------
            const int64_t t2 = t + 2 * (int64_t)gridDim.x;
            if (t2 < n_tiles) {
                const int g2 = (int)(t2 / NT), T2 = (int)(t2 - (int64_t)g2 * NT);
                valid = r < __ldg(tile_meta + 2 * T2);
                const int b2 = (int)((k + 1) % NBUF);
                mbar_wait(&bars->full[q][b2], (uint32_t)(((k + 1) / NBUF) & 1));
                if (OWN_BY_GATHER) {
                    mk = s1_load_mask<BF16>(smem + SM_BUF + (q * NBUF + b2) * F::SB_SIZE, r, valid, packed_mask);
                    mbar_arrive(&bars->empty[q][b2]);
                } else {
                    const float* own_row = (ASSOC && valid) ? tr_own + ((int64_t)g2 * S + __ldg(tile_rows + (int64_t)T2 * ROWS + r)) * 32 : nullptr;
                    mk = s1_own_operands<BF16, ASSOC>(smem + SM_BUF + (q * NBUF + b2) * F::SB_SIZE, r, valid, key, inv11, lane_base,
                                                      packed_mask, own_row);
                    mbar_arrive(&bars->empty[q][b2]);
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bars->opA_full[q]);
                }
                if (r == 0) S1_TRACE(19);
            }
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

static int launch_s1(const genie_plan* p, bool assoc, const float* blob, const float* pfeat, const float* msrc, const float* mask,
                     float* zc, float* va, float* vb, const float* edge_sta, const float* edge_src, const float* tr_own,
                     const float* mask_out, bool export_halo, cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    const int64_t n_tiles = (int64_t)g.n_sta_tiles * (g.n_grid_owned > 0 ? g.n_grid_owned : g.n_grid);
    static PerDeviceOnce attr_set;
    if (attr_set.need()) {
#define GENIE_S1_ATTR(E, B, A, X)                                                                                    \
    GENIE_CUDA_CHECK(cudaFuncSetAttribute(da_layer1_s_kernel<E, B, A, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          Fmt<B>::SM_TOTAL));
        GENIE_S1_ATTR(false, false, false, false) GENIE_S1_ATTR(true, false, false, false) GENIE_S1_ATTR(false, true, false, false)
        GENIE_S1_ATTR(true, true, false, false) GENIE_S1_ATTR(false, false, true, false) GENIE_S1_ATTR(true, false, true, false)
        GENIE_S1_ATTR(false, false, false, true) GENIE_S1_ATTR(true, false, false, true) GENIE_S1_ATTR(false, true, false, true)
        GENIE_S1_ATTR(true, true, false, true)
#undef GENIE_S1_ATTR
        attr_set.mark();
    }
    const int64_t grid = n_tiles < p->sm_count ? n_tiles : p->sm_count;
    const bool edge = edge_sta != nullptr, bf = !assoc && p->storage == GENIE_STORAGE_BF16;
    const bool exp = export_halo && p->exp_ptr != nullptr;
    TimedLaunch tl(assoc ? KID_ASSOC_LAYER1 : KID_DA_LAYER1_S, st);
#define GENIE_S1_LAUNCH(E, B, A, X)                                                                                            \
    if (edge == E && bf == B && assoc == A && exp == X)                                                                         \
        da_layer1_s_kernel<E, B, A, X><<<(unsigned)grid, S1_THREADS, Fmt<B>::SM_TOTAL, st>>>(                                   \
            blob, pfeat, msrc, mask, zc, va, vb, g.n_sta, g.n_sta_tiles, g.sta_tile_rows, g.sta_tile_meta, g.sta_tile_nbr,      \
            g.sta_tile_invdeg, n_tiles, edge_sta, edge_src, tr_own, mask_out, p->exp_ptr, p->exp_peer,                              \
            p->exp_row, p->peer_base, g_s1_trace, g_s1_trace_tiles, g_s1_trace_start);
    GENIE_S1_LAUNCH(false, false, false, false) GENIE_S1_LAUNCH(true, false, false, false) GENIE_S1_LAUNCH(false, true, false, false)
    GENIE_S1_LAUNCH(true, true, false, false) GENIE_S1_LAUNCH(false, false, true, false) GENIE_S1_LAUNCH(true, false, true, false)
    GENIE_S1_LAUNCH(false, false, false, true) GENIE_S1_LAUNCH(true, false, false, true) GENIE_S1_LAUNCH(false, true, false, true)
    GENIE_S1_LAUNCH(true, true, false, true)
#undef GENIE_S1_LAUNCH
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_da_layer1_s(const genie_plan* p, const float* packed, const float* pfeat, const float* msrc,
                       const float* mask, float* zc, float* va, float* vb, cudaStream_t st) {
    return launch_s1(p, false, packed + T2_BASE, pfeat, msrc, mask, zc, va, vb, p->edge_sta, p->edge_src, nullptr, nullptr, true, st);
}

// Layer 1 of the association phase on a plan with tiling tables.  blob: T2_FLOATS + 96 floats built by launch_assoc_pack_t2;
// a1 = PReLU(l1_t1_1 tr) rows, msrc = mean over source neighbours of PReLU(l1_t2_1 tr), mask [P,4], mask_out [G].
int launch_assoc_layer1_s(const genie_plan* p, const float* blob, const float* tr, const float* a1, const float* msrc,
                          const float* mask, const float* mask_out, float* zc, float* va, float* vb, cudaStream_t st) {
    return launch_s1(p, true, blob, a1, msrc, mask, zc, va, vb, p->assoc_edge_sta, p->assoc_edge_src, tr, mask_out, false, st);
}
