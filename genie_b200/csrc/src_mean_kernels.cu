// Source pass of DataAggregation (module.py:91, 95: `propagate(A_in_src, x=...)` with mean aggregation) — CARTESIAN graphs.
//
//   out[g, s, :] = 1/deg(g) * sum_{g' in N_src(g)} X[g', s, :]          (0 when g has no in-edges: PyG's mean of nothing)
//
// The product edge (g,s) <- (g',s) keeps the station, so for a slab of stations the pass is a sparse G x G matrix applied
// to dense rows.  The kernel walks compact groups of grid nodes (genie_graph_desc_t.grid_grp_*, built by recursive
// bisection of the grid graph): the source-neighbour sets of the <= 256 nodes of a group overlap heavily (union ~600 rows
// instead of 256 x 15), and ONE 1024-thread CTA per SM works on one (group, 8-station slab) tile at a time, so the union
// of neighbour rows is re-read through that SM's L1 / the L2 while the group's nodes are summed: DRAM sees every row once
// (tiles are visited slab-major: all groups of one slab, G x 1 KB = 51 MB at C4, stay L2 resident).  Group size and slab
// width were swept on the GPU; the L1 data pipe is 88 % / 95 % busy (profiles/r1zl_ncu_full_summary.md).
// Lanes map to 16-byte chunks of a row (8 lanes per 128-byte row), so every request is a fully used 128-byte line; there
// is no shared memory, no barrier and no atomics: warps drift apart freely, and the leaders pull the next tile's rows
// into L1 while the stragglers finish.
#include "bf16.cuh"
#include "common.cuh"
#include <climits>

using namespace gl;

namespace {

constexpr int SM_THREADS = 1024;

// W = floats per row (32: layer-0 features p, 16: layer-2 messages v_b); SB = stations per slab (even)
template <int W, int SB>
__global__ void __launch_bounds__(SM_THREADS, 1)
    src_mean_kernel(const float* __restrict__ X, float* __restrict__ out, int S, const int64_t* __restrict__ rowptr,
                    const int32_t* __restrict__ col, const int32_t* __restrict__ grp_ptr,
                    const int32_t* __restrict__ grp_nodes, int n_groups, int n_slabs, const float* __restrict__ gate,
                    const float* __restrict__ halo, int n_owned) {
    if (gate != nullptr && *gate == 0.f) return;     // the one-pass kernels run instead (layout.h TCS_OK)
    constexpr int LPR = W / 4;                       // lanes (16-byte chunks) per row
    constexpr int PAIRS = SB / 2;                    // every lane sums the same chunk of TWO stations (s and s + PAIRS)
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    float4* __restrict__ O4 = reinterpret_cast<float4*>(out);
    const uint32_t gstride = (uint32_t)S * LPR;      // float4 units between consecutive grid nodes (P * LPR < 2^32 checked)
    // grid-sharded plans: rows of halo grid nodes (ids >= n_owned; n_owned = INT_MAX otherwise) live in the landing buffer
    const float4* __restrict__ H4 = reinterpret_cast<const float4*>(halo);
    const int64_t n_tiles = (int64_t)n_groups * n_slabs;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int slab = (int)(t / n_groups);
        const int grp = (int)(t - (int64_t)slab * n_groups);
        const int gbeg = __ldg(grp_ptr + grp);
        const int gcnt = __ldg(grp_ptr + grp + 1) - gbeg;
        const int s0 = slab * SB;
        const int items = gcnt * PAIRS * LPR;
        for (int i = threadIdx.x; i < items; i += SM_THREADS) {
            const int c = i % LPR;
            const int pr = (i / LPR) % PAIRS;
            const int gl = i / (LPR * PAIRS);
            const int s = s0 + pr;                 // the lanes of one load cover CONSECUTIVE stations: 128-byte lines fully used
            if (s >= S) continue;
            const bool two = s + PAIRS < S;        // second station of this lane: PAIRS further on
            const int g = __ldg(grp_nodes + gbeg + gl);
            const int beg = (int)__ldg(rowptr + g);
            const int deg = (int)__ldg(rowptr + g + 1) - beg;
            const uint32_t off = (uint32_t)s * LPR + c;                   // chunk c of station s inside a grid node's block
            const uint32_t off2 = two ? off + PAIRS * LPR : off;          // same chunk of station s + PAIRS (clamped at the edge)
            const int32_t* __restrict__ cp = col + beg;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            int j = 0;
            for (; j + 5 <= deg; j += 5) {                                // k = 15 in-edges: three unmasked batches of five
                float4 v0[5], v1[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int cj = __ldg(cp + j + u);
                    const float4* __restrict__ B = cj >= n_owned ? H4 : X4;
                    const uint32_t base = (uint32_t)(cj >= n_owned ? cj - n_owned : cj) * gstride;
                    v0[u] = __ldg(B + (base + off));
                    v1[u] = __ldg(B + (base + off2));
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    a0.x += v0[u].x; a0.y += v0[u].y; a0.z += v0[u].z; a0.w += v0[u].w;
                    a1.x += v1[u].x; a1.y += v1[u].y; a1.z += v1[u].z; a1.w += v1[u].w;
                }
            }
            for (; j < deg; ++j) {
                const int cj = __ldg(cp + j);
                const float4* __restrict__ B = cj >= n_owned ? H4 : X4;
                const uint32_t base = (uint32_t)(cj >= n_owned ? cj - n_owned : cj) * gstride;
                const float4 v0 = __ldg(B + (base + off)), v1 = __ldg(B + (base + off2));
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            }
            const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
            const uint32_t o = (uint32_t)g * gstride + off;
            __stcs(O4 + o, make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv));
            if (two) __stcs(O4 + (o + PAIRS * LPR), make_float4(a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv));
        }
    }
}

// The same pass over bf16 rows (genie_plan_set_storage, GENIE_STORAGE_BF16): W channels = W / 8 chunks of 16 bytes per row,
// fp32 accumulation, bf16 result.  Half the bytes through the L1 data pipe that bounds the fp32 kernel.
template <int W, int SB>
__global__ void __launch_bounds__(SM_THREADS, 1)
    src_mean_bf16_kernel(const uint4* __restrict__ X4, uint4* __restrict__ O4, int S, const int64_t* __restrict__ rowptr,
                         const int32_t* __restrict__ col, const int32_t* __restrict__ grp_ptr,
                         const int32_t* __restrict__ grp_nodes, int n_groups, int n_slabs, const float* __restrict__ gate,
                         const uint4* __restrict__ H4, int n_owned) {
    if (gate != nullptr && *gate == 0.f) return;
    constexpr int LPR = W / 8;                       // lanes (16-byte chunks) per row
    constexpr int PAIRS = SB / 2;
    const uint32_t gstride = (uint32_t)S * LPR;      // uint4 units between consecutive grid nodes
    const int64_t n_tiles = (int64_t)n_groups * n_slabs;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int slab = (int)(t / n_groups);
        const int grp = (int)(t - (int64_t)slab * n_groups);
        const int gbeg = __ldg(grp_ptr + grp);
        const int gcnt = __ldg(grp_ptr + grp + 1) - gbeg;
        const int s0 = slab * SB;
        const int items = gcnt * PAIRS * LPR;
        for (int i = threadIdx.x; i < items; i += SM_THREADS) {
            const int c = i % LPR;
            const int pr = (i / LPR) % PAIRS;
            const int gl = i / (LPR * PAIRS);
            const int s = s0 + pr;
            if (s >= S) continue;
            const bool two = s + PAIRS < S;
            const int g = __ldg(grp_nodes + gbeg + gl);
            const int beg = (int)__ldg(rowptr + g);
            const int deg = (int)__ldg(rowptr + g + 1) - beg;
            const uint32_t off = (uint32_t)s * LPR + c;
            const uint32_t off2 = two ? off + PAIRS * LPR : off;
            const int32_t* __restrict__ cp = col + beg;
            float a0[8], a1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
            int j = 0;
            for (; j + 3 <= deg; j += 3) {                                // k = 15 in-edges: five unmasked batches of three
                uint4 v0[3], v1[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int cj = __ldg(cp + j + u);
                    const uint4* __restrict__ B = cj >= n_owned ? H4 : X4;
                    const uint32_t base = (uint32_t)(cj >= n_owned ? cj - n_owned : cj) * gstride;
                    v0[u] = __ldg(B + (base + off));
                    v1[u] = __ldg(B + (base + off2));
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    float f[8];
                    bf16_unpack8(v0[u], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) a0[e] += f[e];
                    bf16_unpack8(v1[u], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) a1[e] += f[e];
                }
            }
            for (; j < deg; ++j) {
                const int cj = __ldg(cp + j);
                const uint4* __restrict__ B = cj >= n_owned ? H4 : X4;
                const uint32_t base = (uint32_t)(cj >= n_owned ? cj - n_owned : cj) * gstride;
                float f[8];
                bf16_unpack8(__ldg(B + (base + off)), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) a0[e] += f[e];
                bf16_unpack8(__ldg(B + (base + off2)), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) a1[e] += f[e];
            }
            const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                a0[e] *= inv;
                a1[e] *= inv;
            }
            const uint32_t o = (uint32_t)g * gstride + off;
            __stcs(O4 + o, bf16_pack8(a0));
            if (two) __stcs(O4 + (o + PAIRS * LPR), bf16_pack8(a1));
        }
    }
}

}  // namespace

bool split_supported(const genie_plan* p) {
    const genie_graph_desc_t& g = p->g;
    return g.mode == GENIE_GRAPH_CARTESIAN && g.n_sta_tiles > 0 && g.n_grid_groups > 0 && g.sta_tile_rows &&
           g.sta_tile_meta && g.sta_tile_nbr && g.sta_tile_invdeg && g.grid_grp_ptr && g.grid_grp_nodes && g.n_prod > 0 &&
           g.n_prod < (int64_t)0x7fffff00 / 4;     // 32-bit float4 indices in the source pass
}

template <int W, int SB>
static void launch_src_mean_t(const genie_plan* p, const float* X, float* out, const float* gate, cudaStream_t st,
                              const float* halo) {
    const genie_graph_desc_t& g = p->g;
    const int n_slabs = (g.n_sta + SB - 1) / SB;
    const int64_t n_tiles = (int64_t)g.n_grid_groups * n_slabs;
    const unsigned grid = (unsigned)(n_tiles < p->sm_count ? n_tiles : p->sm_count);
    src_mean_kernel<W, SB><<<grid, SM_THREADS, 0, st>>>(X, out, g.n_sta, g.src_rowptr, g.src_col, g.grid_grp_ptr,
                                                        g.grid_grp_nodes, g.n_grid_groups, n_slabs, gate, halo,
                                                        halo ? g.n_grid_owned : INT_MAX);
}

template <int W, int SB>
static void launch_src_mean_bf16_t(const genie_plan* p, const float* X, float* out, const float* gate, cudaStream_t st,
                                   const float* halo) {
    const genie_graph_desc_t& g = p->g;
    const int n_slabs = (g.n_sta + SB - 1) / SB;
    const int64_t n_tiles = (int64_t)g.n_grid_groups * n_slabs;
    const unsigned grid = (unsigned)(n_tiles < p->sm_count ? n_tiles : p->sm_count);
    src_mean_bf16_kernel<W, SB><<<grid, SM_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(X), reinterpret_cast<uint4*>(out),
                                                             g.n_sta, g.src_rowptr, g.src_col, g.grid_grp_ptr, g.grid_grp_nodes,
                                                             g.n_grid_groups, n_slabs, gate, reinterpret_cast<const uint4*>(halo),
                                                             halo ? g.n_grid_owned : INT_MAX);
}

// X / out: fp32 rows, or (storage == GENIE_STORAGE_BF16) bf16 rows of the same channel count
int launch_src_mean(const genie_plan* p, int width, const float* X, float* out, const float* gate, cudaStream_t st,
                    int storage, const float* halo) {
    // One slab = 8 stations of every neighbour row (1024 B of 128-byte rows, 512 B of 64-byte rows): measured best on B200
    // together with groups of 256 grid nodes (sweeps of 64..4096 nodes x 256..2048 bytes, gpurun r1zd-r1zf: 5.2 -> 3.9 ms and
    // 3.6 -> 3.3 ms at C4).  bf16 rows: 16 / 32 stations, the same bytes per slab.
    const bool bf = storage == GENIE_STORAGE_BF16;
    if (width == 32) {
        TimedLaunch tl(KID_SRC_MEAN32, st);
        if (bf) launch_src_mean_bf16_t<32, 16>(p, X, out, gate, st, halo);
        else launch_src_mean_t<32, 8>(p, X, out, gate, st, halo);
    } else if (width == 16) {
        TimedLaunch tl(KID_SRC_MEAN16, st);
        if (bf) launch_src_mean_bf16_t<16, 16>(p, X, out, gate, st, halo);
        else launch_src_mean_t<16, 8>(p, X, out, gate, st, halo);
    } else {
        set_error("launch_src_mean: unsupported row width");
        return GENIE_ERR_INVALID;
    }
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
