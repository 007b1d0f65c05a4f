// Source pass of DataAggregation (module.py:91, 95: `propagate(A_in_src, x=...)` with mean aggregation) — CARTESIAN graphs.
//
//   out[g, s, :] = 1/deg(g) * sum_{g' in N_src(g)} X[g', s, :]          (0 when g has no in-edges: PyG's mean of nothing)
//
// The product edge (g,s) <- (g',s) keeps the station, so for a slab of stations the pass is a sparse G x G matrix applied
// to dense rows.  The kernel walks compact groups of grid nodes (genie_graph_desc_t.grid_grp_*, built by recursive
// bisection of the grid graph): the source-neighbour sets of the ~64 nodes of a group overlap heavily (union ~200 rows
// instead of 64 x 15), and ONE 1024-thread CTA per SM works on one (group, station slab) tile at a time, so the union of
// neighbour rows (<= 265 x 512 B) sits in that SM's L1 while the group's nodes re-read it: L2 sees every row ~3 times
// instead of 15, DRAM once (tiles are visited slab-major: all groups of one slab, G x 512 B = 26 MB, stay L2 resident).
// Lanes map to 16-byte chunks of a row (8 lanes per 128-byte row), so every request is a fully used 128-byte line; there
// is no shared memory, no barrier and no atomics: warps drift apart freely, and the leaders pull the next tile's rows
// into L1 while the stragglers finish.
#include "common.cuh"

using namespace gl;

namespace {

constexpr int SM_THREADS = 1024;

// W = floats per row (32: layer-0 features p, 16: layer-2 messages v_b); SB = stations per slab (even)
template <int W, int SB, bool PREFETCH>
__global__ void __launch_bounds__(SM_THREADS, 1)
    src_mean_kernel(const float* __restrict__ X, float* __restrict__ out, int S, const int64_t* __restrict__ rowptr,
                    const int32_t* __restrict__ col, const int32_t* __restrict__ grp_ptr,
                    const int32_t* __restrict__ grp_nodes, int n_groups, int n_slabs, const float* __restrict__ gate) {
    if (gate != nullptr && *gate == 0.f) return;     // the one-pass kernels run instead (layout.h TCS_OK)
    constexpr int LPR = W / 4;                       // lanes (16-byte chunks) per row
    constexpr int PAIRS = SB / 2;                    // every lane sums the same chunk of TWO consecutive stations
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    float4* __restrict__ O4 = reinterpret_cast<float4*>(out);
    const uint32_t gstride = (uint32_t)S * LPR;      // float4 units between consecutive grid nodes (P * LPR < 2^32 checked)
    const int64_t n_tiles = (int64_t)n_groups * n_slabs;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int slab = (int)(t / n_groups);
        const int grp = (int)(t - (int64_t)slab * n_groups);
        const int gbeg = __ldg(grp_ptr + grp);
        const int gcnt = __ldg(grp_ptr + grp + 1) - gbeg;
        const int s0 = slab * SB;
        const int items = gcnt * PAIRS * LPR;
        if (PREFETCH) {     // measured on B200 (r1r): 5.1 -> 8.5 ms, the extra L1 requests cost more than the misses: kept off
            // Pull the neighbour rows of this CTA's NEXT tile into L1 while this tile is summed (the union of a group's rows
            // then hits instead of stalling the warps on its first touch).  The PAIRS * LPR threads of a grid node share its
            // deg x (slab bytes / 128) lines.
            const int64_t tn = t + gridDim.x;
            if (tn < n_tiles) {
                const int slab_n = (int)(tn / n_groups);
                const int grp_n = (int)(tn - (int64_t)slab_n * n_groups);
                const int gbeg_n = __ldg(grp_ptr + grp_n);
                const int gcnt_n = __ldg(grp_ptr + grp_n + 1) - gbeg_n;
                constexpr int LINES = SB * W * 4 / 128;          // 128-byte lines of one neighbour row inside the slab
                constexpr int TPN = PAIRS * LPR;                 // threads per grid node
                const int i = threadIdx.x;
                if (i < gcnt_n * TPN) {
                    const int gl = i / TPN, sub = i - gl * TPN;
                    const int g = __ldg(grp_nodes + gbeg_n + gl);
                    const int beg = (int)__ldg(rowptr + g);
                    const int deg = (int)__ldg(rowptr + g + 1) - beg;
                    const char* base = reinterpret_cast<const char*>(X) + (size_t)slab_n * SB * W * 4;
                    for (int n = sub; n < deg * LINES; n += TPN) {
                        const int j = n / LINES, line = n - j * LINES;
                        if (slab_n * SB + line * (128 / (W * 4)) >= S) continue;     // ragged last slab
                        const uint32_t nb = (uint32_t)__ldg(col + beg + j);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)nb * gstride * 16 + line * 128));
                    }
                }
            }
        }
        for (int i = threadIdx.x; i < items; i += SM_THREADS) {
            const int c = i % LPR;
            const int pr = (i / LPR) % PAIRS;
            const int gl = i / (LPR * PAIRS);
            const int s = s0 + 2 * pr;
            if (s >= S) continue;
            const bool two = s + 1 < S;
            const int g = __ldg(grp_nodes + gbeg + gl);
            const int beg = (int)__ldg(rowptr + g);
            const int deg = (int)__ldg(rowptr + g + 1) - beg;
            const uint32_t off = (uint32_t)s * LPR + c;                   // chunk c of station s inside a grid node's block
            const uint32_t off2 = two ? off + LPR : off;                  // same chunk of station s + 1 (clamped at the edge)
            const int32_t* __restrict__ cp = col + beg;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            int j = 0;
            for (; j + 5 <= deg; j += 5) {                                // k = 15 in-edges: three unmasked batches of five
                float4 v0[5], v1[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const uint32_t base = (uint32_t)__ldg(cp + j + u) * gstride;
                    v0[u] = __ldg(X4 + (base + off));
                    v1[u] = __ldg(X4 + (base + off2));
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    a0.x += v0[u].x; a0.y += v0[u].y; a0.z += v0[u].z; a0.w += v0[u].w;
                    a1.x += v1[u].x; a1.y += v1[u].y; a1.z += v1[u].z; a1.w += v1[u].w;
                }
            }
            for (; j < deg; ++j) {
                const uint32_t base = (uint32_t)__ldg(cp + j) * gstride;
                const float4 v0 = __ldg(X4 + (base + off)), v1 = __ldg(X4 + (base + off2));
                a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            }
            const float inv = deg > 0 ? 1.f / (float)deg : 0.f;
            const uint32_t o = (uint32_t)g * gstride + off;
            __stcs(O4 + o, make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv));
            if (two) __stcs(O4 + (o + LPR), make_float4(a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv));
        }
    }
}

// ---- quad variant -------------------------------------------------------------------------------------------------------
// The nodes of a group are matched into quads whose merged neighbour list (genie_graph_desc_t.src_quad_*) holds every
// distinct row once with a 4-bit membership mask: a row shared by several nodes of the quad is loaded once and added to
// each of their accumulators (k = 15 nearest neighbours in 3-D: ~31 loads per quad instead of 60 — the L1 wavefronts, the
// bound of the pass, halve).  One warp = one quad x one slab of 512 bytes of every row (lanes = 16-byte chunks), so the
// masks are warp-uniform; a 512-thread CTA covers the <= 16 quads of one group, two CTAs per SM work on neighbouring
// groups of the same slab, and the union of a group's rows is still shared through L1.
constexpr int SQ_THREADS = 512;

__device__ __forceinline__ void masked_add(f32x4_t (&a)[4], const float4& v, uint32_t w) {
    if (w & (1u << 28)) fadd4(a[0], v);
    if (w & (2u << 28)) fadd4(a[1], v);
    if (w & (4u << 28)) fadd4(a[2], v);
    if (w & (8u << 28)) fadd4(a[3], v);
}

template <int W>
__global__ void __launch_bounds__(SQ_THREADS, 2)
    src_mean_quad_kernel(const float* __restrict__ X, float* __restrict__ out, int S, const int32_t* __restrict__ quad_ptr,
                         const uint32_t* __restrict__ quad_list, const int32_t* __restrict__ quad_nodes,
                         const float* __restrict__ quad_invdeg, const int32_t* __restrict__ grp_quad_ptr, int n_groups,
                         int n_slabs, const float* __restrict__ gate) {
    if (gate != nullptr && *gate == 0.f) return;     // the one-pass kernels run instead (layout.h TCS_OK)
    constexpr int LPR = W / 4;                       // lanes (16-byte chunks) per row
    constexpr int SB = 32 / LPR;                     // stations per slab: one warp spans 512 bytes of every row
    constexpr uint32_t ID = 0x0fffffffu;
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    float4* __restrict__ O4 = reinterpret_cast<float4*>(out);
    const uint32_t gstride = (uint32_t)S * LPR;      // float4 units between consecutive grid nodes (P * LPR < 2^32 checked)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (int64_t)n_groups * n_slabs;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int slab = (int)(t / n_groups);
        const int grp = (int)(t - (int64_t)slab * n_groups);
        const int q0 = __ldg(grp_quad_ptr + grp);
        const int nq = __ldg(grp_quad_ptr + grp + 1) - q0;
        const int s = slab * SB + lane / LPR;
        if (s >= S) continue;                        // ragged last slab
        const uint32_t off = (uint32_t)s * LPR + (lane % LPR);
        for (int qi = warp; qi < nq; qi += SQ_THREADS / 32) {
            const int quad = q0 + qi;
            int e = __ldg(quad_ptr + quad);
            const int end = __ldg(quad_ptr + quad + 1);
            f32x4_t a[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) a[b].lo = a[b].hi = 0ull;
            for (; e + 8 <= end; e += 8) {
                const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(quad_list + e));
                const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(quad_list + e + 4));
                const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(X4 + ((w[u] & ID) * gstride + off));
#pragma unroll
                for (int u = 0; u < 8; ++u) masked_add(a, v[u], w[u]);
            }
            if (e < end) {                           // segments are padded to a multiple of four entries
                const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(quad_list + e));
                const uint32_t w[4] = {w0.x, w0.y, w0.z, w0.w};
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(X4 + ((w[u] & ID) * gstride + off));
#pragma unroll
                for (int u = 0; u < 4; ++u) masked_add(a, v[u], w[u]);
            }
            const int4 nd = __ldg(reinterpret_cast<const int4*>(quad_nodes) + quad);
            const float4 inv = __ldg(reinterpret_cast<const float4*>(quad_invdeg) + quad);
            const int nds[4] = {nd.x, nd.y, nd.z, nd.w};
            const float invs[4] = {inv.x, inv.y, inv.z, inv.w};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (nds[b] < 0) continue;
                const float4 r = to_float4(a[b]);
                __stcs(O4 + ((uint32_t)nds[b] * gstride + off),
                       make_float4(r.x * invs[b], r.y * invs[b], r.z * invs[b], r.w * invs[b]));
            }
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
static void launch_src_mean_t(const genie_plan* p, const float* X, float* out, const float* gate, cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    const int n_slabs = (g.n_sta + SB - 1) / SB;
    const int64_t n_tiles = (int64_t)g.n_grid_groups * n_slabs;
    const unsigned grid = (unsigned)(n_tiles < p->sm_count ? n_tiles : p->sm_count);
    src_mean_kernel<W, SB, false><<<grid, SM_THREADS, 0, st>>>(X, out, g.n_sta, g.src_rowptr, g.src_col, g.grid_grp_ptr,
                                                        g.grid_grp_nodes, g.n_grid_groups, n_slabs, gate);
}

template <int W>
static void launch_src_mean_quad(const genie_plan* p, const float* X, float* out, const float* gate, cudaStream_t st) {
    const genie_graph_desc_t& g = p->g;
    constexpr int SB = 32 / (W / 4);
    const int n_slabs = (g.n_sta + SB - 1) / SB;
    const int64_t n_tiles = (int64_t)g.n_grid_groups * n_slabs;
    const int64_t cap = 2 * (int64_t)p->sm_count;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    src_mean_quad_kernel<W><<<grid, SQ_THREADS, 0, st>>>(X, out, g.n_sta, g.src_quad_ptr, g.src_quad_list, g.src_quad_nodes,
                                                         g.src_quad_invdeg, g.src_grp_quad_ptr, g.n_grid_groups, n_slabs, gate);
}

int launch_src_mean(const genie_plan* p, int width, const float* X, float* out, const float* gate, cudaStream_t st) {
    const genie_graph_desc_t& gd = p->g;
    const bool quads = gd.n_src_quads > 0 && gd.src_quad_ptr && gd.src_quad_list && gd.src_quad_nodes &&
                       gd.src_quad_invdeg && gd.src_grp_quad_ptr && gd.n_grid < (1 << 28);
    if (quads && (width == 32 || width == 16)) {
        TimedLaunch tl(width == 32 ? KID_SRC_MEAN32 : KID_SRC_MEAN16, st);
        if (width == 32) launch_src_mean_quad<32>(p, X, out, gate, st);
        else launch_src_mean_quad<16>(p, X, out, gate, st);
        GENIE_LAUNCH_CHECK();
        return GENIE_OK;
    }
    // one slab = 512 bytes of every neighbour row: 4 stations of 128-byte rows, 8 stations of 64-byte rows
    // (64 grid nodes x 2 station pairs x 8 lanes = 1024 items: one item per thread of the CTA)
    if (width == 32) {
        TimedLaunch tl(KID_SRC_MEAN32, st);
        launch_src_mean_t<32, 4>(p, X, out, gate, st);
    } else if (width == 16) {
        TimedLaunch tl(KID_SRC_MEAN16, st);
        launch_src_mean_t<16, 8>(p, X, out, gate, st);
    } else {
        set_error("launch_src_mean: unsupported row width");
        return GENIE_ERR_INVALID;
    }
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
