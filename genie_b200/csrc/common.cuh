// Device helpers shared by the front-end kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ float prelu(float x, float a) { return x >= 0.f ? x : a * x; }

// acc[0..29] += a * wrow[0..29]; wrow is a 32-float, 16-byte aligned shared-memory row (same address for all lanes of
// a warp -> broadcast, no bank conflicts).
__device__ __forceinline__ void fma_row30(float (&acc)[30], float a, const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        float4 w = w4[c];
        acc[4 * c + 0] = fmaf(a, w.x, acc[4 * c + 0]);
        acc[4 * c + 1] = fmaf(a, w.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(a, w.z, acc[4 * c + 2]);
        acc[4 * c + 3] = fmaf(a, w.w, acc[4 * c + 3]);
    }
    float2 w = *reinterpret_cast<const float2*>(wrow + 28);
    acc[28] = fmaf(a, w.x, acc[28]);
    acc[29] = fmaf(a, w.y, acc[29]);
}

// Packed pair of fp32 values for Blackwell's two-wide FFMA2 (fma.rn.f32x2): acc += {a, a} * {w0, w1}.  With a scalar `a`
// and weights from the constant bank ptxas emits FFMA2 R, Ra.F32, URw.F32x2, R: one issue slot for two FMAs.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(f32x2_t& acc, float a, float w0, float w1) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pack2(a, a)), "l"(pack2(w0, w1)));
}

__device__ __forceinline__ void fadd2(f32x2_t& acc, f32x2_t v) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(v)); }
// Two-wide accumulation of 16-byte chunks: acc += v (four floats, two FADD2).
struct f32x4_t {
    f32x2_t lo, hi;
};
__device__ __forceinline__ void fadd4(f32x4_t& acc, const float4& v) {
    fadd2(acc.lo, pack2(v.x, v.y));
    fadd2(acc.hi, pack2(v.z, v.w));
}
__device__ __forceinline__ float4 to_float4(const f32x4_t& a) {
    float4 r;
    unpack2(a.lo, r.x, r.y);
    unpack2(a.hi, r.z, r.w);
    return r;
}

// acc[0..15] += a * wrow[0..15] (16-float aligned row).
__device__ __forceinline__ void fma_row16(float (&acc)[16], float a, const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float4 w = w4[c];
        acc[4 * c + 0] = fmaf(a, w.x, acc[4 * c + 0]);
        acc[4 * c + 1] = fmaf(a, w.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(a, w.z, acc[4 * c + 2]);
        acc[4 * c + 3] = fmaf(a, w.w, acc[4 * c + 3]);
    }
}

// Neighbour range of one edge type for product node i.
struct NbrRange {
    int64_t beg, end;   // positions in the col array
    int64_t mul, add;   // neighbour node id = col * mul + add
};

__device__ __forceinline__ void node_ranges(const GraphView& gv, int64_t i, NbrRange& sta, NbrRange& src) {
    if (gv.mode == GENIE_GRAPH_CARTESIAN) {
        const int64_t g = i / gv.S;
        const int64_t s = i - g * gv.S;
        sta.beg = gv.sta_rowptr[s];
        sta.end = gv.sta_rowptr[s + 1];
        sta.mul = 1;
        sta.add = g * gv.S;
        src.beg = gv.src_rowptr[g];
        src.end = gv.src_rowptr[g + 1];
        src.mul = gv.S;
        src.add = s;
    } else {
        sta.beg = gv.sta_rowptr[i];
        sta.end = gv.sta_rowptr[i + 1];
        sta.mul = 1;
        sta.add = 0;
        src.beg = gv.src_rowptr[i];
        src.end = gv.src_rowptr[i + 1];
        src.mul = 1;
        src.add = 0;
    }
}

__device__ __forceinline__ int node_grid(const GraphView& gv, int64_t i) {
    return gv.mode == GENIE_GRAPH_CARTESIAN ? (int)(i / gv.S) : gv.prod_grid[i];
}
