// Warp gathers shared by the generic DataAggregation kernels (da_kernels.cu) and the association-phase kernels
// (assoc_kernels.cu): one warp per target node, lanes = channels, every neighbour row one coalesced request.
#pragma once

#include "common.cuh"

// --------------------------------------------------------------------------------------------------------------------
// gathers (one warp per target node, lanes = channels)
// --------------------------------------------------------------------------------------------------------------------

// mean over the neighbours of PReLU(X[j][lane], slope); X rows are 32 floats.  All 32 lanes participate.
__device__ __forceinline__ float gather_mean32(const float* __restrict__ X, const NbrRange r,
                                               const int32_t* __restrict__ col, float slope, int lane) {
    float acc = 0.f;
    for (int64_t e0 = r.beg; e0 < r.end; e0 += 32) {
        const int cnt = (int)min((int64_t)32, r.end - e0);
        const int32_t c = lane < cnt ? col[e0 + lane] : 0;
        for (int u = 0; u < cnt; u += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t j = (int64_t)__shfl_sync(FULL_MASK, c, (u + q) & 31) * r.mul + r.add;
                v[q] = (u + q) < cnt ? X[j * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += prelu(v[q], slope);
        }
    }
    const int64_t deg = r.end - r.beg;
    return deg > 0 ? acc / (float)deg : 0.f;
}

// Two 16-wide gathers at once: lanes 0-15 average rows of VA over range ra, lanes 16-31 rows of VB over range rb.
__device__ __forceinline__ float gather_mean16x2(const float* __restrict__ VA, const float* __restrict__ VB,
                                                 const NbrRange ra, const NbrRange rb,
                                                 const int32_t* __restrict__ cola, const int32_t* __restrict__ colb,
                                                 int lane) {
    const int half = lane >> 4, l = lane & 15;
    const NbrRange r = half ? rb : ra;
    const float* __restrict__ V = half ? VB : VA;
    const int32_t* __restrict__ col = half ? colb : cola;
    const int64_t deg = r.end - r.beg;
    const int64_t dmax = max(ra.end - ra.beg, rb.end - rb.beg);
    float acc = 0.f;
    for (int64_t e0 = 0; e0 < dmax; e0 += 16) {
        const int cnt = (int)max((int64_t)0, min((int64_t)16, deg - e0));
        const int32_t c = l < cnt ? col[r.beg + e0 + l] : 0;
        const int cmax = (int)min((int64_t)16, dmax - e0);
        for (int u = 0; u < cmax; u += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t j = (int64_t)__shfl_sync(FULL_MASK, c, (lane & 16) | ((u + q) & 15)) * r.mul + r.add;
                v[q] = (u + q) < cnt ? V[j * gl::LD_V + l] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += v[q];
        }
    }
    return deg > 0 ? acc / (float)deg : 0.f;
}

