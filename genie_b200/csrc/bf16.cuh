// bf16 row storage of the fast mode (genie_plan_set_storage, GENIE_STORAGE_BF16): the gathered node-feature rows — the
// layer-0 features p, their source-neighbour mean msrc, the layer-2 source messages v_b and their mean — are kept as bf16
// in HBM and in shared memory (half the bytes on every gather path, which is what bounds the kernels); all arithmetic
// stays fp32 (sums, means, 3xTF32 tensor-core stages).  A 16-byte chunk holds 8 consecutive channels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// fp32 -> bf16 bits, round to nearest even (finite inputs; the feature rows hold no NaN / Inf)
__device__ __forceinline__ uint32_t bf16_bits(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ uint32_t bf16_pack2(float lo, float hi) { return bf16_bits(lo) | (bf16_bits(hi) << 16); }
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 8 floats <-> one 16-byte chunk
__device__ __forceinline__ uint4 bf16_pack8(const float (&v)[8]) {
    return make_uint4(bf16_pack2(v[0], v[1]), bf16_pack2(v[2], v[3]), bf16_pack2(v[4], v[5]), bf16_pack2(v[6], v[7]));
}
__device__ __forceinline__ void bf16_unpack8(const uint4& c, float (&v)[8]) {
    v[0] = bf16_lo(c.x); v[1] = bf16_hi(c.x); v[2] = bf16_lo(c.y); v[3] = bf16_hi(c.y);
    v[4] = bf16_lo(c.z); v[5] = bf16_hi(c.z); v[6] = bf16_lo(c.w); v[7] = bf16_hi(c.w);
}
