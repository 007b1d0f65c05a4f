// a1 device helpers shared by input_kernels.cu (Slice / Mask materialised) and da_kernels.cu (a1 fused into layer 0).
#pragma once

#include "common.cuh"

// Slice row of one product node (process_utils.py:599-614): the node's station `s` (index into ind_use) and its fp32
// travel times `tt` (P, S).  Time bin int((trv + t0 - ref0)/dt) per phase in fp64 (fp32 travel time promoted, :599),
// truncated toward zero as numpy's astype('int'); features {max(P,S) series at the P bin, max(P,S) at the S bin, P series
// at the P bin, S series at the S bin} (:605-608); first and last bin of every series read as zero (:565-568).
// use_sign (:610-614): every feature is multiplied by sign(e[i] - e[i+1]) of the series it was read from (the reference's
// torch.sign(-diff(e)) on the flattened [station][bin] array; at a station's last bin the feature itself is zero).
//   series layout [station][bin][phase] (float2 per bin).
__device__ __forceinline__ float sign_of(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float4 input_slice_row(const genie_input_params_t& prm, int s, float2 tt,
                                                  const float* __restrict__ series, long long& bp, long long& bs) {
    bp = (long long)__ddiv_rn(__dsub_rn(__dadd_rn((double)tt.x, prm.t0), prm.ref0), prm.dt);
    bs = (long long)__ddiv_rn(__dsub_rn(__dadd_rn((double)tt.y, prm.t0), prm.ref0), prm.dt);
    const float2* sr = reinterpret_cast<const float2*>(series) + (int64_t)s * prm.n_ts;      // [bin] = (P series, S series)
    const long long last = (long long)prm.n_ts - 1;
    const bool okp = bp > 0 && bp < last;
    const bool oks = bs > 0 && bs < last;
    const float2 at_p = okp ? __ldg(sr + bp) : make_float2(0.f, 0.f);
    const float2 at_s = oks ? __ldg(sr + bs) : make_float2(0.f, 0.f);
    float4 o;
    o.x = fmaxf(at_p.x, at_p.y);       // either phase at the P bin
    o.y = fmaxf(at_s.x, at_s.y);       // either phase at the S bin
    o.z = at_p.x;                      // P series at the P bin
    o.w = at_s.y;                      // S series at the S bin
    if (prm.use_sign_input) {
        // the next bin of the same series; bin `last` (and anything outside) reads as zero like every edge bin
        const float2 nx_p = (okp && bp + 1 < last) ? __ldg(sr + bp + 1) : make_float2(0.f, 0.f);
        const float2 nx_s = (oks && bs + 1 < last) ? __ldg(sr + bs + 1) : make_float2(0.f, 0.f);
        o.x *= sign_of(o.x - fmaxf(nx_p.x, nx_p.y));
        o.y *= sign_of(o.y - fmaxf(nx_s.x, nx_s.y));
        o.z *= sign_of(o.z - nx_p.x);
        o.w *= sign_of(o.w - nx_s.y);
    }
    return o;
}

__device__ __forceinline__ float4 input_mask_row(const float4& o) {          // :629  Mask = |Slice| > 0.01
    return make_float4(fabsf(o.x) > 0.01f ? 1.f : 0.f, fabsf(o.y) > 0.01f ? 1.f : 0.f, fabsf(o.z) > 0.01f ? 1.f : 0.f,
                       fabsf(o.w) > 0.01f ? 1.f : 0.f);
}

// The four 0/1 mask values as one exactly representable float (bit c = mask channel c): rides in padding channel 30 of the
// layer-0 feature rows, so the station-pass kernel gets the mask with the row it stages anyway.
__device__ __forceinline__ float pack_mask(const float4& m) { return m.x + 2.f * m.y + 4.f * m.z + 8.f * m.w; }
__device__ __forceinline__ float4 unpack_mask(float packed) {
    const int b = (int)packed;
    return make_float4((b & 1) ? 1.f : 0.f, (b & 2) ? 1.f : 0.f, (b & 4) ? 1.f : 0.f, (b & 8) ? 1.f : 0.f);
}

// Per-window parameters either by value (host struct) or from device memory (CUDA-graph replay: the graph's nodes are
// frozen, only the contents of the device block change between replays).
struct WindowParamSrc {
    genie_input_params_t host;
    const genie_window_params_t* dev;      // NULL: use `host`, all picks [0, n_picks)
    int64_t n_picks;
};
__device__ __forceinline__ genie_input_params_t load_params(const WindowParamSrc& w, int64_t& pick_lo, int64_t& pick_hi) {
    if (w.dev == nullptr) {
        pick_lo = 0;
        pick_hi = w.n_picks;
        return w.host;
    }
    pick_lo = w.dev->pick_lo;
    pick_hi = w.dev->pick_hi;
    return w.dev->prm;
}
