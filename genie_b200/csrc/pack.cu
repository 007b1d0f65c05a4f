// Re-lays the reference's nn.Linear / nn.PReLU parameters (module.py:53-83, 214-222, 231-241) into the packed kernel
// layout of layout.h.  One CTA; runs once per weight update.
#include "assoc_layout.h"
#include "common.cuh"

using namespace gl;

namespace {

// dst[(row0 + k) * ldd + o] = W[o * in_total + col0 + k],  k < ncols, o < nout   (W is [nout][in_total] row-major)
__device__ void pack_t(float* dst, int ldd, int row0, const float* W, int in_total, int col0, int ncols, int nout) {
    for (int idx = threadIdx.x; idx < ncols * nout; idx += blockDim.x) {
        const int k = idx / nout, o = idx - k * nout;
        dst[(row0 + k) * ldd + o] = W[o * in_total + col0 + k];
    }
}

__device__ void pack_v(float* dst, const float* b, int n) {
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) dst[idx] = b[idx];
}

// ---- tensor-core blob (layout.h: TC_*) ---------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi_part(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <class F>
__device__ void pack_umma(float* hi, float* lo, int N, int K, F value) {
    for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
        const int n = idx / K, k = idx - n * K;
        const float v = value(n, k);
        const float h = tf32_hi_part(v);
        const int at = ((k >> 2) * N + n) * 4 + (k & 3);
        hi[at] = h;
        lo[at] = v - h;
    }
}

__device__ void pack_tc(const genie_frontend_weights_t& w, float* __restrict__ t) {
    const float *W11 = w.da_l1_t1_2.weight, *W12 = w.da_l1_t2_2.weight;          // [30][64]
    const float *b11 = w.da_l1_t1_2.bias, *b12 = w.da_l1_t2_2.bias;
    const float *W21a = w.da_l2_t1_1.weight, *W22a = w.da_l2_t2_1.weight;        // [30][60]
    const float *W21b = w.da_l2_t1_2.weight, *W22b = w.da_l2_t2_2.weight;        // [15][94]
    pack_umma(t + TC_B1A_HI, t + TC_B1A_LO, 64, 40, [=](int n, int k) -> float {
        const int o = n & 31;
        const float* W = n < 32 ? W11 : W12;
        const float* b = n < 32 ? b11 : b12;
        if (o >= 30) return 0.f;
        if (k < 30) return W[o * 64 + k];
        if (k < 34) return W[o * 64 + 60 + (k - 30)];
        if (k == 34) return b[o];
        return 0.f;
    });
    pack_umma(t + TC_B1B_HI, t + TC_B1B_LO, 32, 32,
              [=](int n, int k) -> float { return (n < 30 && k < 30) ? W11[n * 64 + 30 + k] : 0.f; });
    pack_umma(t + TC_B1C_HI, t + TC_B1C_LO, 32, 32,
              [=](int n, int k) -> float { return (n < 30 && k < 30) ? W12[n * 64 + 30 + k] : 0.f; });
    pack_umma(t + TC_B2_HI, t + TC_B2_LO, 96, 64, [=](int n, int k) -> float {
        // operand column k -> input feature: tr[0:30] at 0-29, mask0,1 at 30,31, tr[30:60] at 32-61, mask2,3 at 62,63
        const int kt = k < 30 ? k : (k >= 32 && k < 62) ? k - 2 : -1;
        const int km = k == 30 ? 0 : k == 31 ? 1 : k == 62 ? 2 : k == 63 ? 3 : -1;
        if (n < 64) {
            const int o = n & 31;
            if (o >= 30 || kt < 0) return 0.f;
            return (n < 32 ? W21a : W22a)[o * 60 + kt];
        }
        const int o = (n - 64) & 15;
        if (o >= 15) return 0.f;
        const float* W = n < 80 ? W21b : W22b;
        if (kt >= 0) return W[o * 94 + kt];
        if (km >= 0) return W[o * 94 + 90 + km];
        return 0.f;
    });
    pack_umma(t + TC_B3A_HI, t + TC_B3A_LO, 16, 32,
              [=](int n, int k) -> float { return (n < 15 && k < 30) ? W21b[n * 94 + 60 + k] : 0.f; });
    pack_umma(t + TC_B3B_HI, t + TC_B3B_LO, 16, 32,
              [=](int n, int k) -> float { return (n < 15 && k < 30) ? W22b[n * 94 + 60 + k] : 0.f; });
    for (int n = threadIdx.x; n < 96; n += blockDim.x) {
        float b = 0.f;
        if (n < 30) b = w.da_l2_t1_1.bias[n];
        else if (n >= 32 && n < 62) b = w.da_l2_t2_1.bias[n - 32];
        else if (n >= 64 && n < 79) b = w.da_l2_t1_2.bias[n - 64];
        else if (n >= 80 && n < 95) b = w.da_l2_t2_2.bias[n - 80];
        t[TC_BIAS2 + n] = b;
    }
    if (threadIdx.x == 0) {
        const float a11 = w.da_activate11[0], a12 = w.da_activate12[0];
        const bool ok = a12 > 1e-3f && a12 < 1e3f && a11 > 1e-3f && a11 < 1e3f;      // false for NaN too
        t[TC_SCAL + TCS_OK] = ok ? 1.f : 0.f;
        t[TC_SCAL + TCS_A1] = w.da_activate1[0];
        t[TC_SCAL + TCS_A21] = w.da_activate21[0];
        t[TC_SCAL + TCS_A22] = w.da_activate22[0];
        t[TC_SCAL + TCS_R11] = ok ? a11 / a12 : 0.f;
        t[TC_SCAL + TCS_INV12] = ok ? 1.f / a12 : 0.f;
        t[TC_SCAL + TCS_A12] = a12;
        t[TC_SCAL + TCS_INV11] = ok ? 1.f / a11 : 0.f;
    }
}

// The k 24-31 block of an [N][K] hi matrix with rows 30 / 31 replaced by the hi / lo parts of a bias vector (layout.h T2_*).
template <class F, class B>
__device__ void pack_bias_block(float* dst, int N, F value, B bias) {
    for (int idx = threadIdx.x; idx < N * 8; idx += blockDim.x) {
        const int n = idx >> 3, k = idx & 7;
        float v;
        if (k == 6) v = tf32_hi_part(bias(n));
        else if (k == 7) v = bias(n) - tf32_hi_part(bias(n));
        else v = tf32_hi_part(value(n, 24 + k));
        dst[((k >> 2) * N + n) * 4 + (k & 3)] = v;
    }
}

__device__ void pack_t2(const genie_frontend_weights_t& w, float* __restrict__ t, const float* __restrict__ tc) {
    const float *W11 = w.da_l1_t1_2.weight, *W12 = w.da_l1_t2_2.weight;          // [30][64]
    const float *b11 = w.da_l1_t1_2.bias, *b12 = w.da_l1_t2_2.bias;
    const float *W21a = w.da_l2_t1_1.weight, *W22a = w.da_l2_t2_1.weight;        // [30][60]
    const float *W21b = w.da_l2_t1_2.weight, *W22b = w.da_l2_t2_2.weight;        // [15][94]
    auto s1a = [=](int n, int k) -> float {
        const int o = n & 31;
        const float* W = n < 32 ? W11 : W12;
        if (o >= 30) return 0.f;
        return k < 30 ? W[o * 64 + k] : W[o * 64 + 60 + (k - 30)];       // k 30, 31: mask0, mask1
    };
    pack_umma(t + T2_S1A_HI, t + T2_S1A_LO, 64, 32, s1a);
    pack_bias_block(t + T2_S1A_BIAS, 64, s1a, [=](int n) -> float {
        const int o = n & 31;
        return o < 30 ? (n < 32 ? b11 : b12)[o] : 0.f;
    });
    pack_umma(t + T2_S1B_HI, t + T2_S1B_LO, 32, 32, [=](int n, int k) -> float {
        if (n >= 30) return 0.f;
        return k < 30 ? W11[n * 64 + 30 + k] : W11[n * 64 + 62 + (k - 30)];      // k 30, 31: mask2, mask3
    });
    pack_umma(t + T2_S1C_HI, t + T2_S1C_LO, 32, 32, [=](int n, int k) -> float {
        if (n >= 30) return 0.f;
        return k < 30 ? W12[n * 64 + 30 + k] : W12[n * 64 + 62 + (k - 30)];
    });
    auto s2 = [=](int n, int k) -> float {
        // operand column k -> input feature: tr[0:30] at 0-29, mask0,1 at 30,31, tr[30:60] at 32-61, mask2,3 at 62,63
        const int kt = k < 30 ? k : (k >= 32 && k < 62) ? k - 2 : -1;
        const int km = k == 30 ? 0 : k == 31 ? 1 : k == 62 ? 2 : k == 63 ? 3 : -1;
        if (n < 64) {
            const int o = n & 31;
            if (o >= 30 || kt < 0) return 0.f;
            return (n < 32 ? W21a : W22a)[o * 60 + kt];
        }
        const int o = (n - 64) & 15;
        if (o >= 15) return 0.f;
        const float* W = n < 80 ? W21b : W22b;
        if (kt >= 0) return W[o * 94 + kt];
        if (km >= 0) return W[o * 94 + 90 + km];
        return 0.f;
    };
    pack_umma(t + T2_S2_HI, t + T2_S2_LO, 96, 64, s2);
    const float *b21a = w.da_l2_t1_1.bias, *b22a = w.da_l2_t2_1.bias, *b21b = w.da_l2_t1_2.bias, *b22b = w.da_l2_t2_2.bias;
    pack_bias_block(t + T2_S2_BIAS, 96, s2, [=](int n) -> float {
        if (n < 30) return b21a[n];
        if (n >= 32 && n < 62) return b22a[n - 32];
        if (n >= 64 && n < 79) return b21b[n - 64];
        if (n >= 80 && n < 95) return b22b[n - 80];
        return 0.f;
    });
    pack_umma(t + T2_S3A_HI, t + T2_S3A_LO, 16, 32,
              [=](int n, int k) -> float { return (n < 15 && k < 30) ? W21b[n * 94 + 60 + k] : 0.f; });
    pack_umma(t + T2_S3B_HI, t + T2_S3B_LO, 16, 32,
              [=](int n, int k) -> float { return (n < 15 && k < 30) ? W22b[n * 94 + 60 + k] : 0.f; });
    __syncthreads();                                        // pack_tc's scalars
    if (threadIdx.x < 8) t[T2_SCAL + threadIdx.x] = tc[TC_SCAL + threadIdx.x];
}

__global__ void pack_weights_kernel(const genie_frontend_weights_t w, float* __restrict__ p) {
    for (int i = threadIdx.x; i < PACKED_FLOATS; i += blockDim.x) p[i] = 0.f;
    __syncthreads();
    // DataAggregation
    pack_t(p + DA_W0, LD, 0, w.da_init_trns.weight, 8, 0, 8, 30);
    pack_v(p + DA_B0, w.da_init_trns.bias, 30);
    pack_t(p + DA_W11, LD, 0, w.da_l1_t1_2.weight, 64, 0, 64, 30);
    pack_v(p + DA_B11, w.da_l1_t1_2.bias, 30);
    pack_t(p + DA_W12, LD, 0, w.da_l1_t2_2.weight, 64, 0, 64, 30);
    pack_v(p + DA_B12, w.da_l1_t2_2.bias, 30);
    pack_t(p + DA_W21A, LD, 0, w.da_l2_t1_1.weight, 60, 0, 60, 30);
    pack_v(p + DA_B21A, w.da_l2_t1_1.bias, 30);
    pack_t(p + DA_W22A, LD, 0, w.da_l2_t2_1.weight, 60, 0, 60, 30);
    pack_v(p + DA_B22A, w.da_l2_t2_1.bias, 30);
    // l2_t*_2 : inputs are [tr(60) ‖ aggregate(30) ‖ mask(4)]
    pack_t(p + DA_WCA, LD16, 0, w.da_l2_t1_2.weight, 94, 0, 60, 15);
    pack_t(p + DA_WCA, LD16, 60, w.da_l2_t1_2.weight, 94, 90, 4, 15);
    pack_v(p + DA_BCA, w.da_l2_t1_2.bias, 15);
    pack_t(p + DA_WVA, LD16, 0, w.da_l2_t1_2.weight, 94, 60, 30, 15);
    pack_t(p + DA_WCB, LD16, 0, w.da_l2_t2_2.weight, 94, 0, 60, 15);
    pack_t(p + DA_WCB, LD16, 60, w.da_l2_t2_2.weight, 94, 90, 4, 15);
    pack_v(p + DA_BCB, w.da_l2_t2_2.bias, 15);
    pack_t(p + DA_WVB, LD16, 0, w.da_l2_t2_2.weight, 94, 60, 30, 15);
    if (threadIdx.x == 0) {
        p[DA_SLOPES + SL_A0] = w.da_activate[0];
        p[DA_SLOPES + SL_A11] = w.da_activate11[0];
        p[DA_SLOPES + SL_A12] = w.da_activate12[0];
        p[DA_SLOPES + SL_A1] = w.da_activate1[0];
        p[DA_SLOPES + SL_A21] = w.da_activate21[0];
        p[DA_SLOPES + SL_A22] = w.da_activate22[0];
        p[DA_SLOPES + SL_A2] = w.da_activate2[0];
        p[RI_SLOPES + 0] = w.ri_activate1[0];
        p[RI_SLOPES + 1] = w.ri_activate2[0];
    }
    // BipartiteGraphOperator
    pack_t(p + RI_WFC1, LD, 0, w.ri_fc1.weight, 33, 0, 33, 30);
    pack_v(p + RI_BFC1, w.ri_fc1.bias, 30);
    pack_t(p + RI_WFC2, LD16, 0, w.ri_fc2.weight, 30, 0, 30, 15);
    pack_v(p + RI_BFC2, w.ri_fc2.bias, 15);
    // SpatialAggregation 1..3 : fc1 inputs [x_j(C) ‖ pos diff(3) ‖ global(5)], fc2 inputs [x_i(C) ‖ aggregate(30)]
    for (int l = 0; l < 3; ++l) {
        const int C = l == 0 ? 15 : 30;
        float* q = p + SA_BASE + l * SA_SIZE;
        pack_t(q + SA_WX, LD, 0, w.sa[l].fc1.weight, C + 8, 0, C, 30);
        pack_t(q + SA_WPG, LD, 0, w.sa[l].fc1.weight, C + 8, C, 8, 30);
        pack_v(q + SA_B1, w.sa[l].fc1.bias, 30);
        pack_t(q + SA_W2, LD, 0, w.sa[l].fc2.weight, 30 + C, 0, C, 30);
        pack_t(q + SA_W2, LD, 30, w.sa[l].fc2.weight, 30 + C, C, 30, 30);
        pack_v(q + SA_B2, w.sa[l].fc2.bias, 30);
        pack_t(q + SA_WGL, 8, 0, w.sa[l].fglobal.weight, C, 0, C, 5);
        pack_v(q + SA_BGL, w.sa[l].fglobal.bias, 5);
        if (threadIdx.x == 0) {
            q[SA_SLOPES + 0] = w.sa[l].activate1[0];
            q[SA_SLOPES + 1] = w.sa[l].activate2[0];
            q[SA_SLOPES + 2] = w.sa[l].activate3[0];
        }
    }
    pack_tc(w, p + TC_BASE);
    pack_t2(w, p + T2_BASE, p + TC_BASE);
}

// Tensor-core blob of the association phase's layer 1 (da_s1_kernel.cu, ASSOC = true): the T2_* layout filled from the
// K-major matrices of the packed association weights (assoc_layout.h), followed by the weight columns of the source mask
// (mask channel 0 of the phase, module.py:388: mask = [mask_out | Mask]) that the epilogues add: [tr1(32) | tr2(32) | c_a(16) | c_b(16)].
__global__ void assoc_pack_t2_kernel(const float* __restrict__ ap, float* __restrict__ t) {
    for (int i = threadIdx.x; i < T2_FLOATS + 96; i += blockDim.x) t[i] = 0.f;
    __syncthreads();
    const float *W11 = ap + as::W11, *W12 = ap + as::W12;          // [65 of 68][32]: rows 0-29 tr, 30-59 mean, 60 mask_out, 61-64 Mask
    const float *W21A = ap + as::W21A, *W22A = ap + as::W22A;      // [60][32]
    const float *WCA = ap + as::WCA, *WCB = ap + as::WCB;          // [65 of 68][16]: rows 0-59 tr, 60 mask_out, 61-64 Mask
    const float *WVA = ap + as::WVA, *WVB = ap + as::WVB;          // [30][16]
    auto s1a = [=](int n, int k) -> float {
        const int o = n & 31;
        if (o >= 30) return 0.f;
        const float* W = n < 32 ? W11 : W12;
        return k < 30 ? W[k * 32 + o] : W[(61 + (k - 30)) * 32 + o];         // k 30, 31: Mask0, Mask1
    };
    pack_umma(t + T2_S1A_HI, t + T2_S1A_LO, 64, 32, s1a);
    pack_bias_block(t + T2_S1A_BIAS, 64, s1a, [=](int n) -> float {
        const int o = n & 31;
        return o < 30 ? ap[(n < 32 ? as::B11 : as::B12) + o] : 0.f;
    });
    pack_umma(t + T2_S1B_HI, t + T2_S1B_LO, 32, 32, [=](int n, int k) -> float {
        if (n >= 30) return 0.f;
        return k < 30 ? W11[(30 + k) * 32 + n] : W11[(63 + (k - 30)) * 32 + n];      // k 30, 31: Mask2, Mask3
    });
    pack_umma(t + T2_S1C_HI, t + T2_S1C_LO, 32, 32, [=](int n, int k) -> float {
        if (n >= 30) return 0.f;
        return k < 30 ? W12[(30 + k) * 32 + n] : W12[(63 + (k - 30)) * 32 + n];
    });
    auto s2 = [=](int n, int k) -> float {
        const int kt = k < 30 ? k : (k >= 32 && k < 62) ? k - 2 : -1;
        const int km = k == 30 ? 0 : k == 31 ? 1 : k == 62 ? 2 : k == 63 ? 3 : -1;
        if (n < 64) {
            const int o = n & 31;
            if (o >= 30 || kt < 0) return 0.f;
            return (n < 32 ? W21A : W22A)[kt * 32 + o];
        }
        const int o = (n - 64) & 15;
        if (o >= 15) return 0.f;
        const float* W = n < 80 ? WCA : WCB;
        if (kt >= 0) return W[kt * 16 + o];
        if (km >= 0) return W[(61 + km) * 16 + o];
        return 0.f;
    };
    pack_umma(t + T2_S2_HI, t + T2_S2_LO, 96, 64, s2);
    pack_bias_block(t + T2_S2_BIAS, 96, s2, [=](int n) -> float {
        if (n < 30) return ap[as::B21A + n];
        if (n >= 32 && n < 62) return ap[as::B22A + n - 32];
        if (n >= 64 && n < 79) return ap[as::BCA + n - 64];
        if (n >= 80 && n < 95) return ap[as::BCB + n - 80];
        return 0.f;
    });
    pack_umma(t + T2_S3A_HI, t + T2_S3A_LO, 16, 32, [=](int n, int k) -> float { return (n < 15 && k < 30) ? WVA[k * 16 + n] : 0.f; });
    pack_umma(t + T2_S3B_HI, t + T2_S3B_LO, 16, 32, [=](int n, int k) -> float { return (n < 15 && k < 30) ? WVB[k * 16 + n] : 0.f; });
    if (threadIdx.x == 0) {
        t[T2_SCAL + TCS_OK] = 1.f;
        t[T2_SCAL + TCS_A1] = ap[as::SL + as::SL_A1];
        t[T2_SCAL + TCS_A21] = ap[as::SL + as::SL_A21];
        t[T2_SCAL + TCS_A22] = ap[as::SL + as::SL_A22];
    }
    for (int n = threadIdx.x; n < 96; n += blockDim.x) {
        float v = 0.f;
        if (n < 64) {
            const int o = n & 31;
            if (o < 30) v = (n < 32 ? W11 : W12)[60 * 32 + o];
        } else {
            const int o = (n - 64) & 15;
            if (o < 15) v = (n < 80 ? WCA : WCB)[60 * 16 + o];
        }
        t[T2_FLOATS + n] = v;
    }
}

}  // namespace

int launch_pack_weights(const genie_frontend_weights_t* w, float* packed, cudaStream_t st) {
    TimedLaunch tl(KID_PACK, st);
    pack_weights_kernel<<<1, 256, 0, st>>>(*w, packed);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_assoc_pack_t2(const float* assoc_packed, float* blob, cudaStream_t st) {
    assoc_pack_t2_kernel<<<1, 256, 0, st>>>(assoc_packed, blob);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
