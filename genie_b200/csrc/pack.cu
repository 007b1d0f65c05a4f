// Re-lays the reference's nn.Linear / nn.PReLU parameters (module.py:53-83, 214-222, 231-241) into the packed kernel
// layout of layout.h.  One CTA; runs once per weight update.
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
}

}  // namespace

int launch_pack_weights(const genie_frontend_weights_t* w, float* packed, cudaStream_t st) {
    TimedLaunch tl(KID_PACK, st);
    pack_weights_kernel<<<1, 256, 0, st>>>(*w, packed);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
