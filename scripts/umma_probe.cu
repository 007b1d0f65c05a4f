// Development probe (not part of the library): checks, on a real B200, the tcgen05 / TMEM / TMA building blocks the
// tensor-core kernels rely on — 128B-swizzled TMA tile loads, canonical no-swizzle K-major shared-memory operands,
// A operands in tensor memory, 3xTF32 accuracy.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe ...
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../genie_b200/csrc/tc_common.cuh"

using namespace tc;

constexpr int ROWS = 128, KDIM = 32, NKC = KDIM / 4;
constexpr int CS_A = 129 * 16;   // padded chunk stride of the A operand (bytes)

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ W,
                                                    float* __restrict__ out_ts, float* __restrict__ out_ss,
                                                    float* __restrict__ out_1p, int row0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Xs = reinterpret_cast<float*>(smem);                                  // 16 KB, TMA destination
    unsigned char* Ahi = smem + 16384;
    unsigned char* Alo = Ahi + NKC * CS_A;
    unsigned char* Bhi = Alo + NKC * CS_A;
    unsigned char* Blo = Bhi + NKC * N * 16;
    __shared__ uint64_t tma_bar, mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        mbar_init(&tma_bar, 1);
        mbar_init(&mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    if (t == 0) {
        mbar_arrive_expect_tx(&tma_bar, 16384);
        tma_load_2d(Xs, &tmap, 0, row0, &tma_bar);
    }
    mbar_wait(&tma_bar, 0);
    // row t, logical chunk c lives at physical chunk c ^ (t & 7)
    float hi[32], lo[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(smem + t * 128 + ((c ^ (t & 7)) << 4));
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            hi[4 * c + q] = tf32_hi(x[q]);
            lo[4 * c + q] = x[q] - hi[4 * c + q];
        }
        *reinterpret_cast<float4*>(Ahi + c * CS_A + t * 16) = make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<float4*>(Alo + c * CS_A + t * 16) = make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    }
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    {
        float a[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = hi[16 * h + i];
            tmem_st16(lane_base + 16 * h, a);
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = lo[16 * h + i];
            tmem_st16(lane_base + 32 + 16 * h, a);
        }
    }
    if (t < N) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(W + t * KDIM + 4 * c);
            float4 h4 = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            float4 l4 = make_float4(v.x - h4.x, v.y - h4.y, v.z - h4.z, v.w - h4.w);
            *reinterpret_cast<float4*>(Bhi + c * N * 16 + t * 16) = h4;
            *reinterpret_cast<float4*>(Blo + c * N * 16 + t * 16) = l4;
        }
    }
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_tf32(128, N);
        const uint32_t d_ts = tm + 64, d_ss = tm + 64 + N, d_1p = tm + 64 + 2 * N;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_col = (pass == 1) ? 32 : 0;
            const unsigned char* As = (pass == 1) ? Alo : Ahi;
            const unsigned char* Bs = (pass == 2) ? Blo : Bhi;
            for (int ks = 0; ks < KDIM / 8; ++ks) {
                const uint64_t bd = umma_desc_kmajor(smem_u32(Bs + ks * 2 * N * 16), N * 16, 128);
                const uint64_t ad = umma_desc_kmajor(smem_u32(As + ks * 2 * CS_A), CS_A, 128);
                const uint32_t acc = (pass | ks) ? 1u : 0u;
                umma_tf32_ts(d_ts, tm + a_col + ks * 8, bd, idesc, acc);
                umma_tf32_ss(d_ss, ad, bd, idesc, acc);
                if (pass == 0) umma_tf32_ss(d_1p, ad, bd, idesc, acc);
            }
        }
        umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + 64 + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out_ts[t * N + c0 + i] = v[i];
        tmem_ld16(lane_base + 64 + N + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out_ss[t * N + c0 + i] = v[i];
        tmem_ld16(lane_base + 64 + 2 * N + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out_1p[t * N + c0 + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                                   \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

template <int N>
int run(EncodeFn encode) {
    const int P = 1000, row0 = 300;
    std::vector<float> X((size_t)P * 32), W((size_t)N * 32);
    srand(1234 + N);
    for (auto& v : X) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    float *dX, *dW, *dts, *dss, *d1p;
    CK(cudaMalloc(&dX, X.size() * 4));
    CK(cudaMalloc(&dW, W.size() * 4));
    CK(cudaMalloc(&dts, ROWS * N * 4));
    CK(cudaMalloc(&dss, ROWS * N * 4));
    CK(cudaMalloc(&d1p, ROWS * N * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {32, (cuuint64_t)P};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        return 1;
    }
    const size_t smem = 16384 + 2 * NKC * CS_A + 2 * NKC * N * 16;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N><<<1, 128, smem>>>(tmap, dW, dts, dss, d1p, row0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> ts(ROWS * N), ss(ROWS * N), p1(ROWS * N);
    CK(cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(p1.data(), d1p, p1.size() * 4, cudaMemcpyDeviceToHost));
    double e_ts = 0, e_ss = 0, e_1p = 0, mx = 0;
    for (int m = 0; m < ROWS; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)X[(size_t)(row0 + m) * 32 + k] * (double)W[n * 32 + k];
            mx = fmax(mx, fabs(ref));
            e_ts = fmax(e_ts, fabs(ts[m * N + n] - ref));
            e_ss = fmax(e_ss, fabs(ss[m * N + n] - ref));
            e_1p = fmax(e_1p, fabs(p1[m * N + n] - ref));
        }
    printf("N=%3d  max|ref|=%.4f  err/max: A-in-TMEM 3xTF32 %.3e | A-in-smem 3xTF32 %.3e | single-pass (hi only) %.3e\n", N,
           mx, e_ts / mx, e_ss / mx, e_1p / mx);
    cudaFree(dX); cudaFree(dW); cudaFree(dts); cudaFree(dss); cudaFree(d1p);
    return (e_ts / mx < 1e-5 && e_ss / mx < 1e-5) ? 0 : 2;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    int rc = 0;
    rc |= run<32>((EncodeFn)fn);
    rc |= run<16>((EncodeFn)fn);
    rc |= run<96>((EncodeFn)fn);
    rc |= run<64>((EncodeFn)fn);
    printf(rc == 0 ? "PROBE OK\n" : "PROBE FAILED rc=%d\n", rc);
    return rc;
}
