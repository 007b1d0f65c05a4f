// Micro-probe: issue rate / completion time of many small cp.async.bulk row copies (global -> shared) on one SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o bulk_probe bulk_probe.cu ; run: ./bulk_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const float* src, int n_rows_total, int rows, int bytes, int lanes, long long* out, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    long long t_issue = 0, t_done = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        long long t0 = clock64();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * bytes) : "memory");
            __syncwarp();
            if ((int)threadIdx.x < lanes) {
                for (int r = threadIdx.x; r < rows; r += lanes) {
                    // scattered rows: pseudo-random row id per (block, it, r)
                    uint32_t h = (uint32_t)(blockIdx.x * 7919u + it * 104729u + r * 2654435761u);
                    const float* s = src + (size_t)(h % (uint32_t)n_rows_total) * 32;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(smem + (size_t)r * bytes)), "l"(s), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
                }
            }
        }
        long long t1 = clock64();
        // everyone waits
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        }
        parity ^= 1;
        long long t2 = clock64();
        if (threadIdx.x == 0) { t_issue += t1 - t0; t_done += t2 - t0; }
    }
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t_issue / iters; out[blockIdx.x * 2 + 1] = t_done / iters; }
}
int main() {
    const int n_rows_total = 1 << 22;   // 512 MB of 128-byte rows
    float* src; cudaMalloc(&src, (size_t)n_rows_total * 128); cudaMemset(src, 0, (size_t)n_rows_total * 128);
    long long* out; cudaMallocManaged(&out, 148 * 2 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int cfgs[][3] = {{256, 128, 32}, {256, 128, 8}, {256, 128, 1}, {512, 128, 32}, {64, 512, 32}, {1024, 16, 32}, {16, 8192, 16}, {256, 64, 32}};
    for (auto& c : cfgs) {
        for (int grid : {1, 148}) {
            probe<<<grid, 128, 200 * 1024>>>(src, n_rows_total, c[0], c[1], c[2], out, 20);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("rows %4d x %4d B, %2d issuing lanes, grid %3d: issue %6lld clk, all landed %6lld clk (%.1f clk/row, %.1f B/clk/SM)\n", c[0], c[1], c[2], grid,
                   out[0], out[1], (double)out[1] / c[0], (double)c[0] * c[1] / out[1]);
        }
    }
    return 0;
}
