// Caller-side output stacking of the streaming loop (process_continuous_days.py:797-805):
//   Out_2[:, ip_need[t]] += x[:, t, 0] / n_overlap / n_scale_x_grid        for the first n_use query times of a window
// kept on the device, so that a day of windows copies back one [Q, n_steps] array instead of 28 800 x [Q, T].
// Consecutive windows write overlapping columns, but windows are launched in stream order, and inside one launch every
// (q, t) pair has its own thread.  The columns of one window are normally distinct; when the caller's solution grid is too
// short and the nearest-index search clips two query times onto one column, numpy's fancy-index `+=` keeps only the LAST
// duplicate (read-modify-write per index, no accumulation) — the kernel does the same: a thread whose column re-appears
// at a later query time does not write.  No atomics, no race, bit-reproducible.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) stack_output_kernel(const float* __restrict__ x, int Q, int T, int n_use,
                                                           const int32_t* __restrict__ col, float scale,
                                                           float* __restrict__ out, int64_t ld_out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)Q * n_use) return;
    const int q = (int)(i / n_use), t = (int)(i - (int64_t)q * n_use);
    const int32_t c = col[t];
    for (int t2 = t + 1; t2 < n_use; ++t2)
        if (col[t2] == c) return;             // numpy: the last duplicate wins
    // the reference divides twice (/ n_overlap / n_scale_x_grid) in fp64 after .cpu(); here scale = 1 / (n_overlap * n_scale)
    out[(int64_t)q * ld_out + c] += x[(int64_t)q * T + t] * scale;
}

}  // namespace

int launch_stack_output(const float* x, int Q, int T, int n_use, const int32_t* col, float scale, float* out, int64_t ld_out,
                        cudaStream_t st) {
    const int64_t n = (int64_t)Q * n_use;
    if (n == 0) return GENIE_OK;
    TimedLaunch tl(KID_STACK, st);
    stack_output_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, Q, T, n_use, col, scale, out, ld_out);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
