// Brute-force k-nearest-neighbour search on the device: replaces torch_cluster.knn at its three call
// sites on the path — the station and source kNN graphs of extract_inputs_adjacencies (process_utils.py:718-719) and the
// query -> grid-node edges of SpatialAttention (module.py:282).
//
//   idx[q][0..k) = the k rows of X nearest to Y[q], nearest first           (torch_cluster.knn(x, y, k) row 1, grouped by y)
//
// One thread per query point; the X points stream through shared memory in tiles every thread of the CTA reads with the
// same address (broadcast).  Each thread keeps its K best candidates sorted in registers (fully unrolled compare-and-swap
// insertion, constant indices only: no local memory); a candidate enters only if it beats the current k-th distance, which
// after the first few tiles happens ~k ln(n / k) times per query, so the inner loop is one distance evaluation per pair.
// Distances are evaluated in fp64 from the fp32 coordinates (the differences are then exact), which reproduces the
// ordering of an fp64 k-d tree query on the same fp32 points; ties keep the lower X index.  n_x * n_y = 2.5e9 pairs at
// G = 50000 is ~25 GFLOP of fp64: a few milliseconds, once per station set.
#include "common.cuh"
#include <climits>

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;       // X points per shared-memory tile (12 KB)

template <int K>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const float* __restrict__ X, int n_x, const float* __restrict__ Y,
                                                          int n_y, int k, int64_t* __restrict__ idx_out) {
    __shared__ float sx[KNN_TILE * 3];
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool live = q < n_y;
    const double y0 = live ? (double)Y[(int64_t)q * 3 + 0] : 0.0;
    const double y1 = live ? (double)Y[(int64_t)q * 3 + 1] : 0.0;
    const double y2 = live ? (double)Y[(int64_t)q * 3 + 2] : 0.0;
    double bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = INFINITY;
        bi[j] = -1;
    }
    for (int t0 = 0; t0 < n_x; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, n_x - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += KNN_THREADS) sx[i] = X[(int64_t)t0 * 3 + i];
        __syncthreads();
        if (!live) continue;
        for (int i = 0; i < cnt; ++i) {
            const double d0 = (double)sx[3 * i] - y0, d1 = (double)sx[3 * i + 1] - y1, d2 = (double)sx[3 * i + 2] - y2;
            const double d = d0 * d0 + d1 * d1 + d2 * d2;
            if (d < bd[K - 1]) {
                bd[K - 1] = d;
                bi[K - 1] = t0 + i;
#pragma unroll
                for (int j = K - 1; j > 0; --j) {
                    if (bd[j] < bd[j - 1]) {           // strict: an equal distance stays behind the earlier index
                        const double td = bd[j];
                        bd[j] = bd[j - 1];
                        bd[j - 1] = td;
                        const int ti = bi[j];
                        bi[j] = bi[j - 1];
                        bi[j - 1] = ti;
                    }
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j < k) idx_out[(int64_t)q * k + j] = bi[j];
    }
}

// Few query points (the association sources of forward_fixed, module.py:980: a handful of queries against 50000 grid nodes —
// 2.7 ms on one thread per query): ONE CTA PER QUERY.  Thread t scans the X points t, t + 128, ... with the same sorted
// insertion, then the 128 sorted lists are merged head by head: k rounds of a block-wide lexicographic (distance, index)
// minimum, the winning thread drops its head.  Same result as the thread-per-query kernel: the k smallest by (distance, index).
template <int K>
__global__ void __launch_bounds__(KNN_THREADS) knn_few_kernel(const float* __restrict__ X, int n_x, const float* __restrict__ Y,
                                                              int k, int64_t* __restrict__ idx_out) {
    __shared__ double sd[KNN_THREADS / 32];
    __shared__ int si[KNN_THREADS / 32];
    __shared__ int s_win;
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double y0 = (double)Y[(int64_t)q * 3 + 0], y1 = (double)Y[(int64_t)q * 3 + 1], y2 = (double)Y[(int64_t)q * 3 + 2];
    double bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = INFINITY;
        bi[j] = INT_MAX;
    }
    for (int i = threadIdx.x; i < n_x; i += KNN_THREADS) {
        const double d0 = (double)__ldg(X + (int64_t)i * 3) - y0, d1 = (double)__ldg(X + (int64_t)i * 3 + 1) - y1,
                     d2 = (double)__ldg(X + (int64_t)i * 3 + 2) - y2;
        const double d = d0 * d0 + d1 * d1 + d2 * d2;
        if (d < bd[K - 1]) {
            bd[K - 1] = d;
            bi[K - 1] = i;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                if (bd[j] < bd[j - 1]) {
                    const double td = bd[j];
                    bd[j] = bd[j - 1];
                    bd[j - 1] = td;
                    const int ti = bi[j];
                    bi[j] = bi[j - 1];
                    bi[j - 1] = ti;
                }
            }
        }
    }
    for (int r = 0; r < k; ++r) {
        double d = bd[0];
        int i = bi[0];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double od = __shfl_xor_sync(FULL_MASK, d, s);
            const int oi = __shfl_xor_sync(FULL_MASK, i, s);
            if (od < d || (od == d && oi < i)) {
                d = od;
                i = oi;
            }
        }
        if (lane == 0) {
            sd[warp] = d;
            si[warp] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double wd = sd[0];
            int wi = si[0];
#pragma unroll
            for (int w = 1; w < KNN_THREADS / 32; ++w)
                if (sd[w] < wd || (sd[w] == wd && si[w] < wi)) {
                    wd = sd[w];
                    wi = si[w];
                }
            s_win = wi;
            idx_out[(int64_t)q * k + r] = wi == INT_MAX ? -1 : wi;
        }
        __syncthreads();
        if (bi[0] == s_win && s_win != INT_MAX) {        // indices are unique across threads: exactly one list advances
#pragma unroll
            for (int j = 0; j < K - 1; ++j) {
                bd[j] = bd[j + 1];
                bi[j] = bi[j + 1];
            }
            bd[K - 1] = INFINITY;
            bi[K - 1] = INT_MAX;
        }
    }
}

constexpr int KNN_FEW = 2048;        // up to this many queries: one CTA per query

}  // namespace

int launch_knn(const float* x, int n_x, const float* y, int n_y, int k, int64_t* idx_out, cudaStream_t st) {
    if (n_y == 0) return GENIE_OK;
    const unsigned grid = (unsigned)((n_y + KNN_THREADS - 1) / KNN_THREADS);
    TimedLaunch tl(KID_KNN, st);
    if (n_y <= KNN_FEW && k <= 32) {
        if (k <= 8) knn_few_kernel<8><<<n_y, KNN_THREADS, 0, st>>>(x, n_x, y, k, idx_out);
        else if (k <= 12) knn_few_kernel<12><<<n_y, KNN_THREADS, 0, st>>>(x, n_x, y, k, idx_out);
        else if (k <= 16) knn_few_kernel<16><<<n_y, KNN_THREADS, 0, st>>>(x, n_x, y, k, idx_out);
        else knn_few_kernel<32><<<n_y, KNN_THREADS, 0, st>>>(x, n_x, y, k, idx_out);
        GENIE_LAUNCH_CHECK();
        return GENIE_OK;
    }
    if (k <= 8)
        knn_kernel<8><<<grid, KNN_THREADS, 0, st>>>(x, n_x, y, n_y, k, idx_out);
    else if (k <= 12)
        knn_kernel<12><<<grid, KNN_THREADS, 0, st>>>(x, n_x, y, n_y, k, idx_out);
    else if (k <= 16)
        knn_kernel<16><<<grid, KNN_THREADS, 0, st>>>(x, n_x, y, n_y, k, idx_out);
    else if (k <= 32)
        knn_kernel<32><<<grid, KNN_THREADS, 0, st>>>(x, n_x, y, n_y, k, idx_out);
    else {
        set_error("genie_knn_fwd: k must be <= 32");
        return GENIE_ERR_INVALID;
    }
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}
