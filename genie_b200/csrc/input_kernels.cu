// a1: pick window -> per-station Gaussian time series -> Slice / Mask   (process_utils.py:460-629).
//
// input_series_kernel: one thread per (pick, bin offset).  The pick's nearest bin is int((t - ref0)/dt) in fp64 with
//   truncation toward zero (numpy astype('int'), :515); bins nearest-n_extra .. nearest+n_extra receive
//   exp(-0.5 (t - ref[bin])^2 / sigma^2), evaluated in fp64 and rounded to fp32 as the reference does (:546, :563), and
//   are combined with a scatter-MAX.  The values are non-negative, so the max is an integer atomicMax on the float's bit
//   pattern: order-independent, hence deterministic.  ref[bin] = ref0 + bin*ref_step reproduces numpy.arange (two
//   separately rounded fp64 operations — no FMA contraction).
// input_gather_kernel: one thread per product node; time bin int((trv + t0 - ref0)/dt) per phase (:599, fp32 travel time
//   promoted to fp64), reads {max(P,S) series at the P bin, max(P,S) at the S bin, P series at the P bin, S series at the
//   S bin} (:605-608); first and last bin of every series read as zero (:565-568); Mask = |Slice| > 0.01 (:629).
#include "common.cuh"
#include "input.cuh"

namespace {

// One thread per (pick row of the window's range, bin offset).  The parameters and the pick range come by value or — for
// a window captured in a CUDA graph — from the device block `ws.dev`; the grid is then sized for the largest window.
__global__ void input_series_kernel(const WindowParamSrc ws, const double* __restrict__ picks,
                                    const int32_t* __restrict__ sta_perm, float* __restrict__ series) {
    int64_t lo, hi;
    const genie_input_params_t prm = load_params(ws, lo, hi);
    const int n_off = 2 * prm.n_extra + 1;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = lo + tid / n_off;
    if (p >= hi) return;
    const int off = (int)(tid % n_off) - prm.n_extra;
    const double t = picks[p * 5 + 0];
    if (!((t > (prm.t0 - 2.0 * prm.kernel_sig_t)) && (t < (prm.t0 + prm.max_t + 2.0 * prm.kernel_sig_t)))) return;
    const long long sta_abs = (long long)picks[p * 5 + 1];
    if (sta_abs < 0 || sta_abs >= prm.n_locs) return;
    const int s = sta_perm[sta_abs];
    if (s < 0) return;
    const double phd = picks[p * 5 + 4];
    int ph;
    if (phd == 0.0) ph = 0;
    else if (phd == 1.0) ph = 1;
    else return;
    const long long nearest = (long long)(__ddiv_rn(__dsub_rn(t, prm.ref0), prm.dt));
    const long long bin = nearest + off;
    if (bin <= 0 || bin >= (long long)prm.n_ts - 1) return;   // out of range, or an edge bin that is zeroed anyway
    const double ref = __dadd_rn(prm.ref0, __dmul_rn((double)bin, prm.ref_step));
    const double d = __dsub_rn(t, ref);
    const double arg = __ddiv_rn(__dmul_rn(-0.5, __dmul_rn(d, d)), __dmul_rn(prm.kernel_sig_t, prm.kernel_sig_t));
    const float v = (float)exp(arg);
    // scratch layout [station][bin][phase]: the gather kernel reads both phases of a bin with one 8-byte load
    int* cell = reinterpret_cast<int*>(series + (((int64_t)s * prm.n_ts + bin) * 2 + ph));
    atomicMax(cell, __float_as_int(v));
}

__global__ void input_gather_kernel(const genie_input_params_t prm, int mode, int S, int64_t P,
                                    const int32_t* __restrict__ ind_use, const float* __restrict__ trv,
                                    const int32_t* __restrict__ node_sta, const int32_t* __restrict__ node_grid,
                                    const float* __restrict__ series, float* __restrict__ slice_out,
                                    float* __restrict__ mask_out, int64_t* __restrict__ time_bin_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int s, g;
    if (mode == GENIE_GRAPH_CARTESIAN && node_sta == nullptr) {
        if (P < (int64_t)0x7fffffff) {             // 32-bit division: the 64-bit one costs ~100 instructions
            g = (int)((uint32_t)i / (uint32_t)S);
            s = (int)((uint32_t)i - (uint32_t)g * (uint32_t)S);
        } else {
            g = (int)(i / S);
            s = (int)(i - (int64_t)g * S);
        }
    } else {
        s = node_sta[i];
        g = node_grid[i];
    }
    const int sta_abs = ind_use[s];
    const float2 tt = *reinterpret_cast<const float2*>(trv + ((int64_t)g * prm.n_locs + sta_abs) * 2);
    long long bp, bs;
    const float4 o = input_slice_row(prm, s, tt, series, bp, bs);
    __stcs(reinterpret_cast<float4*>(slice_out) + i, o);     // read once, by the next kernel
    reinterpret_cast<float4*>(mask_out)[i] = input_mask_row(o);
    if (time_bin_out != nullptr) {
        time_bin_out[i * 2 + 0] = bp;
        time_bin_out[i * 2 + 1] = bs;
    }
}

// a1': nearest-pick features (process_utils.py:194-268).  One thread per (sample, product node).
__device__ __forceinline__ double nearest_distance(const double* __restrict__ times, int64_t n, double q) {
    int64_t lo = 0, hi = n;                      // numpy.searchsorted(times, q) (side='left'): first index with times[i] >= q
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(times + mid) < q) lo = mid + 1;
        else hi = mid;
    }
    const int64_t i0 = min(max(lo - 1, (int64_t)0), n - 1), i1 = min(max(lo, (int64_t)0), n - 1);     // :199-203
    return fmin(fabs(__dsub_rn(q, __ldg(times + i0))), fabs(__dsub_rn(q, __ldg(times + i1))));
}

__global__ void input_nearest_kernel(const genie_nearest_params_t prm, const double* __restrict__ t_all,
                                     const double* __restrict__ t_p, const double* __restrict__ t_s,
                                     const int32_t* __restrict__ ind_use, const float* __restrict__ trv,
                                     float* __restrict__ slice_out, float* __restrict__ mask_out) {
    const int64_t n_node = (int64_t)prm.n_grid * prm.n_sta_use;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_node * prm.n_batch) return;
    const int b = (int)(tid / n_node);
    const int64_t i = tid - (int64_t)b * n_node;
    const int g = (int)(i / prm.n_sta_use), s = (int)(i - (int64_t)g * prm.n_sta_use);
    const int sta_abs = __ldg(ind_use + s);
    const float2 tt = __ldg(reinterpret_cast<const float2*>(trv + ((int64_t)g * prm.n_locs + sta_abs) * 2));
    const double shift = __dmul_rn((double)sta_abs, prm.offset_per_station);
    const double qp = __dadd_rn(__dadd_rn((double)tt.x, __dmul_rn((double)b, prm.offset_per_batch)), shift);      // :194
    const double qs = __dadd_rn(__dadd_rn((double)tt.y, __dmul_rn((double)b, prm.offset_per_batch)), shift);      // :195
    const double s2 = __dmul_rn(prm.kernel_sig_t, prm.kernel_sig_t);
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (prm.n_all > 0) {
        const double dp = nearest_distance(t_all, prm.n_all, qp), ds = nearest_distance(t_all, prm.n_all, qs);
        v[0] = exp(__ddiv_rn(__dmul_rn(-0.5, __dmul_rn(dp, dp)), s2));
        v[1] = exp(__ddiv_rn(__dmul_rn(-0.5, __dmul_rn(ds, ds)), s2));
    }
    if (prm.n_p > 0) {
        const double d = nearest_distance(t_p, prm.n_p, qp);
        v[2] = exp(__ddiv_rn(__dmul_rn(-0.5, __dmul_rn(d, d)), s2));
    }
    if (prm.n_s > 0) {
        const double d = nearest_distance(t_s, prm.n_s, qs);
        v[3] = exp(__ddiv_rn(__dmul_rn(-0.5, __dmul_rn(d, d)), s2));
    }
    reinterpret_cast<float4*>(slice_out)[tid] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    reinterpret_cast<float4*>(mask_out)[tid] = make_float4(v[0] > 0.01 ? 1.f : 0.f, v[1] > 0.01 ? 1.f : 0.f,
                                                           v[2] > 0.01 ? 1.f : 0.f, v[3] > 0.01 ? 1.f : 0.f);
}

}  // namespace

int launch_input_nearest(const genie_nearest_params_t* prm, const double* t_all, const double* t_p, const double* t_s,
                         const int32_t* ind_use, const float* trv_times, float* slice_out, float* mask_out, cudaStream_t st) {
    const int64_t total = (int64_t)prm->n_grid * prm->n_sta_use * prm->n_batch;
    if (total == 0) return GENIE_OK;
    TimedLaunch tl(KID_INPUT_GATHER, st);
    input_nearest_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*prm, t_all, t_p, t_s, ind_use, trv_times,
                                                                           slice_out, mask_out);
    GENIE_LAUNCH_CHECK();
    return GENIE_OK;
}

int launch_input_series(const WindowParamSrc& ws, int64_t max_picks, const double* picks, const int32_t* sta_perm, float* series,
                        size_t series_bytes, int n_extra, cudaStream_t st) {
    GENIE_CUDA_CHECK(cudaMemsetAsync(series, 0, series_bytes, st));
    if (max_picks > 0) {
        const int64_t total = max_picks * (2 * (int64_t)n_extra + 1);
        const int64_t blocks = (total + 255) / 256;
        TimedLaunch tl(KID_INPUT_SERIES, st);
        input_series_kernel<<<(unsigned)blocks, 256, 0, st>>>(ws, picks, sta_perm, series);
        GENIE_LAUNCH_CHECK();
    }
    return GENIE_OK;
}

int launch_input_scatter(const genie_plan* p, const genie_input_params_t* prm, const double* picks, int64_t n_picks,
                         const int32_t* sta_perm, const int32_t* ind_use, const float* trv_times,
                         const int32_t* node_sta, const int32_t* node_grid, float* series, float* slice_out,
                         float* mask_out, int64_t* time_bin_out, cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    WindowParamSrc ws;
    ws.host = *prm;
    ws.dev = nullptr;
    ws.n_picks = n_picks;
    int rc = launch_input_series(ws, n_picks, picks, sta_perm, series, (size_t)2 * prm->n_sta_use * prm->n_ts * sizeof(float),
                                 prm->n_extra, st);
    if (rc) return rc;
    if (P > 0) {
        const int64_t blocks = (P + 255) / 256;
        TimedLaunch tl(KID_INPUT_GATHER, st);
        input_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(*prm, p->g.mode, p->g.n_sta, P, ind_use, trv_times,
                                                               node_sta, node_grid, series, slice_out, mask_out,
                                                               time_bin_out);
        GENIE_LAUNCH_CHECK();
    }
    return GENIE_OK;
}
