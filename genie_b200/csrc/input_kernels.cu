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

namespace {

__global__ void input_series_kernel(const genie_input_params_t prm, const double* __restrict__ picks, int64_t n_picks,
                                    const int32_t* __restrict__ sta_perm, float* __restrict__ series) {
    const int n_off = 2 * prm.n_extra + 1;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = tid / n_off;
    if (p >= n_picks) return;
    const int off = (int)(tid - p * n_off) - prm.n_extra;
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
    const long long bp = (long long)__ddiv_rn(__dsub_rn(__dadd_rn((double)tt.x, prm.t0), prm.ref0), prm.dt);
    const long long bs = (long long)__ddiv_rn(__dsub_rn(__dadd_rn((double)tt.y, prm.t0), prm.ref0), prm.dt);
    const float2* sr = reinterpret_cast<const float2*>(series) + (int64_t)s * prm.n_ts;      // [bin] = (P series, S series)
    const bool okp = bp > 0 && bp < (long long)prm.n_ts - 1;
    const bool oks = bs > 0 && bs < (long long)prm.n_ts - 1;
    const float2 at_p = okp ? __ldg(sr + bp) : make_float2(0.f, 0.f);
    const float2 at_s = oks ? __ldg(sr + bs) : make_float2(0.f, 0.f);
    const float pp = at_p.x;   // P series at the P bin
    const float sp_ = at_p.y;  // S series at the P bin
    const float ps = at_s.x;   // P series at the S bin
    const float ss_ = at_s.y;  // S series at the S bin
    float4 o;
    o.x = fmaxf(pp, sp_);
    o.y = fmaxf(ps, ss_);
    o.z = pp;
    o.w = ss_;
    float4 m;
    m.x = fabsf(o.x) > 0.01f ? 1.f : 0.f;
    m.y = fabsf(o.y) > 0.01f ? 1.f : 0.f;
    m.z = fabsf(o.z) > 0.01f ? 1.f : 0.f;
    m.w = fabsf(o.w) > 0.01f ? 1.f : 0.f;
    __stcs(reinterpret_cast<float4*>(slice_out) + i, o);     // read once, by the next kernel
    reinterpret_cast<float4*>(mask_out)[i] = m;
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

int launch_input_scatter(const genie_plan* p, const genie_input_params_t* prm, const double* picks, int64_t n_picks,
                         const int32_t* sta_perm, const int32_t* ind_use, const float* trv_times,
                         const int32_t* node_sta, const int32_t* node_grid, float* series, float* slice_out,
                         float* mask_out, int64_t* time_bin_out, cudaStream_t st) {
    const int64_t P = p->g.n_prod;
    const size_t series_bytes = (size_t)2 * prm->n_sta_use * prm->n_ts * sizeof(float);
    GENIE_CUDA_CHECK(cudaMemsetAsync(series, 0, series_bytes, st));
    if (n_picks > 0) {
        const int64_t total = n_picks * (2 * (int64_t)prm->n_extra + 1);
        const int64_t blocks = (total + 255) / 256;
        TimedLaunch tl(KID_INPUT_SERIES, st);
        input_series_kernel<<<(unsigned)blocks, 256, 0, st>>>(*prm, picks, n_picks, sta_perm, series);
        GENIE_LAUNCH_CHECK();
    }
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
