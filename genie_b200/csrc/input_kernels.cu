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

}  // namespace

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
