// Packed (kernel-side) weight layout of the front end.  All offsets are in floats and multiples of 4 (16-byte aligned
// for float4 shared-memory loads).  Every matrix is stored K-major ("transposed": row k holds the weights that multiply
// input feature k), with the output dimension padded to LD (30 -> 32) or LD16 (15 -> 16); padding is zero.
#pragma once

namespace gl {

constexpr int H = 30;      // hidden width of the reference (n_hidden, module.py:53)
constexpr int LD = 32;     // padded 30-wide output rows
constexpr int LD16 = 16;   // padded 15-wide output rows

// ---- DataAggregation (module.py:52-98) ------------------------------------------------------------------------------
constexpr int DA_W0 = 0;                        // init_trns          [8][32]   rows 0-3 Slice, 4-7 Mask
constexpr int DA_B0 = DA_W0 + 8 * LD;           //                    [32]
constexpr int DA_W11 = DA_B0 + LD;              // l1_t1_2            [64][32]  rows 0-29 tr0, 30-59 mean_sta, 60-63 Mask
constexpr int DA_B11 = DA_W11 + 64 * LD;
constexpr int DA_W12 = DA_B11 + LD;             // l1_t2_2            [64][32]  rows 0-29 tr0, 30-59 mean_src, 60-63 Mask
constexpr int DA_B12 = DA_W12 + 64 * LD;
constexpr int DA_W21A = DA_B12 + LD;            // l2_t1_1            [60][32]
constexpr int DA_B21A = DA_W21A + 60 * LD;
constexpr int DA_W22A = DA_B21A + LD;           // l2_t2_1            [60][32]
constexpr int DA_B22A = DA_W22A + 60 * LD;
constexpr int DA_WCA = DA_B22A + LD;            // l2_t1_2[:, 0:60 | 90:94]   [64][16]  rows 0-59 tr, 60-63 Mask
constexpr int DA_BCA = DA_WCA + 64 * LD16;      // l2_t1_2.bias       [16]
constexpr int DA_WCB = DA_BCA + LD16;           // l2_t2_2[:, 0:60 | 90:94]   [64][16]
constexpr int DA_BCB = DA_WCB + 64 * LD16;
constexpr int DA_WVA = DA_BCB + LD16;           // l2_t1_2[:, 60:90]  [30][16]  applied BEFORE the mean (linearity)
constexpr int DA_WVB = DA_WVA + 30 * LD16;      // l2_t2_2[:, 60:90]  [30][16]
constexpr int DA_SLOPES = DA_WVB + 30 * LD16;   // [8]: activate, 11, 12, 1, 21, 22, 2, -
constexpr int DA_END = DA_SLOPES + 8;
enum { SL_A0 = 0, SL_A11 = 1, SL_A12 = 2, SL_A1 = 3, SL_A21 = 4, SL_A22 = 5, SL_A2 = 6 };

// ---- BipartiteGraphOperator (module.py:214-229) ---------------------------------------------------------------------
constexpr int RI_WFC1 = DA_END;                 // fc1  [33][32]  rows 0-29 x_latent, 30-32 edge attr
constexpr int RI_BFC1 = RI_WFC1 + 33 * LD;
constexpr int RI_WFC2 = RI_BFC1 + LD;           // fc2  [30][16]
constexpr int RI_BFC2 = RI_WFC2 + 30 * LD16;
constexpr int RI_SLOPES = RI_BFC2 + LD16;       // [4]: activate1, activate2, -, -
constexpr int RI_END = RI_SLOPES + 4;

// ---- SpatialAggregation x3 (module.py:231-249) ----------------------------------------------------------------------
constexpr int SA_WX = 0;                        // fc1[:, 0:C]        [30][32]  rows >= C zero
constexpr int SA_WPG = SA_WX + 30 * LD;         // fc1[:, C:C+8]      [8][32]   rows 0-2 pos diff, 3-7 global
constexpr int SA_B1 = SA_WPG + 8 * LD;
constexpr int SA_W2 = SA_B1 + LD;               // fc2                [60][32]  rows 0..C-1 x_i, rows 30-59 aggregate
constexpr int SA_B2 = SA_W2 + 60 * LD;
constexpr int SA_WGL = SA_B2 + LD;              // fglobal            [30][8]   rows >= C zero
constexpr int SA_BGL = SA_WGL + 30 * 8;
constexpr int SA_SLOPES = SA_BGL + 8;           // [4]: activate1, activate2, activate3, -
constexpr int SA_SIZE = SA_SLOPES + 4;
constexpr int SA_BASE = RI_END;

constexpr int GENERIC_END = SA_BASE + 3 * SA_SIZE;

// ---- tensor-core blob of DataAggregation layer 1 (da_tc_kernels.cu) ----------------------------------------------------
// Every B operand of tcgen05.mma is stored [N rows][K] K-major in the canonical no-swizzle UMMA layout
//   float index = ((k / 4) * N + n) * 4 + (k % 4)            (8-row x 16-byte core matrices, LBO = 16 N bytes, SBO = 128)
// twice: the tf32 "hi" part (low 13 mantissa bits cleared) and the fp32 remainder "lo" (3xTF32 split).
//   B1A  N=64 K=40  rows 0-29 l1_t1_2, rows 32-61 l1_t2_2; k 0-29 tr0, k 30-33 mask, k 34 bias (operand column = 1)
//   B1B  N=32 K=32  l1_t1_2[:, 30:60]  (mean over station neighbours)
//   B1C  N=32 K=32  l1_t2_2[:, 30:60]  (mean over source neighbours)
//   B2   N=96 K=64  rows 0-29 l2_t1_1, 32-61 l2_t2_1, 64-78 l2_t1_2[:, tr|mask], 80-94 l2_t2_2[:, tr|mask];
//                   k 0-29 tr[0:30], k 30,31 mask0,1, k 32-61 tr[30:60], k 62,63 mask2,3
//   B3A  N=16 K=32  l2_t1_2[:, 60:90];  B3B  N=16 K=32  l2_t2_2[:, 60:90]
constexpr int TC_BASE = (GENERIC_END + 63) / 64 * 64;   // 256-byte aligned
constexpr int TC_B1A_HI = 0, TC_B1A_LO = TC_B1A_HI + 64 * 40;
constexpr int TC_B1B_HI = TC_B1A_LO + 64 * 40, TC_B1B_LO = TC_B1B_HI + 32 * 32;
constexpr int TC_B1C_HI = TC_B1B_LO + 32 * 32, TC_B1C_LO = TC_B1C_HI + 32 * 32;
constexpr int TC_B2_HI = TC_B1C_LO + 32 * 32, TC_B2_LO = TC_B2_HI + 96 * 64;
constexpr int TC_B3A_HI = TC_B2_LO + 96 * 64, TC_B3A_LO = TC_B3A_HI + 16 * 32;
constexpr int TC_B3B_HI = TC_B3A_LO + 16 * 32, TC_B3B_LO = TC_B3B_HI + 16 * 32;
constexpr int TC_BIAS2 = TC_B3B_LO + 16 * 32;           // [96] bias of the B2 columns
constexpr int TC_SCAL = TC_BIAS2 + 96;                  // [8], see enum below
constexpr int TC_FLOATS = TC_SCAL + 8;
// TC_OK = 1 when activate12's slope is > 1e-3: layer 0 then stores p = PReLU12(tr0) (sums over source neighbours need
// no per-edge activation; tr0 and PReLU11(tr0) are recovered from p exactly up to rounding), else the generic kernels run.
// TC_OK = 1 also requires activate11's slope > 1e-3: the station-pass kernel stages PReLU11(tr0) and recovers tr0 from it.
enum { TCS_OK = 0, TCS_A1 = 1, TCS_A21 = 2, TCS_A22 = 3, TCS_R11 = 4, TCS_INV12 = 5, TCS_A12 = 6, TCS_INV11 = 7 };

// ---- weight blob of the two-pipeline station-pass kernel (da_s1_kernel.cu) ---------------------------------------------
// Same canonical no-swizzle K-major UMMA layout and hi / lo split as above, with every A operand 32 columns wide:
//   S1A  N=64 K=32  rows 0-29 l1_t1_2, rows 32-61 l1_t2_2; k 0-29 tr0, k 30,31 mask0, mask1
//   S1B  N=32 K=32  l1_t1_2[:, 30:60 | mask2, mask3]      (mean over station neighbours in k 0-29, mask2,3 in k 30,31)
//   S1C  N=32 K=32  l1_t2_2[:, 30:60 | mask2, mask3]      (mean over source neighbours)
//   S2   N=96 K=64  as B2
//   S3A / S3B       as B3A / B3B
// Biases ride on the "lo" pass of the A operand (pass A_lo x B_hi): the lo parts of the mask columns are zero, so the
// kernel writes 1.0 into lo columns 30 and 31 and uses, for that pass only, a copy of the k 24-31 block of B_hi whose
// rows 30 / 31 hold the hi / lo parts of the bias:   S1A_BIAS [N=64][K=8],  S2_BIAS [N=96][K=8].
constexpr int T2_BASE = (TC_BASE + TC_FLOATS + 63) / 64 * 64;
constexpr int T2_S1A_HI = 0, T2_S1A_LO = T2_S1A_HI + 64 * 32, T2_S1A_BIAS = T2_S1A_LO + 64 * 32;
constexpr int T2_S1B_HI = T2_S1A_BIAS + 64 * 8, T2_S1B_LO = T2_S1B_HI + 32 * 32;
constexpr int T2_S1C_HI = T2_S1B_LO + 32 * 32, T2_S1C_LO = T2_S1C_HI + 32 * 32;
constexpr int T2_S2_HI = T2_S1C_LO + 32 * 32, T2_S2_LO = T2_S2_HI + 96 * 64, T2_S2_BIAS = T2_S2_LO + 96 * 64;
constexpr int T2_S3A_HI = T2_S2_BIAS + 96 * 8, T2_S3A_LO = T2_S3A_HI + 16 * 32;
constexpr int T2_S3B_HI = T2_S3A_LO + 16 * 32, T2_S3B_LO = T2_S3B_HI + 16 * 32;
constexpr int T2_SCAL = T2_S3B_LO + 16 * 32;            // [8] copy of the TCS_* scalars
constexpr int T2_FLOATS = T2_SCAL + 8;

constexpr int PACKED_FLOATS = T2_BASE + T2_FLOATS;

static_assert(DA_W11 % 4 == 0 && DA_END % 4 == 0 && RI_END % 4 == 0 && SA_SIZE % 4 == 0 && TC_FLOATS % 4 == 0 &&
                  T2_FLOATS % 4 == 0,
              "16-byte alignment");

// ---- read-out heads (heads_kernels.cu; module.py:251-331) ---------------------------------------------------------------
// Every matrix K-major [n_in][ld]; packed on the host (genie_b200/ops.py HeadsWeights) at the offsets genie_heads_layout reports.
constexpr int HD_SD_W = 0;                          // SpatialDirect.f_direct                  [30][32]
constexpr int HD_SD_B = HD_SD_W + 30 * 32;          //                                         [32]
constexpr int HD_SD_SL = HD_SD_B + 32;              // [4]: activate
constexpr int HD_TA_WC1 = HD_SD_SL + 4;             // TemporalAttention.f_context_1           [30][32]
constexpr int HD_TA_BC1 = HD_TA_WC1 + 30 * 32;
constexpr int HD_TA_WV1 = HD_TA_BC1 + 32;           // f_values_1                              [30][32]
constexpr int HD_TA_BV1 = HD_TA_WV1 + 30 * 32;
constexpr int HD_TA_WV2 = HD_TA_BV1 + 32;           // f_values_2                              [30][76]
constexpr int HD_TA_BV2 = HD_TA_WV2 + 30 * 76;      //                                         [76]
constexpr int HD_TA_WP1 = HD_TA_BV2 + 76;           // proj_1                                  [15][32]
constexpr int HD_TA_BP1 = HD_TA_WP1 + 15 * 32;
constexpr int HD_TA_WP2 = HD_TA_BP1 + 32;           // proj_2 (30 -> 1)                        [32]
constexpr int HD_TA_BP2 = HD_TA_WP2 + 32;           //                                         [4]
constexpr int HD_TA_SL = HD_TA_BP2 + 4;             // [4]: activate1, activate2, activate4, activate5
constexpr int HD_SA_WQ = HD_TA_SL + 4;              // SpatialAttention.f_queries              [3][76]
constexpr int HD_SA_BQ = HD_SA_WQ + 3 * 76;
constexpr int HD_SA_WC = HD_SA_BQ + 76;             // f_context  rows 0-29 x_j, 30-32 edge attr [33][76]
constexpr int HD_SA_BC = HD_SA_WC + 33 * 76;
constexpr int HD_SA_WV = HD_SA_BC + 76;             // f_values                                [33][76]
constexpr int HD_SA_BV = HD_SA_WV + 33 * 76;
constexpr int HD_SA_WP = HD_SA_BV + 76;             // proj                                    [15][32]
constexpr int HD_SA_BP = HD_SA_WP + 15 * 32;
constexpr int HD_SA_SL = HD_SA_BP + 32;             // [4]: activate1, activate2
constexpr int HD_FLOATS = HD_SA_SL + 4;
static_assert(HD_FLOATS % 4 == 0 && HD_TA_WV2 % 4 == 0 && HD_SA_WC % 4 == 0 && HD_SA_WP % 4 == 0, "16-byte alignment");

// ---- node-feature row strides in the workspace (floats) -------------------------------------------------------------
constexpr int LD_TR0 = 32;   // tr0 rows padded to 128 B: one L2 line per gathered neighbour row
constexpr int LD_ZC = 32;    // [ca(15) 0 | cb(15) 0]
constexpr int LD_V = 16;     // va / vb rows (64 B = two sectors)

}  // namespace gl
