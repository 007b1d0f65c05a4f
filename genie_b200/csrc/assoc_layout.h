// Layout of the packed association weights (assoc_kernels.cu, reported to the host by genie_assoc_layout and packed in
// genie_b200/ops.py AssocWeights), shared with the packer of the tensor-core blob of the tiled layer-1 pass (pack.cu).
#pragma once

namespace as {
// every matrix K-major [n_in][ld]
constexpr int SD_W = 0;                        // SpatialDirect.f_direct                         [30][32]
constexpr int SD_B = SD_W + 30 * 32;           //                                                [32]
constexpr int RO_WY = SD_B + 32;               // read-out fc1[:, 0:30]  (y_latent)              [30][32]
constexpr int RO_B1 = RO_WY + 30 * 32;         //                                                [32]
constexpr int RO_WA = RO_B1 + 32;              // read-out fc1[:, 30:33] (edge attr)             [4][32] (3 used)
constexpr int RO_W2 = RO_WA + 4 * 32;          // read-out fc2                                   [30][16]
constexpr int RO_B2 = RO_W2 + 30 * 16;         //                                                [16]
constexpr int AI_W = RO_B2 + 16;               // init_trns: rows 0-14 s0, 15-44 x_latent, 45 mask_out, 46-49 Mask   [52][32] (50 used)
constexpr int AI_B = AI_W + 52 * 32;
constexpr int M11_W = AI_B + 32;               // l1_t1_1                                        [30][32]
constexpr int M11_B = M11_W + 30 * 32;
constexpr int M12_W = M11_B + 32;              // l1_t2_1                                        [30][32]
constexpr int M12_B = M12_W + 30 * 32;
constexpr int INIT_END = M12_B + 32;
// ---- layer-1 block (copied to shared memory as one piece) ----
constexpr int W11 = INIT_END;                  // l1_t1_2: rows 0-29 tr, 30-59 mean_sta, 60-64 mask5   [68][32] (65 used)
constexpr int W12 = W11 + 68 * 32;             // l1_t2_2
constexpr int B11 = W12 + 68 * 32;
constexpr int B12 = B11 + 32;
constexpr int W21A = B12 + 32;                 // l2_t1_1                                        [60][32]
constexpr int W22A = W21A + 60 * 32;           // l2_t2_1
constexpr int B21A = W22A + 60 * 32;
constexpr int B22A = B21A + 32;
constexpr int WVA = B22A + 32;                 // l2_t1_2[:, 60:90]                              [30][16]
constexpr int WVB = WVA + 30 * 16;             // l2_t2_2[:, 60:90]
constexpr int WCA = WVB + 30 * 16;             // l2_t1_2[:, 0:60 | 90:95]: rows 0-59 tr, 60-64 mask5   [68][16] (65 used)
constexpr int WCB = WCA + 68 * 16;
constexpr int BCA = WCB + 68 * 16;
constexpr int BCB = BCA + 16;
constexpr int L1_END = BCB + 16;
// ---- LocalSliceLgCollapse P / S ----
constexpr int CP_W1 = L1_END;                  // fc1: rows 0-29 s_j, 30 (t_pick - t_j)/eps, 31 phase  [32][32]
constexpr int CP_B1 = CP_W1 + 32 * 32;
constexpr int CP_W2 = CP_B1 + 32;              // fc2                                            [30][16]
constexpr int CP_B2 = CP_W2 + 30 * 16;
constexpr int CS_W1 = CP_B2 + 16;
constexpr int CS_B1 = CS_W1 + 32 * 32;
constexpr int CS_W2 = CS_B1 + 32;
constexpr int CS_B2 = CS_W2 + 30 * 16;
constexpr int C_SIZE = CS_W1 - CP_W1;
constexpr int SL = CS_B2 + 16;                 // slopes [16], see enum
constexpr int FLOATS = SL + 16;
enum { SL_SD = 0, SL_RO1, SL_RO2, SL_A, SL_A11, SL_A12, SL_A1, SL_A21, SL_A22, SL_A2, SL_CP1, SL_CP2, SL_CS1, SL_CS2 };
static_assert(W11 % 4 == 0 && WVA % 4 == 0 && WCA % 4 == 0 && CP_W1 % 4 == 0 && SL % 4 == 0 && RO_W2 % 4 == 0, "alignment");

constexpr int LD_S = 32;                       // s rows: [o1(15) 0 | o2(15) 0]
}  // namespace as
