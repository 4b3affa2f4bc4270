// Segmented (implicit-)GEMM descriptor shared by the engine and the backends.
//
//   D[b,h][m,n] = alpha * sum_seg sum_k A_seg[b,h][m,k] * B_seg[b,h][n,k]  + bias[n] + beta * R[b,h][m,n]
//
// All matrices are K-major (row-major [rows][K]), fp32 or fp16; the contraction runs on the tcgen05 tensor
// cores as TF32 or F16 with fp32 accumulation in TMEM.
//
//  * plain mode: A is [M][K] with leading dimension lda and two batch strides (b = tangent index,
//    h = attention head); a stride of 0 broadcasts the operand over that batch dimension.
//  * conv mode (3x3, stride 1, pad 1, NHWC): A is the activation tensor [nb][H][W][C] (pixel stride lda);
//    M = nb*H*W rows; K runs over 9 taps x C channels and B is the packed filter [N][9*C]
//    (tap-major, channel fastest).  Padding is produced by TMA out-of-bounds zero fill.
#pragma once
#include <cstdint>

enum { PB_GEMM_F32 = 0, PB_GEMM_F16 = 1 };

// leading dimensions and batch strides are in ELEMENTS of the operand type
struct PbGemmSeg {
  const void* A; long lda, sAb, sAh;
  const void* B; long ldb, sBb, sBh;
  int K;
};

struct PbGemm {
  int M, N;
  int nseg;
  PbGemmSeg seg[2];
  void* D; long ldd, sDb, sDh;
  const void* R; long ldr, sRb, sRh;    // optional residual of D's type (may alias D element-for-element)
  const float* bias;                     // optional [N]
  float alpha, beta;
  int nb, nh;                            // batch extents (plain mode); conv mode: nb images, nh = 1
  int conv;                              // 0 plain, 1 = 3x3/s1/p1 NHWC implicit GEMM
  int H, W;                              // conv mode spatial extent
  int round_tf32;                        // round the stored result to TF32 (RNA) for a GEMM-only consumer
  int precise;                           // 0: one TF32 pass; 1: error-compensated 3xTF32 (A and B split hi/lo);
                                         // 2: B (weights) split hi/lo, A as given
  // optional split-K scratch owned by the caller (nullptr: never split): room for the partial tiles
  float* ws; long ws_floats;
  // element types: operands A/B fp32 (TF32 tensor cores) or fp16 (kind::f16); D/R fp32 or fp16, independently of the
  // operand type; bias, alpha, beta and the accumulation are always fp32
  int ab_dtype, d_dtype;
  // problem slots (plain mode): batch index b belongs to problem b / k_slot; an operand whose batch stride (sAb / sBb) is 0 is
  // PRIMAL and the copy of problem s starts p_stride BYTES after problem s - 1's.  k_slot <= 0 or >= nb: one problem.
  int k_slot; long p_stride;
  // GEGLU tangent epilogue (JVP of ff1 fused with the GEGLU linearisation; nullptr: off).  B holds the 2 F weight rows INTERLEAVED in
  // blocks of 64 ([32 rows of the a half | the 32 matching rows of the gate half], pbk_interleave_rows16), N = 2 F, and
  //   D[m][f] = acc_a[m][f] * G1[m'][f] + acc_g[m][f] * G2[m'][f],   f < F,  D is [M][ldd >= F] halves,
  // with [G1 | G2] = gg[m'] the prepared factor cache [gelu(g) | a gelu'(g)] of pbk_geglu_fwd ([gg_rows_p][2 F] floats per
  // problem; m' = m % gg_rows_p, problem = (m / gg_rows_p) / gg_k_slot at gg_p_stride floats).  fp16 operands and output, plain
  // mode, nb = nh = 1, no residual / bias, alpha = 1, F % 128 == 0; pbk_gemm_geglu_supported() tells whether the backend has it.
  const float* gg; int gg_F; long gg_rows_p; int gg_k_slot; long gg_p_stride;
};

static inline PbGemm pb_gemm_init() {
  PbGemm g{};
  g.nseg = 1; g.alpha = 1.f; g.beta = 0.f; g.nb = 1; g.nh = 1;
  return g;
}

// tuning hooks of the sm_100a GEMM (scripts/bench_gemm.py): force a tile width (0 = heuristic), the split-K policy
// (0 off, 1 heuristic, > 1 forced split count), allow BN = 160; minimum k-blocks for a split
#ifdef __cplusplus
extern "C" {
#endif
void pb_gemm_tune(int force_bn, int split, int use160);
void pb_gemm_tune_split_min_kb(int kb);
void pb_gemm_tune_pair(int on);          // CTA-pair (cta_group::2) kernels on / off (default on; env PB_GEMM_PAIR)
#ifdef __cplusplus
}
#endif
