// Leaf-kernel interface between the engine (pb_engine.cpp, device-agnostic host C++) and the
// kernel library.  NOT part of the reference-facing ABI (that is pullback_b200.h): these symbols are exported from
// libpullback_b200.so so that every kernel can be unit-tested against torch through ctypes (tests/test_kernels_gpu.py) and
// timed alone (scripts/bench_gemm.py, scripts/bench_gn.py); tests/test_abi_symbols.py checks that each one is exported.  The product library implements every pbk_* as hand-written sm_100a CUDA
// (pb_kernels.cu, pb_gemm_sm100.cu, pb_ortho.cu).  tests/hostsim/ provides a plain-C++ double of
// the same symbols so the engine's sequencing logic can be unit-tested without a GPU; that double
// is test infrastructure only and is never linked into, or reachable from, the product library.
//
// Conventions: fp32, activations are [batch][pixels-or-tokens][channels] (NHWC), `st` is a
// cudaStream_t, every call is stream-ordered and returns nullptr or a static error string.
// `round_tf32` of a producer: 0 store fp32 as is; 1 RNA-round to TF32 (the consumer is a TF32 GEMM); 2 store fp16 --
// the output buffer then holds halves at the same element offsets and its only consumer is an fp16-operand GEMM
// (pbk_gn_lin, pbk_ln_lin, pbk_geglu_jvp/vjp, pbk_im2col_s2, pbk_upsample2x, pbk_transpose; not together with accumulation).
// io flags (the `round_tf32` / `io` argument of the tangent-path kernels): the low two bits are the OUTPUT mode above; bit 2 says the
// tangent / cotangent INPUT holds halves.  The all-fp16 tangent plan of the engine passes PB_IN_F16 | PB_OUT_F16 everywhere
// except at the ends (x_t-shaped V / W and h-shaped U cross the ABI as fp32).
#define PB_RND_MASK 3
#define PB_OUT_F16 2
#define PB_IN_F16 4
// "xp" arguments are PRIMAL tensors cached once per (x_t, t, prompt); "t"/"g" arguments carry the
// nb tangent (JVP) or cotangent (VJP) directions packed on the batch axis.
#pragma once
#include <cstddef>
#include <cstdint>

#include "pb_gemm.h"

typedef void* pb_stream;
#define PBK extern "C" __attribute__((visibility("default"))) const char*

// ---- memory ----
PBK pbk_memset0(void* p, size_t bytes, pb_stream st);
PBK pbk_copy(void* dst, const void* src, size_t bytes, pb_stream st);            // device -> device
PBK pbk_download(void* host_dst, const void* src, size_t bytes, pb_stream st);   // blocking
PBK pbk_upload(void* dst, const void* host_src, size_t bytes, pb_stream st);     // blocking
PBK pbk_sync(pb_stream st);
PBK pbk_backend_name();
// CUDA-graph capture of a launch sequence on `st` (must not be the legacy default stream)
PBK pbk_graph_begin(pb_stream st);
PBK pbk_graph_end(pb_stream st, void** graph_exec, long* kernel_nodes);
PBK pbk_graph_launch(void* graph_exec, pb_stream st);
PBK pbk_graph_destroy(void* graph_exec);

// ---- timing probes (pb_profile_*: per-kernel roofline of bench.py) ----
PBK pbk_event_record(void** ev, pb_stream st);                                    // creates *ev on first use
extern "C" __attribute__((visibility("default"))) float pbk_event_elapsed_ms(void* e0, void* e1);   // waits for e1
PBK pbk_event_destroy(void* ev);

// ---- contraction ----
PBK pbk_gemm(const PbGemm* g, pb_stream st);
// direct 3x3/s1/p1 conv for tiny channel counts (conv_in and its transpose); w is [Cout][9][Cin]
// io: PB_IN_F16 x holds halves, PB_OUT_F16 y holds halves (conv_in keeps its x_t-shaped side in fp32)
PBK pbk_conv3x3_direct(const float* x, int nb, int H, int W, int Cin, const float* w, const float* bias, int Cout,
                       float* y, float beta, int io, pb_stream st);
// stride-2 3x3 conv as im2col + GEMM; input row = 2*o + tap - pad_lo (pad_lo 1: SD, 0: DDPM (0,1,0,1) pad)
PBK pbk_im2col_s2(const float* x, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, float* col, int round_tf32,
                  pb_stream st);
PBK pbk_col2im_s2(const float* col, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, float* gx, float beta,
                  int round_tf32, pb_stream st);

// ---- data movement ----
// dst[r][c] = src[r][c] + beta * dst[r][c]   (channel-slice copies: concat / split / residual fan-in)
PBK pbk_copy2d(float* dst, long ldd, const float* src, long lds, long rows, int cols, float beta, int round_tf32,
               pb_stream st);
// per batch (b, h): src_bh = src + b*sbs + h*shs is [R][lds] (first C columns used) -> dst_bh = dst + b*sbd + h*shd
// is [C][ldd] (ldd >= R);  dst = src^T + beta * dst
PBK pbk_transpose(float* dst, long ldd, long sbd, long shd, const float* src, long lds, long sbs, long shs, int nb,
                  int nh, int R, int C, float beta, int round_tf32, pb_stream st);
PBK pbk_upsample2x(const float* x, int nb, int H, int W, int C, float* y, int round_tf32, pb_stream st);
PBK pbk_upsample2x_vjp(const float* gy, int nb, int H, int W, int C, float* gx, float beta, int round_tf32,
                       pb_stream st);
PBK pbk_round_tf32(float* dst, const float* src, size_t n, pb_stream st);
// fp16-operand GEMMs (kind::f16, same 10-bit mantissa as TF32, half the operand bytes): 1 if the backend has them
extern "C" __attribute__((visibility("default"))) int pbk_has_f16_operands();
PBK pbk_to_f16(void* dst, const float* src, size_t n, pb_stream st);               // n % 4 == 0
PBK pbk_to_f16_scaled(void* dst, const float* src, size_t n, float scale, pb_stream st);
PBK pbk_to_f32(float* dst, const void* src, size_t n, pb_stream st);              // n % 8 == 0

// ---- GroupNorm (+ optional SiLU) ----
// tmp: pbk_gn_tmp_floats(HW, C, G, nb) floats of scratch (per-chunk partial sums, combined in a fixed order)
extern "C" __attribute__((visibility("default"))) size_t pbk_gn_tmp_floats(int HW, int C, int G, int nb);
PBK pbk_gn_stats(const float* x, int nb, int HW, int C, int G, float eps, float* mean, float* rstd, float* tmp,
                 pb_stream st);
PBK pbk_gn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int nb,
                 int HW, int C, int G, int silu, int round_tf32, float* y, pb_stream st);
// Linearisation of y = act(gamma * xhat + beta) around the cached primal xp (one image, [HW][C]).
//   mode 0 (JVP): out = act'(.) * gamma * rstd * (t - mean_g(t) - xhat * mean_g(xhat * t))
//   mode 1 (VJP): g = t * act'(.) * gamma ; out = rstd * (g - mean_g(g) - xhat * mean_g(xhat * g))
// out = result + acc * out.  tmp: pbk_gn_tmp_floats(HW, C, G, nb) floats of scratch.
// Problem slots (k_slot, p_stride): image b of the tangent batch belongs to problem b / k_slot, whose primal tensors (xp,
// mean, rstd) start p_stride FLOATS after the previous problem's (one primal cache per problem, uniform stride).
// k_slot <= 0 or >= nb: one problem (p_stride ignored).
PBK pbk_gn_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, const float* beta, int HW,
               int C, int G, int silu, const float* t, int nb, int mode, float* out, float acc, int round_tf32,
               float* tmp, int k_slot, long p_stride, pb_stream st);

// ---- LayerNorm over the channel axis ----
PBK pbk_ln_fwd(const float* x, long rows, int C, const float* gamma, const float* beta, float eps, float* y,
               float* mean, float* rstd, int round_tf32, pb_stream st);
PBK pbk_ln_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, long rows_p, int C,
               const float* t, int nb, int mode, float* out, float acc, int round_tf32, int k_slot, long p_stride, pb_stream st);

// the GEMM's GEGLU tangent epilogue (PbGemm::gg) exists in this backend; the interleaved copy of the ff1 weight it takes:
// dst rows [64 j, 64 j + 32) = src rows [32 j, ...), dst rows [64 j + 32, 64 j + 64) = src rows [F + 32 j, ...); halves, F % 32 == 0
extern "C" __attribute__((visibility("default"))) int pbk_gemm_geglu_supported();
PBK pbk_interleave_rows16(void* dst, const void* src, int F, int cols, pb_stream st);
// ---- GEGLU: y = h[:, :F] * gelu_erf(h[:, F:]) ----
// prepare != 0: h = [a | g] is then overwritten IN PLACE with the linearisation factors [gelu(g) | a gelu'(g)], the form
// pbk_geglu_jvp / _vjp read (no erf / exp per element and iteration: the two linearisation kernels were ALU-bound on them)
PBK pbk_geglu_fwd(float* h, long rows, int F, float* y, int round_tf32, int prepare, pb_stream st);
// hp: the PREPARED cache [gelu(g) | a gelu'(g)] of pbk_geglu_fwd;  round_tf32 | PB_IN_F16: the tangent input dh / gy holds halves
PBK pbk_geglu_jvp(const float* hp, long rows_p, const float* dh, int nb, int F, float* dy, int round_tf32,
                  int k_slot, long p_stride, pb_stream st);
PBK pbk_geglu_vjp(const float* hp, long rows_p, const float* gy, int nb, int F, float* gh, int round_tf32,
                  int k_slot, long p_stride, pb_stream st);

// ---- softmax pieces (attention probabilities are materialised per (head, query) row) ----
PBK pbk_softmax_fwd(float* S, long rows, int cols, long ld, int round_tf32, pb_stream st);
// dS <- P * (dS - rowsum(P * dS)); P has rows_p rows and is broadcast over the nb tangents
// problem slots (k_slot, p_stride in FLOATS): tangent b reads the P of problem b / k_slot; k_slot <= 0: one problem
PBK pbk_softmax_lin(const float* P, long rows_p, float* dS, int nb, int cols, long ld, int round_tf32, int k_slot,
                    long p_stride, pb_stream st);
// delta[b][h][i] = sum_c go[b][i][h*d + c] * o[i][h*d + c]
// io & PB_IN_F16: go holds halves (o is primal fp32, delta fp32)
PBK pbk_attn_delta(const float* go, long ldg, const float* o, long ldo, int nb, int N, int H, int d, float* delta,
                   int io, int k_slot, long p_stride, pb_stream st);          // o of problem b / k_slot: o + (b / k_slot) * p_stride
// dP[b][h][r][c] <- scale * P[h][r][c] * (dP[b][h][r][c] - (col_mode ? delta[b][h][c] : delta[b][h][r]))
PBK pbk_attn_ds(const float* P, float* dP, const float* delta, float scale, int nb, int H, int rows, int cols, long ld,
                int col_mode, int round_tf32, int k_slot, long p_stride, pb_stream st);

// ---- fused attention linearisation (no N x N tangent in HBM); implemented in pb_attn_sm100.cu ----
// Per (b, h), rows r in [0, Mr), score columns c in [0, Nc), head dim d (head h occupies columns [h*d, (h+1)*d) of A/B/D/R/O):
//   S[r][c]  = alpha1 * sum_seg  A_seg[b][r][h*d + :] . B_seg[b][c][h*d + :]         (batch / head strides as in PbGemmSeg)
//   T[r][c]  = Pm[h][r][c] * (S[r][c] - delta),  delta = 0 | delta[b][h][r] (mode 1) | delta[b][h][c] (mode 2)
//   E[r][n]  = sum_c Pm[h][r][c] * C2[b][h][n][c]                                     (optional second product on the same P tile)
//   D[b][r][h*d + n] = alpha2 * (sum_c T[r][c] * C1[h][n][c]  +  (D2 ? 0 : E[r][n]))
//                      - (want_rsum ? rowsum_c(T[r][:]) * O[r][h*d + n] : 0)  + beta * R[b][r][h*d + n]
//   D2[b][r][h*d + n] = E[r][n]                                                       (when D2 is given)
// so that the probability matrix is streamed exactly once per (tangent, head) for everything that contracts with it.
struct PbAttnLin {
  int Mr, Nc, d, nb, nh, nseg;
  PbGemmSeg seg[2];                 // seg[i].K is ignored (K = d)
  float alpha1, alpha2, beta;
  const float* Pm; long ldp, sPh;   // [nh][Mr][ldp]
  const float* delta; int delta_mode;
  int want_rsum; const float* O; long ldo;
  const float* C1; long ldc, sCh;   // [nh][d][ldc]
  float* D; long ldd, sDb;
  const float* R; long ldr, sRb;
  int round_tf32;
  const float* C2; long ldc2, sC2h, sC2b;   // optional [nb][nh][d][ldc2]
  float* D2; long ldd2, sD2b;               // optional separate output of the C2 product
  // p16: Pm, C1 and C2 hold HALVES (leading dimensions / strides in elements, rows multiples of 8) and Pm is pre-scaled by
  // p_scale (= Nc keeps softmax probabilities in fp16's normal range); the kernel divides the products by p_scale again
  int p16; float p_scale;
  // s16 (needs p16): the S operands seg[].A / seg[].B hold halves as well (kind::f16 score products; head dim % 8 == 0) and D / D2
  // are written as halves (ldd, sDb, ldd2, sD2b in elements); O, delta stay fp32; no residual R
  int s16;
  // problem slots: tangent b belongs to problem b / k_slot; every PRIMAL operand (Pm, C1, O and the segment operands whose batch
  // stride is 0) of problem s starts p_stride BYTES after problem s - 1's.  k_slot <= 0: one problem.
  int k_slot; long p_stride;
};
PBK pbk_attn_lin_supported(int d, int Mr, int Nc);    // nullptr if pbk_attn_lin handles this geometry
PBK pbk_attn_lin(const PbAttnLin* a, pb_stream st);
// sizeof(PbGemm) / sizeof(PbAttnLin) as this library was compiled: a binding that mirrors the two descriptors (ctypes in
// diffusion_pullback_b200/_native.py) checks its layout against them at load time instead of passing a short struct
extern "C" __attribute__((visibility("default"))) void pbk_struct_sizes(int* gemm_bytes, int* attn_lin_bytes);
// debugging aid of the column-batched kernel (PB_ATTN_TRACE=1, scripts/trace_attn.py): per-substep event clocks of CTA (0, 0)
// of the last launch, [512][16] values; returns the number of values copied
extern "C" __attribute__((visibility("default"))) int pbk_attn16_trace_read(long long* host, int n);

// ---- time embedding (primal only) ----
PBK pbk_timestep_embedding(float t, int dim, int flip_sin_to_cos, float freq_shift, float* out, pb_stream st);
// y = act_out(W[N][K] . act_in(x) + bias); act flags: 1 = SiLU
PBK pbk_gemv(const float* Wm, const float* x, const float* bias, int N, int K, int silu_in, int silu_out, float* y,
             pb_stream st);

// ---- DDIM update (eta = 0): x_next = sqrt(a_next) * (x - sqrt(1 - a_t) * eps) / sqrt(a_t) + sqrt(1 - a_next) * eps ----
PBK pbk_ddim_step(const float* x, const float* eps, float a_t, float a_next, float* x_next, float* pred_x0, long n,
                  pb_stream st);

// out = a x + b y + c z (y, z optional)
PBK pbk_lincomb3(float* out, float a, const float* x, float b, const float* y, float c, const float* z, long n, pb_stream st);

// ---- weight packing (one-time) ----
// w [Co][Ci][3][3] (PyTorch) -> fwd [Co][9][Ci], bwd [Ci][9][Co] with taps flipped (transpose conv)
PBK pbk_pack_conv3x3(const float* w, int Co, int Ci, float* fwd, float* bwd, int round_tf32, pb_stream st);

// ---- subspace re-orthonormalisation (replaces torch.linalg.svd of the k x n_in matrix) ----
// G = W W^T, M = W Vprev^T  (fp64 accumulation), W/Vprev: [k][n]
PBK pbk_gram2(const float* Wm, const float* Vprev, int k, long n, double* G, double* M, pb_stream st);
// Jacobi eigen-decomposition of G (one warp, fp64): sv[i] = lambda_i^(1/4) descending (the reference returns
// sqrt of svdvals(W) = sqrt(sqrt(eig))), Rm[i][j] = sign_i * X[j][i] / sqrt(lambda_i), sign from M.
PBK pbk_jacobi(const double* G, const double* M, int k, float* Rm, float* sv, pb_stream st);
// V = Rm W ; metrics[0] += sum (V - Vprev)^2 ; metrics[1] += #{|V - Vprev| > atol + rtol*|V|}
PBK pbk_rotate(const float* Wm, const float* Rm, const float* Vprev, int k, long n, float atol, float rtol, float* V,
               float* metrics, pb_stream st);
