// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or reachable from the product library.
//
// A plain-C++ double of the leaf-kernel interface (include/pb_kernels.h) so that the HOST logic of the
// engine (csrc/pb_engine.cpp: planning, weight packing, op sequencing, cotangent accumulation order,
// workspace offsets, the iteration loop) can be unit-tested in the CPU-only authoring container
// (`pytest -m "not gpu"`), where no CUDA kernel can run.  tests/hostsim/build.py links it with the
// unmodified pb_engine.cpp into tests/hostsim/libpb_hostsim.so; pb_backend() of that library reports
// "hostsim" and diffusion_pullback_b200._native refuses anything but "cuda-sm100a".
// Semantics mirror the CUDA kernels, including TF32 operand truncation in the GEMM and RNA rounding.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pb_kernels.h"

#include <cstdlib>
namespace {
// PB_HOSTSIM_EXACT=1: keep fp32 everywhere (separates engine-logic errors from TF32 effects)
const bool kExact = std::getenv("PB_HOSTSIM_EXACT") != nullptr;
inline float rna(float x) {
  if (kExact) return x;
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u = (u + 0x1000u) & ~0x1fffu;
  memcpy(&x, &u, 4); return x;
}
inline float trunc_tf32(float x) { if (kExact) return x; uint32_t u; memcpy(&u, &x, 4); u &= ~0x1fffu; memcpy(&x, &u, 4); return x; }
inline float mr(float x, int r) { return r ? rna(x) : x; }
inline float opA(float x, int precise) { return precise == 1 ? x : trunc_tf32(x); }
inline float opB(float x, int precise) { return precise >= 1 ? x : trunc_tf32(x); }
// element accessors for buffers that are fp32 (dt = 0) or fp16 (dt = 1; round-to-nearest-even like cvt.rn.f16.f32)
using h16 = _Float16;
inline float ldx(const void* p, int dt, long i) { return dt ? (float)static_cast<const h16*>(p)[i] : static_cast<const float*>(p)[i]; }
inline void stx(void* p, int dt, long i, float v) { if (dt) static_cast<h16*>(p)[i] = (h16)v; else static_cast<float*>(p)[i] = v; }
inline int in16(int io) { return (io & PB_IN_F16) ? 1 : 0; }
inline int out16(int io) { return (io & PB_RND_MASK) == PB_OUT_F16 ? 1 : 0; }
// store: fp16 when the io flags say so, else fp32 (RNA-rounded when the low bit asks for it)
inline void sto(void* p, int io, long i, float v) { if (out16(io)) static_cast<h16*>(p)[i] = (h16)v; else static_cast<float*>(p)[i] = mr(v, io & 1); }
inline float sigm(float x) { return 1.f / (1.f + std::exp(-x)); }
inline float silu_f(float x) { return x * sigm(x); }
inline float silu_d(float x) { float s = sigm(x); return s * (1.f + x * (1.f - s)); }
inline float gelu_f(float g) { return 0.5f * g * (1.f + std::erf(g * 0.70710678118654752f)); }
inline float gelu_d(float g) { return 0.5f * (1.f + std::erf(g * 0.70710678118654752f)) + g * 0.3989422804014327f * std::exp(-0.5f * g * g); }
}  // namespace

PBK pbk_backend_name() { return "hostsim"; }
PBK pbk_memset0(void* p, size_t bytes, pb_stream) { memset(p, 0, bytes); return nullptr; }
PBK pbk_copy(void* dst, const void* src, size_t bytes, pb_stream) { memmove(dst, src, bytes); return nullptr; }
PBK pbk_download(void* dst, const void* src, size_t bytes, pb_stream) { memmove(dst, src, bytes); return nullptr; }
PBK pbk_upload(void* dst, const void* src, size_t bytes, pb_stream) { memmove(dst, src, bytes); return nullptr; }
PBK pbk_sync(pb_stream) { return nullptr; }
PBK pbk_graph_begin(pb_stream) { return "hostsim: no graphs"; }
PBK pbk_graph_end(pb_stream, void**, long*) { return "hostsim: no graphs"; }
PBK pbk_graph_launch(void*, pb_stream) { return "hostsim: no graphs"; }
PBK pbk_graph_destroy(void*) { return nullptr; }
// PB_HOSTSIM_F16=1: the double models the all-fp16 tangent plan too (halves by _Float16, round to nearest even), so that the
// engine's fp16 sequencing / offsets are testable without a GPU; default: the fp32 / TF32 policy only
extern "C" __attribute__((visibility("default"))) int pbk_has_f16_operands() { return std::getenv("PB_HOSTSIM_F16") != nullptr; }
PBK pbk_to_f16_scaled(void* dst, const float* src, size_t n, float scale, pb_stream) {
  for (size_t i = 0; i < n; ++i) static_cast<h16*>(dst)[i] = (h16)(src[i] * scale);
  return nullptr;
}
PBK pbk_to_f16(void* dst, const float* src, size_t n, pb_stream st) { return pbk_to_f16_scaled(dst, src, n, 1.f, st); }
PBK pbk_to_f32(float* dst, const void* src, size_t n, pb_stream) {
  for (size_t i = 0; i < n; ++i) dst[i] = (float)static_cast<const h16*>(src)[i];
  return nullptr;
}
// timing probes: the double has no clock; a non-null token keeps the engine's bookkeeping exercised
PBK pbk_event_record(void** ev, pb_stream) { static int token; *ev = &token; return nullptr; }
extern "C" __attribute__((visibility("default"))) float pbk_event_elapsed_ms(void*, void*) { return 0.f; }
PBK pbk_event_destroy(void*) { return nullptr; }

PBK pbk_gemm(const PbGemm* gp, pb_stream) {
  const PbGemm& g = *gp;
  if (g.M <= 0 || g.N <= 0) return "gemm: empty problem";
  const int ab = g.ab_dtype, dd = g.d_dtype;
  if (g.conv) {
    const PbGemmSeg& s = g.seg[0];
    const int C = s.K, H = g.H, W = g.W;
    if (C % 8) return "gemm: conv channels must be a multiple of 8";
#pragma omp parallel for collapse(2) schedule(static)
    for (long pix = 0; pix < (long)g.nb * H * W; ++pix)
      for (int n = 0; n < g.N; ++n) {
        const int x = pix % W, y = (pix / W) % H; const long b = pix / ((long)W * H);
        double acc = 0.0;
        for (int tap = 0; tap < 9; ++tap) {
          const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
          if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
          const long a0 = ((b * H + iy) * W + ix) * s.lda;
          const long w0 = (long)n * s.ldb + (long)tap * C;
          float part = 0.f;
          for (int c = 0; c < C; ++c)
            part += ab ? ldx(s.A, 1, a0 + c) * ldx(s.B, 1, w0 + c)
                       : opA(ldx(s.A, 0, a0 + c), g.precise) * opB(ldx(s.B, 0, w0 + c), g.precise);
          acc += part;
        }
        float v = g.alpha * (float)acc;
        if (g.bias) v += g.bias[n];
        if (g.R) v += g.beta * ldx(g.R, dd, pix * g.ldr + n);
        stx(g.D, dd, pix * g.ldd + n, dd ? v : mr(v, g.round_tf32));
      }
    return nullptr;
  }
  const long kslot = (g.k_slot > 0 && g.k_slot < g.nb) ? g.k_slot : g.nb;
  const long pstride = kslot < g.nb ? g.p_stride : 0;
  for (int b = 0; b < g.nb; ++b)
    for (int h = 0; h < g.nh; ++h) {
#pragma omp parallel for collapse(2) schedule(static)
      for (int m = 0; m < g.M; ++m)
        for (int n = 0; n < g.N; ++n) {
          double acc = 0.0;
          for (int si = 0; si < g.nseg; ++si) {
            const PbGemmSeg& s = g.seg[si];
            // problem slots: a primal (batch-broadcast) operand of problem b / k_slot starts p_stride bytes further on
            const long es = ab ? 2 : 4;
            const long a0 = (s.sAb == 0 ? (b / kslot) * (pstride / es) : b * s.sAb) + h * s.sAh + (long)m * s.lda;
            const long w0 = (s.sBb == 0 ? (b / kslot) * (pstride / es) : b * s.sBb) + h * s.sBh + (long)n * s.ldb;
            float part = 0.f;
            for (int k = 0; k < s.K; ++k)
              part += ab ? ldx(s.A, 1, a0 + k) * ldx(s.B, 1, w0 + k)
                         : opA(ldx(s.A, 0, a0 + k), g.precise) * opB(ldx(s.B, 0, w0 + k), g.precise);
            acc += part;
          }
          float v = g.alpha * (float)acc;
          if (g.bias) v += g.bias[n];
          if (g.R) v += g.beta * ldx(g.R, dd, b * g.sRb + h * g.sRh + (long)m * g.ldr + n);
          stx(g.D, dd, b * g.sDb + h * g.sDh + (long)m * g.ldd + n, dd ? v : mr(v, g.round_tf32));
        }
    }
  return nullptr;
}

PBK pbk_conv3x3_direct(const float* x, int nb, int H, int W, int Cin, const float* w, const float* bias, int Cout, float* y,
                       float beta, int io, pb_stream) {
  const int xi = in16(io), yo = out16(io);
#pragma omp parallel for schedule(static)
  for (long pix = 0; pix < (long)nb * H * W; ++pix) {
    const int px = pix % W, py = (pix / W) % H; const long b = pix / ((long)W * H);
    for (int co = 0; co < Cout; ++co) {
      float acc = bias ? bias[co] : 0.f;
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = py + tap / 3 - 1, ix = px + tap % 3 - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const long x0 = ((b * H + iy) * W + ix) * Cin;
        const float* wp = w + ((long)co * 9 + tap) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc += ldx(x, xi, x0 + ci) * wp[ci];
      }
      const long o = pix * Cout + co;
      stx(y, yo, o, beta != 0.f ? acc + beta * ldx(y, yo, o) : acc);
    }
  }
  return nullptr;
}
PBK pbk_im2col_s2(const float* x, int nb, int H, int W, int C, int pad, int Ho, int Wo, float* col, int rnd, pb_stream) {
  for (long b = 0; b < nb; ++b)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox)
        for (int tap = 0; tap < 9; ++tap) {
          const int iy = 2 * oy + tap / 3 - pad, ix = 2 * ox + tap % 3 - pad;
          const long d0 = ((((b * Ho + oy) * Wo + ox) * 9) + tap) * C;
          for (int c = 0; c < C; ++c)
            sto(col, rnd, d0 + c, (iy >= 0 && iy < H && ix >= 0 && ix < W) ? ldx(x, in16(rnd), ((b * H + iy) * W + ix) * C + c) : 0.f);
        }
  return nullptr;
}
PBK pbk_col2im_s2(const float* col, int nb, int H, int W, int C, int pad, int Ho, int Wo, float* gx, float beta, int rnd,
                  pb_stream) {
  std::vector<float> acc((size_t)nb * H * W * C, 0.f);
  for (long b = 0; b < nb; ++b)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox)
        for (int tap = 0; tap < 9; ++tap) {
          const int iy = 2 * oy + tap / 3 - pad, ix = 2 * ox + tap % 3 - pad;
          if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
          const long s0 = ((((b * Ho + oy) * Wo + ox) * 9) + tap) * C;
          float* d = acc.data() + ((b * H + iy) * W + ix) * C;
          for (int c = 0; c < C; ++c) d[c] += ldx(col, in16(rnd), s0 + c);
        }
  for (size_t i = 0; i < acc.size(); ++i) sto(gx, rnd, (long)i, beta != 0.f ? acc[i] + beta * ldx(gx, out16(rnd), (long)i) : acc[i]);
  return nullptr;
}
PBK pbk_copy2d(float* dst, long ldd, const float* src, long lds, long rows, int cols, float beta, int rnd, pb_stream) {
  for (long r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float v = ldx(src, in16(rnd), r * lds + c);
      sto(dst, rnd, r * ldd + c, beta != 0.f ? v + beta * ldx(dst, out16(rnd), r * ldd + c) : v);
    }
  return nullptr;
}
PBK pbk_transpose(float* dst, long ldd, long sbd, long shd, const float* src, long lds, long sbs, long shs, int nb, int nh,
                  int R, int C, float beta, int rnd, pb_stream) {
  for (long b = 0; b < nb; ++b)
    for (long h = 0; h < nh; ++h) {
      const long s0 = b * sbs + h * shs, d0 = b * sbd + h * shd;
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
          float v = ldx(src, in16(rnd), s0 + (long)r * lds + c);
          const long o = d0 + (long)c * ldd + r;
          if (beta != 0.f) v += beta * ldx(dst, out16(rnd), o);
          sto(dst, rnd, o, v);
        }
    }
  return nullptr;
}
PBK pbk_upsample2x(const float* x, int nb, int H, int W, int C, float* y, int rnd, pb_stream) {
  for (long b = 0; b < nb; ++b)
    for (int oy = 0; oy < 2 * H; ++oy)
      for (int ox = 0; ox < 2 * W; ++ox)
        for (int c = 0; c < C; ++c)
          sto(y, rnd, ((b * 2 * H + oy) * 2 * W + ox) * C + c, ldx(x, in16(rnd), ((b * H + oy / 2) * W + ox / 2) * C + c));
  return nullptr;
}
PBK pbk_upsample2x_vjp(const float* gy, int nb, int H, int W, int C, float* gx, float beta, int rnd, pb_stream) {
  for (long b = 0; b < nb; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < C; ++c) {
          float a = 0.f;
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) a += ldx(gy, in16(rnd), ((b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C + c);
          const long o = ((b * H + y) * W + x) * C + c;
          sto(gx, rnd, o, beta != 0.f ? a + beta * ldx(gx, out16(rnd), o) : a);
        }
  return nullptr;
}
PBK pbk_round_tf32(float* dst, const float* src, size_t n, pb_stream) {
  for (size_t i = 0; i < n; ++i) dst[i] = rna(src[i]);
  return nullptr;
}

extern "C" __attribute__((visibility("default"))) size_t pbk_gn_tmp_floats(int, int, int, int) { return 16; }
PBK pbk_gn_stats(const float* x, int nb, int HW, int C, int G, float eps, float* mean, float* rstd, float*, pb_stream) {
  const int cpg = C / G;
  for (int b = 0; b < nb; ++b)
    for (int g = 0; g < G; ++g) {
      double s = 0, q = 0;
      for (int p = 0; p < HW; ++p)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { const double v = x[((long)b * HW + p) * C + c]; s += v; q += v * v; }
      const double n = (double)HW * cpg, m = s / n, var = q / n - m * m;
      mean[b * G + g] = (float)m; rstd[b * G + g] = (float)(1.0 / std::sqrt(var + eps));
    }
  return nullptr;
}
PBK pbk_gn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int nb, int HW,
                 int C, int G, int silu, int rnd, float* y, pb_stream) {
  const int cpg = C / G;
  for (long i = 0; i < (long)nb * HW * C; ++i) {
    const int c = i % C; const long b = i / ((long)HW * C); const int g = c / cpg;
    float v = gamma[c] * ((x[i] - mean[b * G + g]) * rstd[b * G + g]) + beta[c];
    if (silu) v = silu_f(v);
    y[i] = mr(v, rnd);
  }
  return nullptr;
}
PBK pbk_gn_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, const float* beta, int HW, int C, int G,
               int silu, const float* t, int nb, int mode, float* out, float acc, int rnd, float*, int k_slot, long p_stride,
               pb_stream) {
  const int cpg = C / G;
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  const float *xp0 = xp, *mean0 = mean, *rstd0 = rstd;
  for (long b = 0; b < nb; ++b)
    for (int g = 0; g < G; ++g) {
      xp = xp0 + (b / k_slot) * p_stride; mean = mean0 + (b / k_slot) * p_stride; rstd = rstd0 + (b / k_slot) * p_stride;
      double s1 = 0, s2 = 0;
      for (int p = 0; p < HW; ++p)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          const float xh = (xp[(long)p * C + c] - mean[g]) * rstd[g];
          float u = ldx(t, in16(rnd), (b * HW + p) * C + c);
          if (mode == 1) u *= silu ? gamma[c] * silu_d(gamma[c] * xh + beta[c]) : gamma[c];
          s1 += u; s2 += (double)xh * u;
        }
      const float m1 = (float)(s1 / ((double)HW * cpg)), m2 = (float)(s2 / ((double)HW * cpg));
      for (int p = 0; p < HW; ++p)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          const float xh = (xp[(long)p * C + c] - mean[g]) * rstd[g];
          const float f = silu ? gamma[c] * silu_d(gamma[c] * xh + beta[c]) : gamma[c];
          const float tv = ldx(t, in16(rnd), (b * HW + p) * C + c);
          float v = mode == 0 ? f * rstd[g] * (tv - m1 - xh * m2) : rstd[g] * (tv * f - m1 - xh * m2);
          const long o = (b * HW + p) * C + c;
          if (acc != 0.f) v += acc * ldx(out, out16(rnd), o);
          sto(out, rnd, o, v);
        }
    }
  return nullptr;
}
PBK pbk_ln_fwd(const float* x, long rows, int C, const float* gamma, const float* beta, float eps, float* y, float* mean,
               float* rstd, int rnd, pb_stream) {
  for (long r = 0; r < rows; ++r) {
    double s = 0, q = 0;
    for (int c = 0; c < C; ++c) s += x[r * C + c];
    const double m = s / C;
    for (int c = 0; c < C; ++c) q += (x[r * C + c] - m) * (x[r * C + c] - m);
    const float rs = (float)(1.0 / std::sqrt(q / C + eps));
    mean[r] = (float)m; rstd[r] = rs;
    for (int c = 0; c < C; ++c) y[r * C + c] = mr(gamma[c] * ((x[r * C + c] - (float)m) * rs) + beta[c], rnd);
  }
  return nullptr;
}
PBK pbk_ln_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, long rows_p, int C, const float* t,
               int nb, int mode, float* out, float acc, int rnd, int k_slot, long p_stride, pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  const float *xp0 = xp, *mean0 = mean, *rstd0 = rstd;
  for (long r = 0; r < rows_p * nb; ++r) {
    const long rp = r % rows_p, ps = ((r / rows_p) / k_slot) * p_stride;
    xp = xp0 + ps; mean = mean0 + ps; rstd = rstd0 + ps;
    double s1 = 0, s2 = 0;
    for (int c = 0; c < C; ++c) {
      const float xh = (xp[rp * C + c] - mean[rp]) * rstd[rp];
      const float tv = ldx(t, in16(rnd), r * C + c);
      const float u = mode == 1 ? tv * gamma[c] : tv;
      s1 += u; s2 += (double)xh * u;
    }
    const float m1 = (float)(s1 / C), m2 = (float)(s2 / C);
    for (int c = 0; c < C; ++c) {
      const float xh = (xp[rp * C + c] - mean[rp]) * rstd[rp];
      const float tv = ldx(t, in16(rnd), r * C + c);
      float v = mode == 0 ? gamma[c] * rstd[rp] * (tv - m1 - xh * m2) : rstd[rp] * (tv * gamma[c] - m1 - xh * m2);
      const long o = r * C + c;
      if (acc != 0.f) v += acc * ldx(out, out16(rnd), o);
      sto(out, rnd, o, v);
    }
  }
  return nullptr;
}
// the GEGLU tangent epilogue of the GEMM is a device-backend fusion: the engine keeps the two separate ops on this double
extern "C" __attribute__((visibility("default"))) int pbk_gemm_geglu_supported() { return 0; }
extern "C" __attribute__((visibility("default"))) void pbk_struct_sizes(int* gemm_bytes, int* attn_lin_bytes) {
  if (gemm_bytes) *gemm_bytes = (int)sizeof(PbGemm);
  if (attn_lin_bytes) *attn_lin_bytes = (int)sizeof(PbAttnLin);
}
PBK pbk_interleave_rows16(void*, const void*, int, int, pb_stream) { return "interleave_rows16: not in the host double"; }
PBK pbk_geglu_fwd(float* h, long rows, int F, float* y, int rnd, int prepare, pb_stream) {
  for (long r = 0; r < rows; ++r)
    for (int c = 0; c < F; ++c) {
      const float a = h[r * 2 * F + c], g = h[r * 2 * F + F + c];
      y[r * F + c] = mr(a * gelu_f(g), rnd);
      if (prepare) { h[r * 2 * F + c] = gelu_f(g); h[r * 2 * F + F + c] = a * gelu_d(g); }
    }
  return nullptr;
}
PBK pbk_geglu_jvp(const float* hp0, long rows_p, const float* dh, int nb, int F, float* dy, int rnd, int k_slot, long p_stride,
                  pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  for (long r = 0; r < rows_p * nb; ++r) {
    const long rp = r % rows_p;
    const float* hp = hp0 + ((r / rows_p) / k_slot) * p_stride;
    for (int c = 0; c < F; ++c) {
      const float g1 = hp[rp * 2 * F + c], g2 = hp[rp * 2 * F + F + c];      // prepared cache: gelu(g), a gelu'(g)
      sto(dy, rnd, r * F + c, ldx(dh, in16(rnd), r * 2 * F + c) * g1 + g2 * ldx(dh, in16(rnd), r * 2 * F + F + c));
    }
  }
  return nullptr;
}
PBK pbk_geglu_vjp(const float* hp0, long rows_p, const float* gy, int nb, int F, float* gh, int rnd, int k_slot, long p_stride,
                  pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  for (long r = 0; r < rows_p * nb; ++r) {
    const long rp = r % rows_p;
    const float* hp = hp0 + ((r / rows_p) / k_slot) * p_stride;
    for (int c = 0; c < F; ++c) {
      const float g1 = hp[rp * 2 * F + c], g2 = hp[rp * 2 * F + F + c], y = ldx(gy, in16(rnd), r * F + c);
      sto(gh, rnd, r * 2 * F + c, y * g1);
      sto(gh, rnd, r * 2 * F + F + c, y * g2);
    }
  }
  return nullptr;
}
PBK pbk_softmax_fwd(float* S, long rows, int cols, long ld, int rnd, pb_stream) {
  for (long r = 0; r < rows; ++r) {
    float* p = S + r * ld;
    float mx = -INFINITY;
    for (int c = 0; c < cols; ++c) mx = std::max(mx, p[c]);
    double s = 0;
    for (int c = 0; c < cols; ++c) s += std::exp(p[c] - mx);
    for (int c = 0; c < cols; ++c) p[c] = mr((float)(std::exp(p[c] - mx) / s), rnd);
    for (long c = cols; c < ld; ++c) p[c] = 0.f;
  }
  return nullptr;
}
PBK pbk_softmax_lin(const float* P, long rows_p, float* dS, int nb, int cols, long ld, int rnd, int k_slot, long p_stride, pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  for (long r = 0; r < rows_p * nb; ++r) {
    const float* p = P + ((r / rows_p) / k_slot) * p_stride + (r % rows_p) * ld;
    float* d = dS + r * ld;
    double dot = 0;
    for (int c = 0; c < cols; ++c) dot += (double)p[c] * d[c];
    for (int c = 0; c < cols; ++c) d[c] = mr(p[c] * (d[c] - (float)dot), rnd);
  }
  return nullptr;
}
PBK pbk_attn_delta(const float* go, long ldg, const float* o0, long ldo, int nb, int N, int H, int d, float* delta, int io, int k_slot,
                   long p_stride, pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  for (long b = 0; b < nb; ++b)
    for (int h = 0; h < H; ++h) {
      const float* o = o0 + (b / k_slot) * p_stride;
      for (int i = 0; i < N; ++i) {
        double s = 0;
        for (int c = 0; c < d; ++c) s += (double)ldx(go, in16(io), (b * N + i) * ldg + h * d + c) * o[(long)i * ldo + h * d + c];
        delta[(b * H + h) * N + i] = (float)s;
      }
    }
  return nullptr;
}
PBK pbk_attn_ds(const float* P0, float* dP, const float* delta, float scale, int nb, int H, int rows, int cols, long ld,
                int col_mode, int rnd, int k_slot, long p_stride, pb_stream) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  for (long b = 0; b < nb; ++b)
    for (int h = 0; h < H; ++h)
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
          const long bh = b * H + h;
          const float* P = P0 + (b / k_slot) * p_stride;
          const float dl = col_mode ? delta[bh * cols + c] : delta[bh * rows + r];
          float* p = dP + (bh * rows + r) * ld + c;
          *p = mr(scale * P[((long)h * rows + r) * ld + c] * (*p - dl), rnd);
        }
  return nullptr;
}
PBK pbk_attn_lin_supported(int d, int Mr, int Nc) {
  if (d % 4 || d < 8 || d > 96) return "attn_lin: head dim must be a multiple of 4 in [8, 96]";
  return nullptr;
}
PBK pbk_attn_lin(const PbAttnLin* ap, pb_stream) {
  const PbAttnLin& a = *ap;
  const int p16 = a.p16 ? 1 : 0, s16 = a.s16 ? 1 : 0;      // Pm / C1 / C2 halves (Pm pre-scaled); S operands and D / D2 halves
  if (s16 && !p16) return "attn_lin: fp16 S operands need the fp16 probability path";
  const float inv_ps = p16 ? 1.f / a.p_scale : 1.f;
  auto opnd = [&](float v) { return p16 ? v : trunc_tf32(v); };           // operand of an accumulating product
  const int k_slot = (a.k_slot > 0 && a.k_slot < a.nb) ? a.k_slot : a.nb;
  const bool slots = a.nb / k_slot > 1;
  for (long b = 0; b < a.nb; ++b)
    for (long h = 0; h < a.nh; ++h) {
      // primal operands of this tangent's problem (byte stride p_stride per problem)
      const long sb = slots ? (b / k_slot) * a.p_stride : 0;
      const char* Pm = reinterpret_cast<const char*>(a.Pm) + sb;
      const char* C1p = reinterpret_cast<const char*>(a.C1) + sb;
      const float* Op = a.O ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.O) + sb) : nullptr;
#pragma omp parallel for schedule(static)
      for (int r = 0; r < a.Mr; ++r) {
        std::vector<float> T(a.Nc);
        double rs = 0;
        for (int c = 0; c < a.Nc; ++c) {
          float s = 0.f;
          for (int sg = 0; sg < a.nseg; ++sg) {
            const long a0 = b * a.seg[sg].sAb + h * a.seg[sg].sAh + (long)r * a.seg[sg].lda;
            const long b0 = b * a.seg[sg].sBb + h * a.seg[sg].sBh + (long)c * a.seg[sg].ldb;
            const void* Ap = a.seg[sg].sAb == 0 ? static_cast<const void*>(static_cast<const char*>(a.seg[sg].A) + sb) : a.seg[sg].A;
            const void* Bp = a.seg[sg].sBb == 0 ? static_cast<const void*>(static_cast<const char*>(a.seg[sg].B) + sb) : a.seg[sg].B;
            for (int k = 0; k < a.d; ++k)
              s += s16 ? ldx(Ap, 1, a0 + k) * ldx(Bp, 1, b0 + k)
                       : trunc_tf32(ldx(Ap, 0, a0 + k)) * trunc_tf32(ldx(Bp, 0, b0 + k));
          }
          float dl = 0.f;
          if (a.delta && a.delta_mode == 1) dl = a.delta[(b * a.nh + h) * a.Mr + r];
          if (a.delta && a.delta_mode == 2) dl = a.delta[(b * a.nh + h) * a.Nc + c];
          const float t = ldx(Pm, p16, h * a.sPh + (long)r * a.ldp + c) * (a.alpha1 * s - dl);
          T[c] = p16 ? (float)(h16)std::min(std::max(t, -65504.f), 65504.f) : rna(t);
          rs += T[c];
        }
        rs *= inv_ps;
        for (int n = 0; n < a.d; ++n) {
          double acc = 0;
          const long c10 = h * a.sCh + (long)n * a.ldc;
          for (int c = 0; c < a.Nc; ++c) acc += (double)T[c] * opnd(ldx(C1p, p16, c10 + c));
          double e2 = 0;
          const long o2 = b * a.sD2b + (long)r * a.ldd2 + h * a.d + n;
          if (a.C2) {
            const long c20 = b * a.sC2b + h * a.sC2h + (long)n * a.ldc2, p0 = h * a.sPh + (long)r * a.ldp;
            for (int c = 0; c < a.Nc; ++c) e2 += (double)opnd(ldx(Pm, p16, p0 + c)) * opnd(ldx(a.C2, p16, c20 + c));
            if (a.D2) stx(a.D2, s16, o2, s16 ? (float)e2 * inv_ps : mr((float)e2 * inv_ps, a.round_tf32));
            else acc += e2;
          }
          float v = a.alpha2 * inv_ps * (float)acc;
          if (a.want_rsum && Op) v -= (float)rs * Op[(long)r * a.ldo + h * a.d + n];
          if (a.R) v += a.beta * a.R[b * a.sRb + (long)r * a.ldr + h * a.d + n];
          stx(a.D, s16, b * a.sDb + (long)r * a.ldd + h * a.d + n, s16 ? v : mr(v, a.round_tf32));
        }
      }
    }
  return nullptr;
}
PBK pbk_ddim_step(const float* x, const float* eps, float a_t, float a_next, float* x_next, float* pred_x0, long n, pb_stream) {
  if (!(a_t > 0.f) || !(a_next >= 0.f) || a_t > 1.f || a_next > 1.f) return "ddim_step: alphas_cumprod must lie in (0, 1]";
  for (long i = 0; i < n; ++i) {
    const float p0 = (x[i] - std::sqrt(1.f - a_t) * eps[i]) / std::sqrt(a_t);
    if (pred_x0) pred_x0[i] = p0;
    x_next[i] = std::sqrt(a_next) * p0 + std::sqrt(1.f - a_next) * eps[i];
  }
  return nullptr;
}
PBK pbk_lincomb3(float* out, float a, const float* x, float b, const float* y, float c, const float* z, long n, pb_stream) {
  for (long i = 0; i < n; ++i) out[i] = a * x[i] + (y ? b * y[i] : 0.f) + (z ? c * z[i] : 0.f);
  return nullptr;
}
PBK pbk_timestep_embedding(float t, int dim, int flip, float shift, float* out, pb_stream) {
  const int half = dim / 2;
  for (int j = 0; j < half; ++j) {
    const float e = std::exp(-std::log(10000.f) * (float)j / ((float)half - shift));
    const float a = t * e;
    if (flip) { out[j] = std::cos(a); out[half + j] = std::sin(a); } else { out[j] = std::sin(a); out[half + j] = std::cos(a); }
  }
  return nullptr;
}
PBK pbk_gemv(const float* Wm, const float* x, const float* bias, int N, int K, int silu_in, int silu_out, float* y, pb_stream) {
  for (int n = 0; n < N; ++n) {
    double s = 0;
    for (int k = 0; k < K; ++k) s += (double)Wm[(long)n * K + k] * (silu_in ? silu_f(x[k]) : x[k]);
    float v = (float)s + (bias ? bias[n] : 0.f);
    y[n] = silu_out ? silu_f(v) : v;
  }
  return nullptr;
}
PBK pbk_pack_conv3x3(const float* w, int Co, int Ci, float* fwd, float* bwd, int rnd, pb_stream) {
  for (long co = 0; co < Co; ++co)
    for (long ci = 0; ci < Ci; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const float v = mr(w[(co * Ci + ci) * 9 + tap], rnd);
        if (fwd) fwd[(co * 9 + tap) * Ci + ci] = v;
        if (bwd) bwd[(ci * 9 + (8 - tap)) * Co + co] = v;
      }
  return nullptr;
}

PBK pbk_gram2(const float* Wm, const float* Vp, int k, long n, double* G, double* M, pb_stream) {
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      double a = 0, b = 0;
      for (long c = 0; c < n; ++c) { a += (double)Wm[i * n + c] * Wm[j * n + c]; if (Vp) b += (double)Wm[i * n + c] * Vp[j * n + c]; }
      G[i * k + j] = a; if (M) M[i * k + j] = b;
    }
  return nullptr;
}
PBK pbk_jacobi(const double* G, const double* M, int k, float* Rm, float* sv, pb_stream) {
  std::vector<double> A(G, G + k * k), X(k * k, 0.0);
  for (int i = 0; i < k; ++i) X[i * k + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) (i == j ? diag : off) += A[i * k + j] * A[i * k + j];
    if (off <= 1e-30 * diag) break;
    for (int p = 0; p < k - 1; ++p)
      for (int q = p + 1; q < k; ++q) {
        const double apq = A[p * k + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double tau = (A[q * k + q] - A[p * k + p]) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int i = 0; i < k; ++i) {
          const double aip = A[i * k + p], aiq = A[i * k + q];
          A[i * k + p] = c * aip - s * aiq; A[i * k + q] = s * aip + c * aiq;
          const double xip = X[i * k + p], xiq = X[i * k + q];
          X[i * k + p] = c * xip - s * xiq; X[i * k + q] = s * xip + c * xiq;
        }
        for (int i = 0; i < k; ++i) {
          const double api = A[p * k + i], aqi = A[q * k + i];
          A[p * k + i] = c * api - s * aqi; A[q * k + i] = s * api + c * aqi;
        }
      }
  }
  std::vector<int> order(k);
  for (int i = 0; i < k; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return A[a * k + a] > A[b * k + b]; });
  const double lmax = std::max(A[order[0] * k + order[0]], 0.0);
  for (int i = 0; i < k; ++i) {
    const int e = order[i];
    // numerically null direction of W (the Gram matrix squares the conditioning): a zero row with s = 0, not noise / 0
    if (!(A[e * k + e] > 1e-13 * lmax)) { for (int j = 0; j < k; ++j) Rm[i * k + j] = 0.f; sv[i] = 0.f; continue; }
    const double l = std::max(A[e * k + e], 1e-300), inv = 1.0 / std::sqrt(l);
    double dot = 0;
    if (M) for (int j = 0; j < k; ++j) dot += X[j * k + e] * M[j * k + i];
    const double sg = (M && dot < 0) ? -1.0 : 1.0;
    for (int j = 0; j < k; ++j) Rm[i * k + j] = (float)(sg * inv * X[j * k + e]);
    sv[i] = (float)std::sqrt(std::sqrt(l));
  }
  return nullptr;
}
PBK pbk_rotate(const float* Wm, const float* Rm, const float* Vp, int k, long n, float atol, float rtol, float* V, float* metrics,
               pb_stream) {
  double d2 = 0, viol = 0;
  for (long c = 0; c < n; ++c)
    for (int i = 0; i < k; ++i) {
      float v = 0.f;
      for (int j = 0; j < k; ++j) v += Rm[i * k + j] * Wm[j * n + c];
      V[i * n + c] = v;
      if (Vp) { const float d = v - Vp[i * n + c]; d2 += (double)d * d; if (std::fabs(d) > atol + rtol * std::fabs(v)) viol += 1; }
    }
  if (metrics) { metrics[0] = (float)d2; metrics[1] = (float)viol; }
  return nullptr;
}
