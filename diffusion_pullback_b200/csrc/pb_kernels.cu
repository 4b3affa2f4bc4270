// Hand-written sm_100a kernels for everything on the pullback hot path that is not a contraction:
// GroupNorm/LayerNorm/SiLU/GEGLU/softmax forward + their tangent (JVP) and cotangent (VJP)
// linearisations around the cached primal, layout moves, stride-2 im2col, tiny direct convs, the
// time embedding and the weight packers.  All HBM-bound: 128-bit accesses along the channel axis,
// grids sized to cover 148 SMs, reductions by warp shuffles.  See pb_kernels.h for semantics.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "pb_dev.cuh"
#include "pb_host_util.h"
#include "pb_kernels.h"
#include "pb_lin16.h"

const char* pb_gemm_launch(const PbGemm& g, cudaStream_t st);

namespace {

using namespace pbdev;

inline cudaStream_t S(pb_stream st) { return static_cast<cudaStream_t>(st); }
inline bool in16(int io) { return (io & PB_IN_F16) != 0; }
inline bool out16(int io) { return (io & PB_RND_MASK) == PB_OUT_F16; }
__host__ __device__ inline const __half* HP(const float* p) { return reinterpret_cast<const __half*>(p); }
__host__ __device__ inline __half* HP(float* p) { return reinterpret_cast<__half*>(p); }

__device__ __forceinline__ float maybe_round(float x, int r) { return r ? rna_tf32(x) : x; }
__device__ __forceinline__ float4 maybe_round4(float4 v, int r) {
  if (r) { v.x = rna_tf32(v.x); v.y = rna_tf32(v.y); v.z = rna_tf32(v.z); v.w = rna_tf32(v.w); }
  return v;
}
// Store four consecutive elements at element offset `off` (a multiple of 4) of a tensor produced for a GEMM-only consumer:
// rnd 0 fp32 as is, 1 fp32 RNA-rounded to TF32, 2 fp16 (the buffer then holds halves; element offsets are unchanged).
__device__ __forceinline__ void store_out4(float* out, long off, float4 v, int rnd) {
  if (rnd == 2) {
    uint2 h;
    *reinterpret_cast<__half2*>(&h.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(&h.y) = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + off) = h;
  } else {
    *reinterpret_cast<float4*>(out + off) = maybe_round4(v, rnd);
  }
}
// four consecutive elements of a tensor that holds floats or (h16) halves at the same element offsets
__device__ __forceinline__ float4 load_in4(const float* p, long off, bool h16) {
  if (h16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p) + off);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(p + off);
}
__device__ __forceinline__ float load_in1(const float* p, long off, bool h16) {
  return h16 ? __half2float(reinterpret_cast<const __half*>(p)[off]) : p[off];
}

// ------------------------------------------------------------------------------------------------
// data movement
// ------------------------------------------------------------------------------------------------
__global__ void copy2d_v4(float* __restrict__ dst, long ldd, const float* __restrict__ src, long lds, long rows,
                          int cols4, float beta, int rnd) {
  const long total = rows * cols4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / cols4; const int c = int(i % cols4) * 4;
    float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    float4* d = reinterpret_cast<float4*>(dst + r * ldd + c);
    if (beta != 0.f) { float4 o = *d; v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w; }
    *d = maybe_round4(v, rnd);
  }
}
__global__ void copy2d_s(float* __restrict__ dst, long ldd, const float* __restrict__ src, long lds, long rows,
                         int cols, float beta, int rnd) {
  const long total = rows * cols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / cols; const int c = int(i % cols);
    float v = src[r * lds + c];
    float* d = dst + r * ldd + c;
    *d = maybe_round(beta != 0.f ? v + beta * *d : v, rnd);
  }
}

// src_(b,h) [R][lds] -> dst_(b,h) [C][ldd];  blockIdx.z = b * nh + h
__global__ void transpose_k(float* __restrict__ dst, long ldd, long sbd, long shd, const float* __restrict__ src,
                            long lds, long sbs, long shs, int nh, int R, int C, float beta, int rnd, int src16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z / nh, h = blockIdx.z % nh;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const long soff = (long)b * sbs + (long)h * shs;      // src16: src, lds and the batch strides count halves
  float* d = dst + (long)b * sbd + (long)h * shd;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? load_in1(src, soff + (long)r * lds + c, src16 != 0) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      float v = tile[threadIdx.x][i];
      if (rnd == 2) {                                            // halves: dst, ldd and the batch strides count elements
        reinterpret_cast<__half*>(dst)[(long)b * sbd + (long)h * shd + (long)c * ldd + r] = __float2half_rn(v);
      } else {
        float* p = d + (long)c * ldd + r;
        if (beta != 0.f) v += beta * *p;
        *p = maybe_round(v, rnd);
      }
    }
  }
}

// fp16 -> fp16 transpose of narrow per-head matrices (C = head dim, a multiple of 8, <= 128): src_(b,h) [R][lds] halves ->
// dst_(b,h) [C][ldd] halves.  A block moves a tile of 64 rows: 16-byte loads along C, 16-byte stores of 8 consecutive rows
// along R (the generic kernel above moves 32 x 32 tiles one element per thread: 64-byte segments, and two column blocks of which
// the second is 80 % empty at C = 40 -- 0.75 TB/s).  blockIdx.y = b * nh + h
constexpr int TH_ROWS = 64;
__global__ void __launch_bounds__(256) transpose_heads16_k(__half* __restrict__ dst, long ldd, long sbd, long shd,
                                                           const __half* __restrict__ src, long lds, long sbs, long shs,
                                                           int nh, int R, int C) {
  extern __shared__ __half th_tile[];                          // [TH_ROWS][C + 2]
  const int P = C + 2;
  const int b = blockIdx.y / nh, h = blockIdx.y % nh;
  const int r0 = blockIdx.x * TH_ROWS;
  const __half* sp = src + (long)b * sbs + (long)h * shs;
  __half* dp = dst + (long)b * sbd + (long)h * shd;
  const int c8 = C >> 3;
  for (int i = threadIdx.x; i < TH_ROWS * c8; i += blockDim.x) {
    const int r = i / c8, cc = (i % c8) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < R) v = *reinterpret_cast<const uint4*>(sp + (long)(r0 + r) * lds + cc);
    uint32_t* t = reinterpret_cast<uint32_t*>(th_tile + r * P + cc);      // P and cc are even: 4-byte aligned
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * (TH_ROWS / 8); i += blockDim.x) {
    const int c = i / (TH_ROWS / 8), j = (i % (TH_ROWS / 8)) * 8;
    __align__(16) __half e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = th_tile[(j + k) * P + c];
    __half* o = dp + (long)c * ldd + r0 + j;
    if (r0 + j + 8 <= R) {
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(e);
    } else {
      for (int k = 0; k < 8 && r0 + j + k < R; ++k) o[k] = e[k];
    }
  }
}

__global__ void upsample2x_k(const float4* __restrict__ x, int nb, int H, int W, int C4, float4* __restrict__ y,
                             int rnd) {
  const long total = (long)nb * 4 * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C4); t /= C4;
    const int ox = int(t % (2 * W)); t /= 2 * W;
    const int oy = int(t % (2 * H)); t /= 2 * H;
    const int b = int(t);
    store_out4(reinterpret_cast<float*>(y), 4 * i, x[(((long)b * H + oy / 2) * W + ox / 2) * C4 + c], rnd);
  }
}
__global__ void upsample2x_vjp_k(const float4* __restrict__ gy, int nb, int H, int W, int C4, float4* __restrict__ gx,
                                 float beta, int rnd) {
  const long total = (long)nb * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C4); t /= C4;
    const int x = int(t % W); t /= W;
    const int y = int(t % H); t /= H;
    const int b = int(t);
    float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float4 v = gy[(((long)b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C4 + c];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    if (beta != 0.f) { const float4 o = gx[i]; a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w; }
    gx[i] = maybe_round4(a, rnd);
  }
}
__global__ void to_f16_k(__half* __restrict__ dst, const float* __restrict__ src, size_t n4, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    store_out4(reinterpret_cast<float*>(dst), 4 * (long)i, v, 2);
  }
}
__global__ void ddim_step_k(const float* __restrict__ x, const float* __restrict__ eps, float c_x0, float c_eps0, float c_x,
                            float c_eps, float* __restrict__ x_next, float* __restrict__ pred_x0, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float xv = x[i], ev = eps[i];
    const float p0 = c_x0 * xv - c_eps0 * ev;             // (x - sqrt(1 - a_t) eps) / sqrt(a_t)
    if (pred_x0) pred_x0[i] = p0;
    x_next[i] = c_x * p0 + c_eps * ev;
  }
}
__global__ void lincomb3_k(float* __restrict__ out, float a, const float* x, float b, const float* y, float c, const float* z, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = a * x[i];
    if (y) v = fmaf(b, y[i], v);
    if (z) v = fmaf(c, z[i], v);
    out[i] = v;
  }
}
__global__ void round_tf32_k(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = rna_tf32(src[i]);
}

// ------------------------------------------------------------------------------------------------
// stride-2 3x3 conv helpers
// ------------------------------------------------------------------------------------------------
__global__ void im2col_s2_k(const float4* __restrict__ x, int nb, int H, int W, int C4, int pad, int Ho, int Wo,
                            float4* __restrict__ col, int rnd) {
  const long total = (long)nb * Ho * Wo * 9 * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C4); t /= C4;
    const int tap = int(t % 9); t /= 9;
    const int ox = int(t % Wo); t /= Wo;
    const int oy = int(t % Ho); t /= Ho;
    const int b = int(t);
    const int iy = 2 * oy + tap / 3 - pad, ix = 2 * ox + tap % 3 - pad;
    float4 v = make_float4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long)b * H + iy) * W + ix) * C4 + c];
    store_out4(reinterpret_cast<float*>(col), 4 * i, v, rnd);
  }
}
__global__ void col2im_s2_k(const float4* __restrict__ col, int nb, int H, int W, int C4, int pad, int Ho, int Wo,
                            float4* __restrict__ gx, float beta, int rnd) {
  const long total = (long)nb * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C4); t /= C4;
    const int ix = int(t % W); t /= W;
    const int iy = int(t % H); t /= H;
    const int b = int(t);
    float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        const float4 v = col[((((long)b * Ho + oy) * Wo + ox) * 9 + ky * 3 + kx) * C4 + c];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    if (beta != 0.f) { const float4 o = gx[i]; a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w; }
    gx[i] = maybe_round4(a, rnd);
  }
}

// thread per output element; for tiny Cin (conv_in)
__global__ void conv3x3_direct_thin_in(const float* __restrict__ x, int nb, int H, int W, int Cin,
                                       const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                       float* __restrict__ y, float beta, bool x16, bool y16) {
  const long total = (long)nb * H * W * Cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int co = int(t % Cout); t /= Cout;
    const int px = int(t % W); t /= W;
    const int py = int(t % H); t /= H;
    const int b = int(t);
    float acc = bias ? bias[co] : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = py + tap / 3 - 1, ix = px + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const long xo = (((long)b * H + iy) * W + ix) * Cin;
      const float* wp = w + ((long)co * 9 + tap) * Cin;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(load_in1(x, xo + ci, x16), wp[ci], acc);
    }
    if (beta != 0.f) acc += beta * load_in1(y, i, y16);
    if (y16) HP(y)[i] = __float2half_rn(acc); else y[i] = acc;
  }
}
// warp per output pixel, lanes split Cin; for tiny Cout (transpose of conv_in)
template <int MAXCO>
__global__ void conv3x3_direct_thin_out(const float* __restrict__ x, int nb, int H, int W, int Cin,
                                        const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                        float* __restrict__ y, float beta, bool x16, bool y16) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long total = (long)nb * H * W;
  for (long pix = warp; pix < total; pix += nwarps) {
    long t = pix;
    const int px = int(t % W); t /= W;
    const int py = int(t % H); t /= H;
    const int b = int(t);
    float acc[MAXCO];
#pragma unroll
    for (int co = 0; co < MAXCO; ++co) acc[co] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = py + tap / 3 - 1, ix = px + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const long xo = (((long)b * H + iy) * W + ix) * Cin;
      for (int ci = lane; ci < Cin; ci += 32) {
        const float xv = load_in1(x, xo + ci, x16);
#pragma unroll
        for (int co = 0; co < MAXCO; ++co)
          if (co < Cout) acc[co] = fmaf(xv, w[((long)co * 9 + tap) * Cin + ci], acc[co]);
      }
    }
#pragma unroll
    for (int co = 0; co < MAXCO; ++co) {
      if (co < Cout) {
        float v = warp_sum(acc[co]);
        if (lane == 0) {
          if (bias) v += bias[co];
          const long yo = pix * Cout + co;
          if (beta != 0.f) v += beta * load_in1(y, yo, y16);
          if (y16) HP(y)[yo] = __float2half_rn(v); else y[yo] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
// Per-(image, channel) sums over a chunk of pixels.  Threads own fixed channel quads so partials
// live in registers; pixel lanes are combined through shared memory, chunks through atomics.
//   MODE 0 (fwd):  s1 = sum (x - pivot_c), s2 = sum (x - pivot_c)^2,  pivot_c = x[b][0][c]
//   MODE 1 (jvp):  u = t                    ; s1 = sum u, s2 = sum xhat*u
//   MODE 2 (vjp):  u = t * act'(y) * gamma  ; s1 = sum u, s2 = sum xhat*u
struct GnGeom { int TP, QPT, R; };   // threads per pixel row, quads per thread, pixel lanes
static GnGeom gn_geom(int C) {
  const int Q = C / 4;
  int TP = 1;
  for (int d = 1; d <= Q && d <= 256; ++d)
    if (Q % d == 0) TP = d;
  GnGeom g;
  g.TP = TP; g.QPT = Q / TP; g.R = std::max(1, 256 / TP);
  return g;
}
constexpr int GN_MAX_QPT = 8;

template <int MODE>
__global__ void gn_sums_k(const float* __restrict__ xp, const float* __restrict__ mean, const float* __restrict__ rstd,
                          const float* __restrict__ gamma, const float* __restrict__ beta_, int HW, int C, int G,
                          int silu, const float* __restrict__ t, int TP, int QPT, int R, int pix_per_block,
                          float* __restrict__ sums /* [nb][C][2] */) {
  extern __shared__ float sh[];                 // [R][C][2]
  const int b = blockIdx.y;
  const int tp = threadIdx.x % TP, pl = threadIdx.x / TP;
  const int cpg = C / G;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const float* xb = MODE == 0 ? xp + (long)b * HW * C : xp;       // primal is a single image in lin modes
  const float* tb = MODE == 0 ? nullptr : t + (long)b * HW * C;
  const float* mb = MODE == 0 ? nullptr : mean;
  const float* rb = MODE == 0 ? nullptr : rstd;
  float a1[GN_MAX_QPT][4], a2[GN_MAX_QPT][4];
#pragma unroll
  for (int j = 0; j < GN_MAX_QPT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { a1[j][e] = 0.f; a2[j][e] = 0.f; }
  for (int pix = p0 + pl; pix < p1; pix += R) {
#pragma unroll
    for (int j = 0; j < GN_MAX_QPT; ++j) {
      if (j >= QPT) break;
      const int c = 4 * (tp + j * TP);
      const float4 xv = *reinterpret_cast<const float4*>(xb + (long)pix * C + c);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      if (MODE == 0) {
        const float4 pv = *reinterpret_cast<const float4*>(xb + c);
        const float ps[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float d = xs[e] - ps[e]; a1[j][e] += d; a2[j][e] = fmaf(d, d, a2[j][e]); }
      } else {
        const float4 tv = *reinterpret_cast<const float4*>(tb + (long)pix * C + c);
        const float ts[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int g = (c + e) / cpg;
          const float xh = (xs[e] - mb[g]) * rb[g];
          float u = ts[e];
          if (MODE == 2) {
            const float ga = gamma[c + e];
            u *= silu ? ga * silu_d(fmaf(ga, xh, beta_[c + e])) : ga;
          }
          a1[j][e] += u; a2[j][e] = fmaf(xh, u, a2[j][e]);
        }
      }
    }
  }
  // combine pixel lanes
#pragma unroll
  for (int j = 0; j < GN_MAX_QPT; ++j) {
    if (j >= QPT) break;
    const int c = 4 * (tp + j * TP);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sh[((long)pl * C + c + e) * 2 + 0] = a1[j][e];
      sh[((long)pl * C + c + e) * 2 + 1] = a2[j][e];
    }
  }
  __syncthreads();
  // per-chunk partials, combined in a fixed order by the finalize kernels (no atomics: runs are bit-reproducible)
  float* part = sums + ((long)blockIdx.x * gridDim.y + b) * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += sh[(long)r * C * 2 + i];
    part[i] = s;
  }
}

// deterministic block reduction of two doubles (blockDim.x = 128)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[2 * w] = a; sh[2 * w + 1] = b; }
  __syncthreads();
  a = sh[0] + sh[2] + sh[4] + sh[6];
  b = sh[1] + sh[3] + sh[5] + sh[7];
}

// one block (128 threads) per (image, group): sums the per-chunk partials in a fixed order
__global__ void gn_finalize_fwd_k(const float* __restrict__ x, const float* __restrict__ part, int chunks, int nb, int HW,
                                  int C, int G, float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ double sh[8];
  __shared__ double s_mg;
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const double n = HW;
  // pass 1: group mean from per-channel pivoted sums
  double msum = 0.0, dummy = 0.0;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    double s1 = 0.0;
    for (int k = 0; k < chunks; ++k) s1 += part[(((long)k * nb + b) * C + c) * 2];
    msum += (double)x[(long)b * HW * C + c] + s1 / n;
  }
  block_sum2(msum, dummy, sh);
  if (threadIdx.x == 0) s_mg = msum / cpg;
  __syncthreads();
  const double mg = s_mg;
  double m2 = 0.0; dummy = 0.0;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = g * cpg + i;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < chunks; ++k) {
      s1 += part[(((long)k * nb + b) * C + c) * 2];
      s2 += part[(((long)k * nb + b) * C + c) * 2 + 1];
    }
    const double mc = (double)x[(long)b * HW * C + c] + s1 / n;
    m2 += (s2 - s1 * s1 / n) + n * (mc - mg) * (mc - mg);
  }
  __syncthreads();
  block_sum2(m2, dummy, sh);
  if (threadIdx.x == 0) {
    const double var = m2 / (n * cpg);
    mean[b * G + g] = (float)mg;
    rstd[b * G + g] = (float)(1.0 / sqrt(var + (double)eps));
  }
}
// tmp[b][g] = (mean_g(u), mean_g(xhat*u))
__global__ void gn_finalize_lin_k(const float* __restrict__ part, int chunks, int nb, int HW, int C, int G,
                                  float* __restrict__ tmp) {
  __shared__ double sh[8];
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  double s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < cpg * chunks; i += blockDim.x) {
    const int k = i / cpg, c = g * cpg + i % cpg;
    s1 += part[(((long)k * nb + b) * C + c) * 2];
    s2 += part[(((long)k * nb + b) * C + c) * 2 + 1];
  }
  block_sum2(s1, s2, sh);
  if (threadIdx.x == 0) {
    const double n = (double)HW * cpg;
    tmp[((long)b * G + g) * 2] = (float)(s1 / n);
    tmp[((long)b * G + g) * 2 + 1] = (float)(s2 / n);
  }
}

// One-launch GroupNorm linearisation: a block owns one (image, group) pair, sums over it (pass 1, HBM), and applies
// (pass 2, the same bytes again from L2).  Replaces partial sums + finalize + apply (three launches and a round trip
// of per-chunk partials) whenever there are enough (image, group) pairs to fill the machine.  Threads are laid out
// [pixel lane][channel vector] so that no index needs a division inside the loops; V = 4 / 2 / 1 channels per access.
// SPLIT > 1: the pixels of one (image, group) pair are shared by a cluster of SPLIT blocks (grid.z) that exchange their
// partial sums through distributed shared memory in rank order (deterministic), for the layers with many pixels per group.
template <int MODE, int V, int SPLIT>
__device__ __forceinline__ void gn_lin_group_body(const float* __restrict__ xp, const float* __restrict__ mean,
                                                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta_, int HW, int C, int G, int silu,
                                                  const float* __restrict__ t, float* __restrict__ out, float acc, int rnd) {
  __shared__ double sh[2][16];
  __shared__ double s_part[2];
  __shared__ float s_m[2];
  const int g = blockIdx.x, b = blockIdx.y;
  const int pix_lo = SPLIT > 1 ? int(blockIdx.z) * (HW / SPLIT) : 0;
  const int pix_hi = SPLIT > 1 ? pix_lo + HW / SPLIT : HW;
  const int cpg = C / G, cv = cpg / V;
  const int rows = blockDim.x / cv;                      // pixel lanes
  const int tid = threadIdx.x;
  const bool active = tid < rows * cv;
  const int j = active ? (tid % cv) * V : 0, p0 = tid / cv;
  const float mu = mean[g], rs = rstd[g];
  const float* xg = xp + g * cpg + j;
  const float* tg = t + (long)b * HW * C + g * cpg + j;
  float* og = out + (long)b * HW * C + g * cpg + j;
  float ga[V], be[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { ga[k] = gamma[g * cpg + j + k]; be[k] = beta_[g * cpg + j + k]; }
  auto load = [&](const float* q_, float (&v)[V]) {
    if constexpr (V == 4) { const float4 q = *reinterpret_cast<const float4*>(q_); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else if constexpr (V == 2) { const float2 q = *reinterpret_cast<const float2*>(q_); v[0] = q.x; v[1] = q.y; }
    else v[0] = q_[0];
  };
  const long step = (long)rows * C;                      // pointer increment per loop iteration (no index math inside)
  const long first = (long)(pix_lo + p0) * C;
  float s1 = 0.f, s2 = 0.f;
  if (active) {
    const float* xq = xg + first;
    const float* tq = tg + first;
#pragma unroll 4
    for (int p = pix_lo + p0; p < pix_hi; p += rows, xq += step, tq += step) {   // unrolled: four pixels' loads in flight
      float xv[V], tv[V];
      load(xq, xv); load(tq, tv);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float xh = (xv[k] - mu) * rs;
        float u = tv[k];
        if (MODE == 1) u *= silu ? ga[k] * silu_d(fmaf(ga[k], xh, be[k])) : ga[k];
        s1 += u; s2 = fmaf(xh, u, s2);
      }
    }
  }
  double d1 = s1, d2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
  if ((tid & 31) == 0) { sh[0][tid >> 5] = d1; sh[1][tid >> 5] = d2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; c += sh[1][w]; }
    s_part[0] = a; s_part[1] = c;
  }
  if constexpr (SPLIT > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                        // every block's partial is in its shared memory
    if (tid == 0) {
      double a = 0, c = 0;
      for (int r = 0; r < SPLIT; ++r) {
        const double* rp = cluster.map_shared_rank(s_part, r);
        a += rp[0]; c += rp[1];
      }
      const double n = (double)HW * cpg;
      s_m[0] = (float)(a / n); s_m[1] = (float)(c / n);
    }
    cluster.sync();                                        // nobody leaves (or reuses s_part) while a peer may still read it
  } else {
    __syncthreads();
    if (tid == 0) {
      const double n = (double)HW * cpg;
      s_m[0] = (float)(s_part[0] / n); s_m[1] = (float)(s_part[1] / n);
    }
    __syncthreads();
  }
  const float m1 = s_m[0], m2 = s_m[1];
  if (!active) return;
  const float* xq = xg + first;
  const float* tq = tg + first;
  float* op = og + first;
  __half* hp = reinterpret_cast<__half*>(out) + ((long)b * HW * C + g * cpg + j) + first;
#pragma unroll 4
  for (int p = pix_lo + p0; p < pix_hi; p += rows, xq += step, tq += step, op += step, hp += step) {
    float xv[V], tv[V], o[V];
    load(xq, xv); load(tq, tv);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float xh = (xv[k] - mu) * rs;
      const float f = silu ? ga[k] * silu_d(fmaf(ga[k], xh, be[k])) : ga[k];
      o[k] = MODE == 0 ? f * rs * (tv[k] - m1 - xh * m2) : rs * (tv[k] * f - m1 - xh * m2);
    }
    if (acc != 0.f) {
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] += acc * op[k];
    }
    if (rnd == 2) {
      if constexpr (V == 4) {
        uint2 hv;
        *reinterpret_cast<__half2*>(&hv.x) = __floats2half2_rn(o[0], o[1]);
        *reinterpret_cast<__half2*>(&hv.y) = __floats2half2_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(hp) = hv;
      } else if constexpr (V == 2) {
        *reinterpret_cast<__half2*>(hp) = __floats2half2_rn(o[0], o[1]);
      } else {
        hp[0] = __float2half_rn(o[0]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = maybe_round(o[k], rnd);
      if constexpr (V == 4) *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
      else if constexpr (V == 2) *reinterpret_cast<float2*>(op) = make_float2(o[0], o[1]);
      else op[0] = o[0];
    }
  }
}

template <int MODE, int V>
__global__ void __launch_bounds__(512) gn_lin_group_k(const float* __restrict__ xp, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta_, int HW, int C, int G, int silu,
                                                      const float* __restrict__ t, float* __restrict__ out, float acc, int rnd) {
  gn_lin_group_body<MODE, V, 1>(xp, mean, rstd, gamma, beta_, HW, C, G, silu, t, out, acc, rnd);
}
template <int MODE, int V, int SPLIT>
__global__ void __cluster_dims__(1, 1, SPLIT) __launch_bounds__(512)
gn_lin_group_cluster_k(const float* __restrict__ xp, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta_, int HW, int C, int G, int silu,
                       const float* __restrict__ t, float* __restrict__ out, float acc, int rnd) {
  gn_lin_group_body<MODE, V, SPLIT>(xp, mean, rstd, gamma, beta_, HW, C, G, silu, t, out, acc, rnd);
}

__global__ void gn_apply_fwd_k(const float* __restrict__ x, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                               const float* __restrict__ beta_, long total4, int HW, int C, int G, int silu, int rnd,
                               float* __restrict__ y) {
  const int C4 = C / 4, cpg = C / G;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int c = int(i % C4) * 4;
    const int b = int(i / ((long)HW * C4));
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (c + e) / cpg;
      float v = fmaf(gamma[c + e], (xs[e] - mean[b * G + g]) * rstd[b * G + g], beta_[c + e]);
      if (silu) v = silu_f(v);
      o[e] = maybe_round(v, rnd);
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

template <int MODE>
__global__ void gn_apply_lin_k(const float* __restrict__ xp, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                               const float* __restrict__ beta_, const float* __restrict__ tmp, long total4, int HW,
                               int C, int G, int silu, const float* __restrict__ t, float* __restrict__ out, float acc,
                               int rnd) {
  const int C4 = C / 4, cpg = C / G;
  const long per_img4 = (long)HW * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int c = int(i % C4) * 4;
    const int b = int(i / per_img4);
    const float4 xv = reinterpret_cast<const float4*>(xp)[i % per_img4];
    const float4 tv = reinterpret_cast<const float4*>(t)[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float ts[4] = {tv.x, tv.y, tv.z, tv.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (c + e) / cpg;
      const float rs = rstd[g];
      const float xh = (xs[e] - mean[g]) * rs;
      const float m1 = tmp[((long)b * G + g) * 2], m2 = tmp[((long)b * G + g) * 2 + 1];
      const float ga = gamma[c + e];
      const float f = silu ? ga * silu_d(fmaf(ga, xh, beta_[c + e])) : ga;
      float v;
      if (MODE == 0) v = f * rs * (ts[e] - m1 - xh * m2);
      else v = rs * (ts[e] * f - m1 - xh * m2);
      o[e] = v;
    }
    if (acc != 0.f) { const float4 p = reinterpret_cast<const float4*>(out)[i]; o[0] += acc * p.x; o[1] += acc * p.y; o[2] += acc * p.z; o[3] += acc * p.w; }
    store_out4(out, 4 * i, make_float4(o[0], o[1], o[2], o[3]), rnd);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per token row
// ------------------------------------------------------------------------------------------------
__global__ void ln_fwd_k(const float* __restrict__ x, long rows, int C, const float* __restrict__ gamma,
                         const float* __restrict__ beta_, float eps, float* __restrict__ y, float* __restrict__ mean,
                         float* __restrict__ rstd, int rnd) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long r = warp; r < rows; r += nw) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) { const float4 v = *reinterpret_cast<const float4*>(xr + c); s += v.x + v.y + v.z + v.w; }
    const float m = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      q += (v.x - m) * (v.x - m) + (v.y - m) * (v.y - m) + (v.z - m) * (v.z - m) + (v.w - m) * (v.w - m);
    }
    const float rs = rsqrtf(warp_sum(q) / C + eps);
    if (lane == 0) { mean[r] = m; rstd[r] = rs; }
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 bb = *reinterpret_cast<const float4*>(beta_ + c);
      float4 o = make_float4(fmaf(g.x, (v.x - m) * rs, bb.x), fmaf(g.y, (v.y - m) * rs, bb.y),
                             fmaf(g.z, (v.z - m) * rs, bb.z), fmaf(g.w, (v.w - m) * rs, bb.w));
      *reinterpret_cast<float4*>(y + r * C + c) = maybe_round4(o, rnd);
    }
  }
}
template <int MODE>
__global__ void ln_lin_k(const float* __restrict__ xp, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, long rows_p, int C, const float* __restrict__ t, long rows,
                         float* __restrict__ out, float acc, int rnd) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long r = warp; r < rows; r += nw) {
    const long rp = r % rows_p;
    const float* xr = xp + rp * C;
    const float* tr = t + r * C;
    const float m = mean[rp], rs = rstd[rp];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      float4 tv = *reinterpret_cast<const float4*>(tr + c);
      if (MODE == 1) { const float4 g = *reinterpret_cast<const float4*>(gamma + c); tv.x *= g.x; tv.y *= g.y; tv.z *= g.z; tv.w *= g.w; }
      s1 += tv.x + tv.y + tv.z + tv.w;
      s2 += (xv.x - m) * rs * tv.x + (xv.y - m) * rs * tv.y + (xv.z - m) * rs * tv.z + (xv.w - m) * rs * tv.w;
    }
    const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const float4 tv = *reinterpret_cast<const float4*>(tr + c);
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w}, gs[4] = {g.x, g.y, g.z, g.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (xs[e] - m) * rs;
        o[e] = MODE == 0 ? gs[e] * rs * (ts[e] - m1 - xh * m2) : rs * (ts[e] * gs[e] - m1 - xh * m2);
      }
      if (acc != 0.f) { const float4 p = *reinterpret_cast<const float4*>(out + r * C + c); o[0] += acc * p.x; o[1] += acc * p.y; o[2] += acc * p.z; o[3] += acc * p.w; }
      store_out4(out, r * C + c, make_float4(o[0], o[1], o[2], o[3]), rnd);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GEGLU
// ------------------------------------------------------------------------------------------------
// ff1 weight [2 F][cols] halves -> blocks of 64 rows [32 rows of the a half | the 32 matching rows of the gate half] (PbGemm::gg)
__global__ void interleave_rows16_k(uint4* __restrict__ dst, const uint4* __restrict__ src, int F, int c8) {
  const long total = 2L * F * c8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int row = int(i / c8), c = int(i % c8);
    const int j = row >> 6, e = row & 63;
    const int srow = e < 32 ? 32 * j + e : F + 32 * j + (e - 32);
    dst[i] = src[(long)srow * c8 + c];
  }
}
// prepare: [a | g] -> [gelu(g) | a gelu'(g)] in place, the factors of the linearisation (read by geglu_jvp_k / geglu_vjp_k)
__global__ void geglu_fwd_k(float* __restrict__ h, long rows, int F, float* __restrict__ y, int rnd, int prepare) {
  const int F4 = F / 4;
  const long total = rows * F4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / F4; const int c = int(i % F4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(h + r * 2 * F + c);
    const float4 g = *reinterpret_cast<const float4*>(h + r * 2 * F + F + c);
    const float4 gf = make_float4(gelu_f(g.x), gelu_f(g.y), gelu_f(g.z), gelu_f(g.w));
    float4 o = make_float4(a.x * gf.x, a.y * gf.y, a.z * gf.z, a.w * gf.w);
    *reinterpret_cast<float4*>(y + r * F + c) = maybe_round4(o, rnd);
    if (prepare) {
      *reinterpret_cast<float4*>(h + r * 2 * F + c) = gf;
      *reinterpret_cast<float4*>(h + r * 2 * F + F + c) =
          make_float4(a.x * gelu_d(g.x), a.y * gelu_d(g.y), a.z * gelu_d(g.z), a.w * gelu_d(g.w));
    }
  }
}
// IN16: the tangent dh / gy holds halves; rnd 2: the result is stored as halves.  Problem slots: tangent image r / rows_p
// belongs to problem (r / rows_p) / k_slot, whose primal tensor starts p_stride floats after the previous problem's.
template <bool IN16>
__global__ void geglu_jvp_k(const float* __restrict__ hp, long rows_p, const float* __restrict__ dh, long rows, int F,
                            float* __restrict__ dy, int rnd, int k_slot, long p_stride) {
  // 32-bit index arithmetic (the launcher checks rows * F / 4 < 2^31): four 64-bit divisions per float4 cost more instructions than
  // the whole linearisation
  const unsigned F4 = F / 4, total = unsigned(rows) * F4, rp32 = unsigned(rows_p);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / F4; const int c = int(i - r * F4) * 4;
    const unsigned img = r / rp32;
    const float* hr = hp + (long)(img / unsigned(k_slot)) * p_stride + (long)(r - img * rp32) * 2 * F;
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(hr + c));          // gelu(g)
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(hr + F + c));      // a gelu'(g)
    const float4 da = load_in4(dh, (long)r * 2 * F + c, IN16), dg = load_in4(dh, (long)r * 2 * F + F + c, IN16);
    float4 o = make_float4(da.x * g1.x + g2.x * dg.x, da.y * g1.y + g2.y * dg.y, da.z * g1.z + g2.z * dg.z, da.w * g1.w + g2.w * dg.w);
    store_out4(dy, (long)r * F + c, o, rnd);
  }
}
template <bool IN16>
__global__ void geglu_vjp_k(const float* __restrict__ hp, long rows_p, const float* __restrict__ gy, long rows, int F,
                            float* __restrict__ gh, int rnd, int k_slot, long p_stride) {
  const unsigned F4 = F / 4, total = unsigned(rows) * F4, rp32 = unsigned(rows_p);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / F4; const int c = int(i - r * F4) * 4;
    const unsigned img = r / rp32;
    const float* hr = hp + (long)(img / unsigned(k_slot)) * p_stride + (long)(r - img * rp32) * 2 * F;
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(hr + c));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(hr + F + c));
    const float4 y = load_in4(gy, (long)r * F + c, IN16);
    float4 ga = make_float4(y.x * g1.x, y.y * g1.y, y.z * g1.z, y.w * g1.w);
    float4 gg = make_float4(y.x * g2.x, y.y * g2.y, y.z * g2.z, y.w * g2.w);
    store_out4(gh, (long)r * 2 * F + c, ga, rnd);
    store_out4(gh, (long)r * 2 * F + F + c, gg, rnd);
  }
}

// ------------------------------------------------------------------------------------------------
// softmax (rows of attention scores).  NT threads cooperate on one row; up to 4 float4 per thread
// are cached in registers so HBM is touched once per element, longer rows are re-read.
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ float row_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  if (NT == 32) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? sh[l] : 0.f;
  return warp_sum(t);
}
template <int NT>
__device__ __forceinline__ float row_reduce_max(float v, float* sh) {
  v = warp_max(v);
  if (NT == 32) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? sh[l] : -INFINITY;
  return warp_max(t);
}

// NT == 32: one warp per row, blockDim = 256 (8 rows per block); NT == 256: one block per row.
template <int NT>
__global__ void softmax_fwd_k(float* __restrict__ Sm, long rows, int cols, long ld, int rnd) {
  __shared__ float sh[8];
  const int tid = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
  const long row0 = NT == 32 ? (blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  const long rstep = NT == 32 ? (long)gridDim.x * (blockDim.x >> 5) : gridDim.x;
  for (long r = row0; r < rows; r += rstep) {
    float* p = Sm + r * ld;
    float mx = -INFINITY;
    for (int c = tid; c < cols; c += NT) mx = fmaxf(mx, p[c]);
    mx = row_reduce_max<NT>(mx, sh);
    float s = 0.f;
    for (int c = tid; c < cols; c += NT) s += __expf(p[c] - mx);
    s = row_reduce_sum<NT>(s, sh);
    const float inv = 1.f / s;
    for (int c = tid; c < cols; c += NT) p[c] = maybe_round(__expf(p[c] - mx) * inv, rnd);
    for (int c = cols + tid; c < ld; c += NT) p[c] = 0.f;
  }
}

template <int NT>
__global__ void softmax_lin_k(const float* __restrict__ P, long rows_p, float* __restrict__ dS, long rows, int cols,
                              long ld, int rnd, int k_slot, long p_stride) {
  __shared__ float sh[8];
  const int tid = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
  const long row0 = NT == 32 ? (blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  const long rstep = NT == 32 ? (long)gridDim.x * (blockDim.x >> 5) : gridDim.x;
  const int cols4 = (cols + 3) / 4;               // ld % 4 == 0 and ld >= cols, so float4 loads stay in the row
  for (long r = row0; r < rows; r += rstep) {
    const float4* pp = reinterpret_cast<const float4*>(P + ((r / rows_p) / k_slot) * p_stride + (r % rows_p) * ld);
    float4* dp = reinterpret_cast<float4*>(dS + r * ld);
    float4 pc[4], dc[4];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c4 = tid + j * NT;
      if (c4 < cols4) {
        pc[j] = pp[c4]; dc[j] = dp[c4];
        const int c = c4 * 4;
        if (c + 3 < cols) dot += pc[j].x * dc[j].x + pc[j].y * dc[j].y + pc[j].z * dc[j].z + pc[j].w * dc[j].w;
        else {
          dot += pc[j].x * dc[j].x;
          if (c + 1 < cols) dot += pc[j].y * dc[j].y;
          if (c + 2 < cols) dot += pc[j].z * dc[j].z;
        }
      }
    }
    for (int c4 = tid + 4 * NT; c4 < cols4; c4 += NT) {      // long rows: uncached tail
      const float4 a = pp[c4], b = dp[c4];
      const int c = c4 * 4;
      dot += a.x * b.x;
      if (c + 1 < cols) dot += a.y * b.y;
      if (c + 2 < cols) dot += a.z * b.z;
      if (c + 3 < cols) dot += a.w * b.w;
    }
    dot = row_reduce_sum<NT>(dot, sh);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c4 = tid + j * NT;
      if (c4 < cols4) {
        float4 o = make_float4(pc[j].x * (dc[j].x - dot), pc[j].y * (dc[j].y - dot), pc[j].z * (dc[j].z - dot),
                               pc[j].w * (dc[j].w - dot));
        dp[c4] = maybe_round4(o, rnd);
      }
    }
    for (int c4 = tid + 4 * NT; c4 < cols4; c4 += NT) {
      const float4 a = pp[c4], b = dp[c4];
      dp[c4] = maybe_round4(make_float4(a.x * (b.x - dot), a.y * (b.y - dot), a.z * (b.z - dot), a.w * (b.w - dot)), rnd);
    }
  }
}

// thread per (tangent, row, head): heads are contiguous along a row, so a warp reads whole rows with 128-bit loads
__global__ void attn_delta_k(const float* __restrict__ go, long ldg, const float* __restrict__ o, long ldo, int nb,
                             int N, int H, int d, float* __restrict__ delta, bool go16, int k_slot, long p_stride) {
  const long total = (long)nb * N * H;
  const int d4 = d / 4;
  for (long w = blockIdx.x * (long)blockDim.x + threadIdx.x; w < total; w += (long)gridDim.x * blockDim.x) {
    long t = w;
    const int h = int(t % H); t /= H;
    const int i = int(t % N); t /= N;
    const int b = int(t);
    const long goff = ((long)b * N + i) * ldg + h * d;
    const float4* oo = reinterpret_cast<const float4*>(o + (long)(b / k_slot) * p_stride + (long)i * ldo + h * d);
    float s = 0.f;
    for (int c = 0; c < d4; ++c) {
      const float4 a = load_in4(go, goff + 4 * c, go16), q = oo[c];
      s = fmaf(a.x, q.x, fmaf(a.y, q.y, fmaf(a.z, q.z, fmaf(a.w, q.w, s))));
    }
    delta[((long)b * H + h) * N + i] = s;
  }
}
__global__ void attn_ds_k(const float* __restrict__ P, float* __restrict__ dP, const float* __restrict__ delta,
                          float scale, int nb, int H, int rows, int cols, long ld, int col_mode, int rnd, int k_slot, long p_stride4) {
  const long ld4 = ld / 4;
  const long per_b = (long)H * rows * ld4;
  const long total = (long)nb * per_b;
  const int ndelta = col_mode ? cols : rows;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = int(i % ld4) * 4;
    const long rr = i / ld4;                  // (b*H + h)*rows + r
    const int r = int(rr % rows);
    const long bh = rr / rows;
    const float4 p = reinterpret_cast<const float4*>(P)[((i / per_b) / k_slot) * p_stride4 + i % per_b];
    float4 v = reinterpret_cast<float4*>(dP)[i];
    const float* dl = delta + bh * ndelta;
    float d0, d1, d2, d3;
    if (col_mode) {
      d0 = c < cols ? dl[c] : 0.f; d1 = c + 1 < cols ? dl[c + 1] : 0.f;
      d2 = c + 2 < cols ? dl[c + 2] : 0.f; d3 = c + 3 < cols ? dl[c + 3] : 0.f;
    } else {
      d0 = d1 = d2 = d3 = dl[r];
    }
    v.x = scale * p.x * (v.x - d0); v.y = scale * p.y * (v.y - d1);
    v.z = scale * p.z * (v.z - d2); v.w = scale * p.w * (v.w - d3);
    reinterpret_cast<float4*>(dP)[i] = maybe_round4(v, rnd);
  }
}

// ------------------------------------------------------------------------------------------------
// time embedding / gemv / packing
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_k(float t, int dim, int flip, float shift, float* __restrict__ out) {
  const int half = dim / 2;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float e = expf(-logf(10000.f) * (float)j / ((float)half - shift));
    const float a = t * e;
    const float s = sinf(a), c = cosf(a);
    if (flip) { out[j] = c; out[half + j] = s; } else { out[j] = s; out[half + j] = c; }
  }
}
__global__ void gemv_k(const float* __restrict__ Wm, const float* __restrict__ x, const float* __restrict__ bias,
                       int N, int K, int silu_in, int silu_out, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < N; n += nw) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) {
      float xv = x[k];
      if (silu_in) xv = silu_f(xv);
      s = fmaf(Wm[(long)n * K + k], xv, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (bias) s += bias[n];
      y[n] = silu_out ? silu_f(s) : s;
    }
  }
}
__global__ void pack_conv3x3_k(const float* __restrict__ w, int Co, int Ci, float* __restrict__ fwd,
                               float* __restrict__ bwd, int rnd) {
  const long total = (long)Co * Ci * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int tap = int(t % 9); t /= 9;
    const int ci = int(t % Ci); t /= Ci;
    const int co = int(t);
    const float v = maybe_round(w[i], rnd);
    if (fwd) fwd[((long)co * 9 + tap) * Ci + ci] = v;
    if (bwd) bwd[((long)ci * 9 + (8 - tap)) * Co + co] = v;
  }
}


// conv_in and its transpose, register blocked over a STRIP of 8 pixels of an image row (the two "thin" kernels above are the
// generic fallback).  Both are 1.2 GFMA at 25 x 64 x 64 pixels, 33 us of fp32 FMA issue; one pixel per loop iteration with the
// filter re-read from shared memory per input (thin-out) or 64-bit index arithmetic per pixel (thin-in) ran them at 430-500 us.
// thin-in: each thread owns TWO output channels (its 2 x 9 x CIN filter taps stay in registers); the 3 x 10 input patch of a
// strip is 30 broadcast loads per thread, 576 FMAs.  w: [Cout][9][CIN].
constexpr int THIN_PW = 8;
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_thin_in_reg_k(const float* __restrict__ x, int nb, int H, int W,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int Cout, float* __restrict__ y, float beta, bool x16, bool y16) {
  const int cp = blockIdx.y * blockDim.x + threadIdx.x;       // output-channel pair
  if (2 * cp >= Cout) return;
  float wr[2][9 * CIN];
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int i = 0; i < 9 * CIN; ++i) wr[e][i] = w[(long)(2 * cp + e) * 9 * CIN + i];
  const float b0 = bias ? bias[2 * cp] : 0.f, b1 = bias ? bias[2 * cp + 1] : 0.f;
  const int spr = (W + THIN_PW - 1) / THIN_PW;                // strips per image row
  const int total = nb * H * spr;
  for (int sidx = blockIdx.x; sidx < total; sidx += gridDim.x) {
    const int sx = sidx % spr, row = sidx / spr;              // row = b * H + py
    const int py = row % H, px0 = sx * THIN_PW;
    float a0[THIN_PW], a1[THIN_PW];
#pragma unroll
    for (int q = 0; q < THIN_PW; ++q) { a0[q] = b0; a1[q] = b1; }
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      if (py + dy < 0 || py + dy >= H) continue;
      const long rbase = (long)(row + dy) * W;
      float xv[THIN_PW + 2][CIN];
#pragma unroll
      for (int i = 0; i < THIN_PW + 2; ++i) {
        const int xx = px0 - 1 + i;
        const bool ok = xx >= 0 && xx < W;
        if (CIN == 4) {
          const float4 v = ok ? load_in4(x, (rbase + xx) * 4, x16) : make_float4(0.f, 0.f, 0.f, 0.f);
          xv[i][0] = v.x; xv[i][1] = v.y; xv[i][2] = v.z; xv[i][CIN - 1] = v.w;
        } else {
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) xv[i][ci] = ok ? load_in1(x, (rbase + xx) * CIN + ci, x16) : 0.f;
        }
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = (dy + 1) * 3 + dx;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float w0 = wr[0][tap * CIN + ci], w1 = wr[1][tap * CIN + ci];
#pragma unroll
          for (int q = 0; q < THIN_PW; ++q) {
            a0[q] = fmaf(xv[q + dx][ci], w0, a0[q]);
            a1[q] = fmaf(xv[q + dx][ci], w1, a1[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < THIN_PW; ++q) {
      if (px0 + q >= W) break;
      const long yo = ((long)row * W + px0 + q) * Cout + 2 * cp;
      float o0 = a0[q], o1 = a1[q];
      if (y16) {
        __half2* yp = reinterpret_cast<__half2*>(HP(y) + yo);
        if (beta != 0.f) { const float2 o = __half22float2(*yp); o0 += beta * o.x; o1 += beta * o.y; }
        *yp = __floats2half2_rn(o0, o1);
      } else {
        float2* yp = reinterpret_cast<float2*>(y + yo);
        if (beta != 0.f) { const float2 o = *yp; o0 += beta * o.x; o1 += beta * o.y; }
        *yp = make_float2(o0, o1);
      }
    }
  }
}
// thin-out: the whole filter ([COUT <= 4][9][Cin], 46 KB for SD) sits in shared memory; one warp per strip of 8 pixels,
// lanes split the input channels in float4 quads: a filter quad read from shared memory serves the 8 pixels (32 FMAs per
// LDS.128), and the 8 x COUT partial sums are reduced across the warp by one 31-shuffle transpose-reduction that leaves
// lane l with output (pixel l / 4, channel l % 4) -- one coalesced store.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_thin_out_smem_k(const float* __restrict__ x, int nb, int H, int W, int Cin,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               float* __restrict__ y, float beta, bool x16, bool y16) {
  extern __shared__ float4 wsm[];                              // [COUT][9][Cin / 4]
  const int C4 = Cin / 4;
  for (int i = threadIdx.x; i < COUT * 9 * C4; i += blockDim.x) wsm[i] = reinterpret_cast<const float4*>(w)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int spr = (W + THIN_PW - 1) / THIN_PW;
  const int total = nb * H * spr;
  for (int sidx = warp; sidx < total; sidx += nwarps) {
    const int sx = sidx % spr, row = sidx / spr;
    const int py = row % H, px0 = sx * THIN_PW;
    float acc[32];                                             // [pixel][4]
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int c4 = lane; c4 < C4; c4 += 32) {
#pragma unroll 1
      for (int dy = -1; dy <= 1; ++dy) {
        if (py + dy < 0 || py + dy >= H) continue;
        const long rbase = (long)(row + dy) * W;
        float4 xv[THIN_PW + 2];
#pragma unroll
        for (int i = 0; i < THIN_PW + 2; ++i) {
          const int xx = px0 - 1 + i;
          xv[i] = (xx >= 0 && xx < W) ? load_in4(x, (rbase + xx) * Cin + 4 * c4, x16) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int tap = (dy + 1) * 3 + dx;
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float4 wv = wsm[(co * 9 + tap) * C4 + c4];
#pragma unroll
            for (int q = 0; q < THIN_PW; ++q) {
              const float4 v = xv[q + dx];
              acc[q * 4 + co] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[q * 4 + co]))));
            }
          }
        }
      }
    }
    // transpose-reduction: after the step of stride s a lane keeps the half of its values whose index has bit s equal to its
    // own lane bit and adds the partner's; lane l ends with the warp total of value l
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float mine = up ? acc[i + s] : acc[i];
        const float other = up ? acc[i] : acc[i + s];
        acc[i] = mine + __shfl_xor_sync(0xffffffffu, other, s);
      }
    }
    const int q = lane >> 2, co = lane & 3;
    if (co < COUT && px0 + q < W) {
      float v = acc[0];
      if (bias) v += bias[co];
      const long yo = ((long)row * W + px0 + q) * COUT + co;
      if (beta != 0.f) v += beta * load_in1(y, yo, y16);
      if (y16) HP(y)[yo] = __float2half_rn(v); else y[yo] = v;
    }
  }
}

}  // namespace

// ==================================================================================================
// C entry points
// ==================================================================================================
#define CHECK_ALIGN4(v, what) \
  if ((v) % 4) return what " must be a multiple of 4"

PBK pbk_backend_name() { return "cuda-sm100a"; }
PBK pbk_memset0(void* p, size_t bytes, pb_stream st) { return cuda_err(cudaMemsetAsync(p, 0, bytes, S(st))); }
PBK pbk_copy(void* dst, const void* src, size_t bytes, pb_stream st) {
  return cuda_err(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, S(st)));
}
PBK pbk_download(void* host_dst, const void* src, size_t bytes, pb_stream st) {
  cudaError_t e = cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, S(st));
  if (e != cudaSuccess) return cuda_err(e);
  return cuda_err(cudaStreamSynchronize(S(st)));
}
PBK pbk_upload(void* dst, const void* host_src, size_t bytes, pb_stream st) {
  cudaError_t e = cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, S(st));
  if (e != cudaSuccess) return cuda_err(e);
  return cuda_err(cudaStreamSynchronize(S(st)));
}
PBK pbk_sync(pb_stream st) { return cuda_err(cudaStreamSynchronize(S(st))); }

PBK pbk_graph_begin(pb_stream st) {
  if (S(st) == nullptr || S(st) == cudaStreamLegacy) return "graph capture needs a non-default stream";
  return cuda_err(cudaStreamBeginCapture(S(st), cudaStreamCaptureModeThreadLocal));
}
PBK pbk_graph_end(pb_stream st, void** graph_exec, long* kernel_nodes) {
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(S(st), &g);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_err(e); }
  size_t n = 0;
  if (kernel_nodes) {
    *kernel_nodes = 0;
    if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n) {
      std::vector<cudaGraphNode_t> nodes(n);
      cudaGraphGetNodes(g, nodes.data(), &n);
      for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++*kernel_nodes;
      }
    }
  }
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return cuda_err(e);
  *graph_exec = ex;
  return nullptr;
}
PBK pbk_graph_launch(void* graph_exec, pb_stream st) {
  return cuda_err(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), S(st)));
}
PBK pbk_graph_destroy(void* graph_exec) {
  return cuda_err(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)));
}

PBK pbk_event_record(void** ev, pb_stream st) {
  if (!*ev) {
    cudaEvent_t e;
    if (const char* err = cuda_err(cudaEventCreate(&e))) return err;
    *ev = e;
  }
  return cuda_err(cudaEventRecord(static_cast<cudaEvent_t>(*ev), S(st)));
}
extern "C" __attribute__((visibility("default"))) float pbk_event_elapsed_ms(void* e0, void* e1) {
  float ms = 0.f;
  if (cudaEventSynchronize(static_cast<cudaEvent_t>(e1)) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(e0), static_cast<cudaEvent_t>(e1)) != cudaSuccess) return -1.f;
  return ms;
}
PBK pbk_event_destroy(void* ev) { return ev ? cuda_err(cudaEventDestroy(static_cast<cudaEvent_t>(ev))) : nullptr; }

PBK pbk_gemm(const PbGemm* g, pb_stream st) {
  static const bool trace = getenv("PB_TRACE_GEMM") != nullptr;     // shape log for scripts/bench_gemm.py
  if (trace)
    printf("GEMM M=%d N=%d K0=%d K1=%d nseg=%d nb=%d nh=%d conv=%d H=%d W=%d res=%d\n", g->M, g->N, g->seg[0].K,
           g->nseg > 1 ? g->seg[1].K : 0, g->nseg, g->nb, g->nh, g->conv, g->H, g->W, g->R != nullptr);
  return pb_gemm_launch(*g, S(st));
}

PBK pbk_conv3x3_direct(const float* x, int nb, int H, int W, int Cin, const float* w, const float* bias, int Cout,
                       float* y, float beta, int io, pb_stream st) {
  const long pix = (long)nb * H * W;
  const bool x16 = in16(io), y16 = out16(io);
  if ((Cin == 3 || Cin == 4) && Cout % 2 == 0 && Cout >= 64) {                  // conv_in
    const int gy = (Cout / 2 + 255) / 256;
    const int bx = std::min(256, ((Cout / 2 + 31) / 32) * 32);
    const long strips = (long)nb * H * ((W + THIN_PW - 1) / THIN_PW);
    dim3 grid((unsigned)std::min<long>(strips, (long)kSMs * 8 / ((Cout / 2 + bx - 1) / bx)), (Cout / 2 + bx - 1) / bx);
    (void)gy;
    if (Cin == 4) conv3x3_thin_in_reg_k<4><<<grid, bx, 0, S(st)>>>(x, nb, H, W, w, bias, Cout, y, beta, x16, y16);
    else conv3x3_thin_in_reg_k<3><<<grid, bx, 0, S(st)>>>(x, nb, H, W, w, bias, Cout, y, beta, x16, y16);
  } else if ((Cout == 3 || Cout == 4) && Cin % 4 == 0 && (size_t)Cout * 9 * Cin * 4 <= 96 * 1024 &&
             (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {   // its transpose
    const int shmem = Cout * 9 * Cin * 4;
    const long strips = (long)nb * H * ((W + THIN_PW - 1) / THIN_PW);
    const unsigned blocks = (unsigned)std::min<long>((strips + 7) / 8, (long)kSMs * 4);
    if (Cout == 4) {
      if (const char* err = pbhost::optin_smem(conv3x3_thin_out_smem_k<4>, 96 * 1024)) return err;
      conv3x3_thin_out_smem_k<4><<<blocks, 256, shmem, S(st)>>>(x, nb, H, W, Cin, w, bias, y, beta, x16, y16);
    } else {
      if (const char* err = pbhost::optin_smem(conv3x3_thin_out_smem_k<3>, 96 * 1024)) return err;
      conv3x3_thin_out_smem_k<3><<<blocks, 256, shmem, S(st)>>>(x, nb, H, W, Cin, w, bias, y, beta, x16, y16);
    }
  } else if (Cout <= 8 && Cin >= 32) {
    conv3x3_direct_thin_out<8><<<grid_for(pix * 32, 256), 256, 0, S(st)>>>(x, nb, H, W, Cin, w, bias, Cout, y, beta, x16, y16);
  } else {
    const long total = pix * Cout;
    conv3x3_direct_thin_in<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(x, nb, H, W, Cin, w, bias, Cout, y, beta, x16, y16);
  }
  return last_err();
}
PBK pbk_im2col_s2(const float* x, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, float* col, int round_tf32,
                  pb_stream st) {
  if (in16(round_tf32) && out16(round_tf32)) {      // pure data movement: [..][C] halves are [..][C / 2] floats
    if (C % 8) return "im2col: C must be a multiple of 8 for fp16 tangents";
    C /= 2; round_tf32 = 0;
  } else if (in16(round_tf32)) return "im2col: fp16 input needs fp16 output";
  CHECK_ALIGN4(C, "im2col: C");
  const long total = (long)nb * Ho * Wo * 9 * (C / 4);
  im2col_s2_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(reinterpret_cast<const float4*>(x), nb, H, W, C / 4, pad_lo,
                                                          Ho, Wo, reinterpret_cast<float4*>(col), round_tf32);
  return last_err();
}
PBK pbk_col2im_s2(const float* col, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, float* gx, float beta,
                  int round_tf32, pb_stream st) {
  if (in16(round_tf32)) {
    if (!out16(round_tf32)) return "col2im: fp16 input needs fp16 output";
    return pb16::col2im_s2(HP(col), nb, H, W, C, pad_lo, Ho, Wo, HP(gx), beta, S(st));
  }
  CHECK_ALIGN4(C, "col2im: C");
  const long total = (long)nb * H * W * (C / 4);
  col2im_s2_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(reinterpret_cast<const float4*>(col), nb, H, W, C / 4, pad_lo,
                                                          Ho, Wo, reinterpret_cast<float4*>(gx), beta, round_tf32);
  return last_err();
}

PBK pbk_copy2d(float* dst, long ldd, const float* src, long lds, long rows, int cols, float beta, int round_tf32,
               pb_stream st) {
  if (rows <= 0 || cols <= 0) return nullptr;
  if (in16(round_tf32)) {
    if (!out16(round_tf32)) return "copy2d: fp16 input needs fp16 output";
    return pb16::copy2d(HP(dst), ldd, HP(src), lds, rows, cols, beta, S(st));
  }
  const bool v4 = (cols % 4 == 0) && (ldd % 4 == 0) && (lds % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  if (v4) copy2d_v4<<<grid_for(rows * (cols / 4), 256, 16), 256, 0, S(st)>>>(dst, ldd, src, lds, rows, cols / 4, beta, round_tf32);
  else copy2d_s<<<grid_for(rows * cols, 256, 16), 256, 0, S(st)>>>(dst, ldd, src, lds, rows, cols, beta, round_tf32);
  return last_err();
}
PBK pbk_transpose(float* dst, long ldd, long sbd, long shd, const float* src, long lds, long sbs, long shs, int nb,
                  int nh, int R, int C, float beta, int round_tf32, pb_stream st) {
  if ((long)nb * nh > 65535 || (R + 31) / 32 > 65535) return "transpose: batch / row extent too large";
  if (out16(round_tf32) && beta != 0.f) return "transpose: fp16 output cannot accumulate";
  if (in16(round_tf32) && out16(round_tf32) && C % 8 == 0 && C <= 128 && R >= 256 && lds % 8 == 0 && ldd % 8 == 0 && sbs % 8 == 0 &&
      shs % 8 == 0 && sbd % 8 == 0 && shd % 8 == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0 &&
      (R + TH_ROWS - 1) / TH_ROWS <= 65535 * 32) {              // narrow per-head matrices of the attention operands
    dim3 grid((R + TH_ROWS - 1) / TH_ROWS, nb * nh);
    transpose_heads16_k<<<grid, 256, TH_ROWS * (C + 2) * 2, S(st)>>>(HP(dst), ldd, sbd, shd, HP(src), lds, sbs, shs, nh, R, C);
    return last_err();
  }
  dim3 grid((C + 31) / 32, (R + 31) / 32, nb * nh), block(32, 8);
  transpose_k<<<grid, block, 0, S(st)>>>(dst, ldd, sbd, shd, src, lds, sbs, shs, nh, R, C, beta, round_tf32 & PB_RND_MASK,
                                         in16(round_tf32) ? 1 : 0);
  return last_err();
}
PBK pbk_upsample2x(const float* x, int nb, int H, int W, int C, float* y, int round_tf32, pb_stream st) {
  if (in16(round_tf32) && out16(round_tf32)) {      // pure data movement: [..][C] halves are [..][C / 2] floats
    if (C % 8) return "upsample: C must be a multiple of 8 for fp16 tangents";
    C /= 2; round_tf32 = 0;
  } else if (in16(round_tf32)) return "upsample: fp16 input needs fp16 output";
  CHECK_ALIGN4(C, "upsample: C");
  const long total = (long)nb * 4 * H * W * (C / 4);
  upsample2x_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(reinterpret_cast<const float4*>(x), nb, H, W, C / 4,
                                                           reinterpret_cast<float4*>(y), round_tf32);
  return last_err();
}
PBK pbk_upsample2x_vjp(const float* gy, int nb, int H, int W, int C, float* gx, float beta, int round_tf32,
                       pb_stream st) {
  if (in16(round_tf32)) {
    if (!out16(round_tf32)) return "upsample_vjp: fp16 input needs fp16 output";
    return pb16::upsample2x_vjp(HP(gy), nb, H, W, C, HP(gx), beta, S(st));
  }
  CHECK_ALIGN4(C, "upsample_vjp: C");
  const long total = (long)nb * H * W * (C / 4);
  upsample2x_vjp_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(reinterpret_cast<const float4*>(gy), nb, H, W, C / 4,
                                                               reinterpret_cast<float4*>(gx), beta, round_tf32);
  return last_err();
}
extern "C" __attribute__((visibility("default"))) int pbk_has_f16_operands() { return 1; }
PBK pbk_to_f16_scaled(void* dst, const float* src, size_t n, float scale, pb_stream st) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(dst) & 7) || (reinterpret_cast<uintptr_t>(src) & 15)) return "to_f16: n % 4 and alignment";
  to_f16_k<<<grid_for((long)(n / 4), 256, 16), 256, 0, S(st)>>>(static_cast<__half*>(dst), src, n / 4, scale);
  return last_err();
}
PBK pbk_to_f16(void* dst, const float* src, size_t n, pb_stream st) { return pbk_to_f16_scaled(dst, src, n, 1.f, st); }
PBK pbk_to_f32(float* dst, const void* src, size_t n, pb_stream st) { return pb16::to_f32(dst, static_cast<const __half*>(src), n, S(st)); }
PBK pbk_ddim_step(const float* x, const float* eps, float a_t, float a_next, float* x_next, float* pred_x0, long n,
                  pb_stream st) {
  if (!(a_t > 0.f) || !(a_next >= 0.f) || a_t > 1.f || a_next > 1.f) return "ddim_step: alphas_cumprod must lie in (0, 1]";
  const float is = 1.f / sqrtf(a_t);
  ddim_step_k<<<grid_for(n, 256, 8), 256, 0, S(st)>>>(x, eps, is, sqrtf(1.f - a_t) * is, sqrtf(a_next), sqrtf(1.f - a_next),
                                                     x_next, pred_x0, n);
  return last_err();
}
PBK pbk_lincomb3(float* out, float a, const float* x, float b, const float* y, float c, const float* z, long n, pb_stream st) {
  lincomb3_k<<<grid_for(n, 256, 8), 256, 0, S(st)>>>(out, a, x, b, y, c, z, n);
  return last_err();
}
PBK pbk_round_tf32(float* dst, const float* src, size_t n, pb_stream st) {
  round_tf32_k<<<grid_for((long)n, 256, 16), 256, 0, S(st)>>>(dst, src, n);
  return last_err();
}

// ---- GroupNorm ----
// floats of scratch pbk_gn_stats / pbk_gn_lin need: per-chunk partials + finalized group means
static int gn_chunks(int HW, int C, int nb) {
  const GnGeom g = gn_geom(C);
  int chunks = std::max(1, std::min((HW + g.R - 1) / g.R, (kSMs * 4 + nb - 1) / nb));
  const int ppb = (HW + chunks - 1) / chunks;
  return (HW + ppb - 1) / ppb;
}
extern "C" __attribute__((visibility("default"))) size_t pbk_gn_tmp_floats(int HW, int C, int G, int nb) {
  return std::max((size_t)gn_chunks(HW, C, nb) * nb * C * 2 + (size_t)nb * G * 2 + 16, pb16::gn_tmp_floats(HW, C, G, nb));
}
static const char* gn_launch_sums(int mode, const float* xp, const float* mean, const float* rstd, const float* gamma,
                                  const float* beta, int HW, int C, int G, int silu, const float* t, int nb,
                                  float* part, int* chunks_out, cudaStream_t st) {
  if (C % 4 || C % G) return "groupnorm: C must be a multiple of 4 and of the group count";
  const GnGeom g = gn_geom(C);
  if (g.QPT > GN_MAX_QPT) return "groupnorm: channel count not supported (too many quads per thread)";
  const int block = g.TP * g.R;
  const size_t shmem = (size_t)g.R * C * 2 * sizeof(float);
  const int chunks = gn_chunks(HW, C, nb);
  const int ppb = (HW + chunks - 1) / chunks;
  *chunks_out = chunks;
  dim3 grid(chunks, nb);
  if (shmem > 48 * 1024) {
    if (mode == 0) cudaFuncSetAttribute(gn_sums_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
    if (mode == 1) cudaFuncSetAttribute(gn_sums_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
    if (mode == 2) cudaFuncSetAttribute(gn_sums_k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
  }
  if (mode == 0) gn_sums_k<0><<<grid, block, shmem, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.TP, g.QPT, g.R, ppb, part);
  else if (mode == 1) gn_sums_k<1><<<grid, block, shmem, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.TP, g.QPT, g.R, ppb, part);
  else gn_sums_k<2><<<grid, block, shmem, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.TP, g.QPT, g.R, ppb, part);
  return last_err();
}

PBK pbk_gn_stats(const float* x, int nb, int HW, int C, int G, float eps, float* mean, float* rstd, float* tmp,
                 pb_stream st) {
  int chunks = 0;
  if (const char* e = gn_launch_sums(0, x, nullptr, nullptr, nullptr, nullptr, HW, C, G, 0, nullptr, nb, tmp, &chunks, S(st)))
    return e;
  gn_finalize_fwd_k<<<nb * G, 128, 0, S(st)>>>(x, tmp, chunks, nb, HW, C, G, eps, mean, rstd);
  return last_err();
}
PBK pbk_gn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int nb,
                 int HW, int C, int G, int silu, int round_tf32, float* y, pb_stream st) {
  const long total4 = (long)nb * HW * (C / 4);
  gn_apply_fwd_k<<<grid_for(total4, 256, 16), 256, 0, S(st)>>>(x, mean, rstd, gamma, beta, total4, HW, C, G, silu,
                                                              round_tf32, y);
  return last_err();
}
PBK pbk_gn_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, const float* beta, int HW,
               int C, int G, int silu, const float* t, int nb, int mode, float* out, float acc, int round_tf32,
               float* tmp, int k_slot, long p_stride, pb_stream st) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  if (in16(round_tf32)) {
    if (!out16(round_tf32)) return "groupnorm: fp16 input needs fp16 output";
    return pb16::gn_lin(xp, mean, rstd, gamma, beta, HW, C, G, silu, HP(t), nb, mode, HP(out), acc, tmp, k_slot, p_stride, S(st));
  }
  if (k_slot != nb) return "groupnorm: problem slots need the fp16-tangent kernels";
  if (round_tf32 == 2 && acc != 0.f) return "groupnorm: fp16 output cannot accumulate";
  if (C % 4 || C % G) return "groupnorm: C must be a multiple of 4 and of the group count";
  {
    // enough (image, group) pairs to fill the SMs, or a tensor small enough that launch count is all that matters:
    // one launch per GroupNorm (gn_lin_group_k); otherwise partial sums over pixel chunks + finalize + apply
    const int cpg = C / G;
    const long pairs = (long)G * nb;
    const int V = cpg % 4 == 0 ? 4 : cpg % 2 == 0 ? 2 : 1;
    if ((pairs >= 120 || (long)nb * HW * C <= (1L << 20)) && cpg / V <= 256) {
      // many pixels per group: a cluster of 2 / 4 / 8 blocks per pair (PB_GN_SPLIT overrides the choice: tuning)
      static const int env_split = getenv("PB_GN_SPLIT") ? atoi(getenv("PB_GN_SPLIT")) : 0;
      int nsplit = env_split ? (HW >= 1024 ? env_split : 1) : (HW >= 4096 ? 4 : 1);   // measured: 4096 pixels 48 -> 41 us, 1024 pixels 19 -> 25 us
      if (nsplit != 2 && nsplit != 4 && nsplit != 8) nsplit = 1;
      if (HW % nsplit) nsplit = 1;
      dim3 grid(G, nb, nsplit);
      const int block = 512;
#define PB_GN_ARGS xp, mean, rstd, gamma, beta, HW, C, G, silu, t, out, acc, round_tf32
#define PB_GN_GROUP(M_, V_)                                                                                   \
  do {                                                                                                        \
    if (nsplit == 2) gn_lin_group_cluster_k<M_, V_, 2><<<grid, block, 0, S(st)>>>(PB_GN_ARGS);               \
    else if (nsplit == 4) gn_lin_group_cluster_k<M_, V_, 4><<<grid, block, 0, S(st)>>>(PB_GN_ARGS);          \
    else if (nsplit == 8) gn_lin_group_cluster_k<M_, V_, 8><<<grid, block, 0, S(st)>>>(PB_GN_ARGS);          \
    else gn_lin_group_k<M_, V_><<<grid, block, 0, S(st)>>>(PB_GN_ARGS);                                      \
  } while (0)
      if (mode == 0) { if (V == 4) PB_GN_GROUP(0, 4); else if (V == 2) PB_GN_GROUP(0, 2); else PB_GN_GROUP(0, 1); }
      else { if (V == 4) PB_GN_GROUP(1, 4); else if (V == 2) PB_GN_GROUP(1, 2); else PB_GN_GROUP(1, 1); }
#undef PB_GN_ARGS
#undef PB_GN_GROUP
      return last_err();
    }
  }
  int chunks = 0;
  float* part = tmp;
  if (const char* e = gn_launch_sums(mode ? 2 : 1, xp, mean, rstd, gamma, beta, HW, C, G, silu, t, nb, part, &chunks, S(st)))
    return e;
  tmp = tmp + (size_t)chunks * nb * C * 2;            // [nb][G][2]
  gn_finalize_lin_k<<<nb * G, 128, 0, S(st)>>>(part, chunks, nb, HW, C, G, tmp);
  const long total4 = (long)nb * HW * (C / 4);
  if (mode == 0)
    gn_apply_lin_k<0><<<grid_for(total4, 256, 16), 256, 0, S(st)>>>(xp, mean, rstd, gamma, beta, tmp, total4, HW, C, G,
                                                                   silu, t, out, acc, round_tf32);
  else
    gn_apply_lin_k<1><<<grid_for(total4, 256, 16), 256, 0, S(st)>>>(xp, mean, rstd, gamma, beta, tmp, total4, HW, C, G,
                                                                   silu, t, out, acc, round_tf32);
  return last_err();
}

// ---- LayerNorm ----
PBK pbk_ln_fwd(const float* x, long rows, int C, const float* gamma, const float* beta, float eps, float* y,
               float* mean, float* rstd, int round_tf32, pb_stream st) {
  CHECK_ALIGN4(C, "layernorm: C");
  ln_fwd_k<<<grid_for(rows * 32, 256, 8), 256, 0, S(st)>>>(x, rows, C, gamma, beta, eps, y, mean, rstd, round_tf32);
  return last_err();
}
PBK pbk_ln_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, long rows_p, int C,
               const float* t, int nb, int mode, float* out, float acc, int round_tf32, int k_slot, long p_stride, pb_stream st) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  if (in16(round_tf32)) {
    if (!out16(round_tf32)) return "layernorm: fp16 input needs fp16 output";
    return pb16::ln_lin(xp, mean, rstd, gamma, rows_p, C, HP(t), nb, mode, HP(out), acc, k_slot, p_stride, S(st));
  }
  if (k_slot != nb) return "layernorm: problem slots need the fp16-tangent kernels";
  CHECK_ALIGN4(C, "layernorm: C");
  if (round_tf32 == 2 && acc != 0.f) return "layernorm: fp16 output cannot accumulate";
  const long rows = rows_p * nb;
  if (mode == 0) ln_lin_k<0><<<grid_for(rows * 32, 256, 8), 256, 0, S(st)>>>(xp, mean, rstd, gamma, rows_p, C, t, rows, out, acc, round_tf32);
  else ln_lin_k<1><<<grid_for(rows * 32, 256, 8), 256, 0, S(st)>>>(xp, mean, rstd, gamma, rows_p, C, t, rows, out, acc, round_tf32);
  return last_err();
}

// ---- GEGLU ----
extern "C" __attribute__((visibility("default"))) int pbk_gemm_geglu_supported() { return 1; }
extern "C" __attribute__((visibility("default"))) void pbk_struct_sizes(int* gemm_bytes, int* attn_lin_bytes) {
  if (gemm_bytes) *gemm_bytes = (int)sizeof(PbGemm);
  if (attn_lin_bytes) *attn_lin_bytes = (int)sizeof(PbAttnLin);
}
PBK pbk_interleave_rows16(void* dst, const void* src, int F, int cols, pb_stream st) {
  if (F % 32 || cols % 8 || ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15))
    return "interleave_rows16: F % 32 == 0, cols % 8 == 0 and 16-byte aligned matrices";
  const long total = 2L * F * (cols / 8);
  interleave_rows16_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), F, cols / 8);
  return last_err();
}
PBK pbk_geglu_fwd(float* h, long rows, int F, float* y, int round_tf32, int prepare, pb_stream st) {
  CHECK_ALIGN4(F, "geglu: F");
  geglu_fwd_k<<<grid_for(rows * (F / 4), 256, 16), 256, 0, S(st)>>>(h, rows, F, y, round_tf32, prepare);
  return last_err();
}
PBK pbk_geglu_jvp(const float* hp, long rows_p, const float* dh, int nb, int F, float* dy, int round_tf32,
                  int k_slot, long p_stride, pb_stream st) {
  CHECK_ALIGN4(F, "geglu: F");
  const long rows = rows_p * nb;
  if (rows * (F / 4) >= (1L << 31)) return "geglu: tensor too large for 32-bit indexing";
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  const unsigned grid = grid_for(rows * (F / 4), 256, 16);
  if (in16(round_tf32)) geglu_jvp_k<true><<<grid, 256, 0, S(st)>>>(hp, rows_p, dh, rows, F, dy, round_tf32 & PB_RND_MASK, k_slot, p_stride);
  else geglu_jvp_k<false><<<grid, 256, 0, S(st)>>>(hp, rows_p, dh, rows, F, dy, round_tf32, k_slot, p_stride);
  return last_err();
}
PBK pbk_geglu_vjp(const float* hp, long rows_p, const float* gy, int nb, int F, float* gh, int round_tf32,
                  int k_slot, long p_stride, pb_stream st) {
  CHECK_ALIGN4(F, "geglu: F");
  const long rows = rows_p * nb;
  if (rows * (F / 4) >= (1L << 31)) return "geglu: tensor too large for 32-bit indexing";
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  const unsigned grid = grid_for(rows * (F / 4), 256, 16);
  if (in16(round_tf32)) geglu_vjp_k<true><<<grid, 256, 0, S(st)>>>(hp, rows_p, gy, rows, F, gh, round_tf32 & PB_RND_MASK, k_slot, p_stride);
  else geglu_vjp_k<false><<<grid, 256, 0, S(st)>>>(hp, rows_p, gy, rows, F, gh, round_tf32, k_slot, p_stride);
  return last_err();
}

// ---- softmax ----
PBK pbk_softmax_fwd(float* Sm, long rows, int cols, long ld, int round_tf32, pb_stream st) {
  if (cols <= 512) softmax_fwd_k<32><<<grid_for(rows, 8, 16), 256, 0, S(st)>>>(Sm, rows, cols, ld, round_tf32);
  else softmax_fwd_k<256><<<(unsigned)std::min<long>(rows, kSMs * 16), 256, 0, S(st)>>>(Sm, rows, cols, ld, round_tf32);
  return last_err();
}
PBK pbk_softmax_lin(const float* P, long rows_p, float* dS, int nb, int cols, long ld, int round_tf32, int k_slot,
                    long p_stride, pb_stream st) {
  CHECK_ALIGN4(ld, "softmax_lin: ld");
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  if (p_stride % 4) return "softmax_lin: p_stride must be a multiple of 4 floats";
  const long rows = rows_p * nb;
  if (cols <= 512) softmax_lin_k<32><<<grid_for(rows, 8, 16), 256, 0, S(st)>>>(P, rows_p, dS, rows, cols, ld, round_tf32, k_slot, p_stride);
  else softmax_lin_k<256><<<(unsigned)std::min<long>(rows, kSMs * 16), 256, 0, S(st)>>>(P, rows_p, dS, rows, cols, ld, round_tf32, k_slot, p_stride);
  return last_err();
}
PBK pbk_attn_delta(const float* go, long ldg, const float* o, long ldo, int nb, int N, int H, int d, float* delta,
                   int io, int k_slot, long p_stride, pb_stream st) {
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  const long total = (long)nb * H * N;
  if (d % 4 || ldg % 4 || ldo % 4 || ((reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(o)) & 15))
    return "attn_delta: head dim and leading dimensions must be multiples of 4 with 16-byte aligned bases";
  attn_delta_k<<<grid_for(total, 256, 8), 256, 0, S(st)>>>(go, ldg, o, ldo, nb, N, H, d, delta, in16(io), k_slot, p_stride);
  return last_err();
}
PBK pbk_attn_ds(const float* P, float* dP, const float* delta, float scale, int nb, int H, int rows, int cols, long ld,
                int col_mode, int round_tf32, int k_slot, long p_stride, pb_stream st) {
  CHECK_ALIGN4(ld, "attn_ds: ld");
  if (k_slot < 1 || k_slot >= nb) { k_slot = nb; p_stride = 0; }
  if (p_stride % 4) return "attn_ds: p_stride must be a multiple of 4 floats";
  const long total = (long)nb * H * rows * (ld / 4);
  attn_ds_k<<<grid_for(total, 256, 16), 256, 0, S(st)>>>(P, dP, delta, scale, nb, H, rows, cols, ld, col_mode, round_tf32, k_slot, p_stride / 4);
  return last_err();
}

// ---- time embedding / gemv / packing ----
PBK pbk_timestep_embedding(float t, int dim, int flip_sin_to_cos, float freq_shift, float* out, pb_stream st) {
  timestep_embedding_k<<<1, 256, 0, S(st)>>>(t, dim, flip_sin_to_cos, freq_shift, out);
  return last_err();
}
PBK pbk_gemv(const float* Wm, const float* x, const float* bias, int N, int K, int silu_in, int silu_out, float* y,
             pb_stream st) {
  gemv_k<<<grid_for((long)N * 32, 256, 8), 256, 0, S(st)>>>(Wm, x, bias, N, K, silu_in, silu_out, y);
  return last_err();
}
PBK pbk_pack_conv3x3(const float* w, int Co, int Ci, float* fwd, float* bwd, int round_tf32, pb_stream st) {
  pack_conv3x3_k<<<grid_for((long)Co * Ci * 9, 256, 16), 256, 0, S(st)>>>(w, Co, Ci, fwd, bwd, round_tf32);
  return last_err();
}
