// Elementwise linearisation kernels over fp16 TANGENTS (the all-fp16 tangent plan of the engine): every tangent /
// cotangent tensor is [image][pixel-or-token][channel] halves, the cached primal quantities stay fp32, all arithmetic is fp32.
// These are HBM-bound kernels: full-row coalesced 8- / 16-byte accesses along the channel axis, several rows in flight per
// thread, grids sized from the SM count, fixed-order reductions (runs are bit-reproducible).
//
//   GroupNorm linearisation : gn16_sums_k  (per-(image, pixel-chunk, group) partial sums, every channel of a pixel row read by
//                             one block so the accesses are whole 2 C-byte rows, not 2 C/G-byte group slices)
//                             gn16_apply_k (re-reduces the few partials of its image in its prologue, then applies; the second
//                             read of the tangent comes from L2)
//   LayerNorm linearisation : ln16_k       (one warp per token, the row is held in registers: one read, one write)
//   data movement with arithmetic: copy2d16_k (accumulating channel-slice copy), col2im16_k, upsample_vjp16_k, to_f32_k
// Primal operands of image b are read at (b / k_slot) * p_stride: problem slots keep one primal cache per problem and the
// tangent batch holds k_slot columns of every problem back to back (pb_set_slots).
#include "pb_lin16.h"

#include "pb_dev.cuh"

namespace pb16 {
using namespace pbdev;

namespace {

// A block owns a pixel chunk x a channel slab of Cblk channels (the whole row unless C > 2048: then nz slabs of whole groups)
struct GnGeom16 { int nz, Cblk, CV, lanes, threads, chunks, ppb; };
GnGeom16 gn_geom16(int HW, int C, int G, int nb) {
  GnGeom16 g;
  const int cpg = C / G;
  // blocks of <= 256 threads, 4 resident per SM (<= 64 registers): these kernels are latency-bound on small tensors, so
  // occupancy and a single wave matter more than per-block efficiency (ncu r2c: 480-thread blocks at 76 registers ran ONE
  // block per SM in 2.9 waves, 23 % of the warp slots active)
  g.nz = 1;
  while ((C / g.nz) / 4 > 256 && g.nz < G) g.nz *= 2;
  if (G % g.nz || ((C / g.nz) % cpg) || (C / g.nz) / 4 > 256) { g.nz = 0; return g; }   // unsupported geometry
  g.Cblk = C / g.nz;
  g.CV = g.Cblk / 4;                                               // channel quads per pixel row of the slab
  g.lanes = std::max(1, std::min(std::min(256 / g.CV, 6144 / g.Cblk), 16));   // pixel lanes; smem = lanes * Cblk * 8 bytes <= 48 KB
  g.threads = (g.CV * g.lanes + 31) / 32 * 32;
  // chunks per (image, slab) rounded DOWN so that the grid fits ONE wave of 4 resident blocks per SM: rounded up, 25 images got
  // 24 chunks = 600 blocks for 592 slots and the 8 left-over blocks ran as a second wave (ncu r2: 102 us for a 52 us kernel); 3 blocks of <= 85 registers per SM
  // since the batched-load loops (4 pixels of operands in registers before the math) do not fit 64
  const int target = std::max(1, (kSMs * 3) / (nb * g.nz));
  const int maxchunks = (HW + g.lanes - 1) / g.lanes;
  g.chunks = std::max(1, std::min(maxchunks, target));
  g.ppb = (HW + g.chunks - 1) / g.chunks;
  g.ppb = (g.ppb + g.lanes - 1) / g.lanes * g.lanes;
  g.chunks = (HW + g.ppb - 1) / g.ppb;
  return g;
}

// MODE 0 (JVP): u = t ; MODE 1 (VJP): u = t * act'(gamma xhat + beta) * gamma.   part[b][chunk][g] = (sum u, sum xhat u)
template <int MODE>
__global__ void __launch_bounds__(256, 3) gn16_sums_k(const float* __restrict__ xp, const float* __restrict__ mean,
                                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta_, int HW, int C, int G, int silu,
                                                   const __half* __restrict__ t, int Cblk, int lanes, int ppb, int k_slot,
                                                   long p_stride, float* __restrict__ part) {
  extern __shared__ float2 sh2[];                                  // [lanes][Cblk] (sum u, sum xhat u) per channel
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int CV = Cblk / 4, cbase = blockIdx.z * Cblk;
  const bool active = tid < CV * lanes;
  const int cv = tid % CV, lane = tid / CV;
  const int c0 = cbase + cv * 4, cpg = C / G;
  const long ps = (long)(b / k_slot) * p_stride;
  if (active) {
    float mu[4], rs[4], ga[4], be[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (c0 + e) / cpg;
      mu[e] = mean[ps + g]; rs[e] = rstd[ps + g];
      ga[e] = MODE == 1 ? gamma[c0 + e] : 1.f; be[e] = MODE == 1 ? beta_[c0 + e] : 0.f;
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const int p0 = chunk * ppb, p1 = min(HW, p0 + ppb);
    const float* xq = xp + ps + (long)(p0 + lane) * C + c0;
    const __half* tq = t + ((long)b * HW + p0 + lane) * C + c0;
    const long step = (long)lanes * C;
    constexpr int U = 4;                                           // pixels per trip, all loads before the math (see gn16_apply_k)
    for (int pix = p0 + lane; pix < p1; pix += lanes * U, xq += U * step, tq += U * step) {
      float4 xv[U]; uint2 tu[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pix + u * lanes < p1) {
          xv[u] = __ldg(reinterpret_cast<const float4*>(xq + u * step));
          tu[u] = __ldg(reinterpret_cast<const uint2*>(tq + u * step));
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pix + u * lanes < p1) {
          const float2 t0 = __half22float2(*reinterpret_cast<const __half2*>(&tu[u].x)), t1 = __half22float2(*reinterpret_cast<const __half2*>(&tu[u].y));
          const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ts[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float xh = (xs[e] - mu[e]) * rs[e];
            float w = ts[e];
            if (MODE == 1) w *= silu ? ga[e] * silu_d(fmaf(ga[e], xh, be[e])) : ga[e];
            s1[e] += w; s2[e] = fmaf(xh, w, s2[e]);
          }
        }
    }
    float2* sp = sh2 + (long)lane * Cblk + cv * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) sp[e] = make_float2(s1[e], s2[e]);
  }
  __syncthreads();
  // one warp per group of the slab: lanes * cpg per-channel partials, summed in a fixed order
  const int w = tid >> 5, l = tid & 31, nw = blockDim.x >> 5;
  const int n = lanes * cpg, gb = Cblk / cpg, g0 = cbase / cpg;
  for (int g = w; g < gb; g += nw) {
    float a = 0.f, c = 0.f;
    for (int i = l; i < n; i += 32) {
      const float2 v = sh2[(long)(i / cpg) * Cblk + g * cpg + i % cpg];
      a += v.x; c += v.y;
    }
    a = warp_sum(a); c = warp_sum(c);
    if (l == 0) {
      float* q = part + (((long)b * gridDim.x + chunk) * G + g0 + g) * 2;
      q[0] = a; q[1] = c;
    }
  }
}

//   MODE 0 (JVP): out = act'(.) gamma rstd (t - m1 - xhat m2)          MODE 1 (VJP): out = rstd (t act'(.) gamma - m1 - xhat m2)
//   (m1, m2) = group means of (u, xhat u);   out = result + acc * out
template <int MODE, bool ACC>
__global__ void __launch_bounds__(256, 3) gn16_apply_k(const float* __restrict__ xp, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta_, int HW, int C, int G, int silu,
                                                    const __half* __restrict__ t, int Cblk, int lanes, int ppb, int k_slot,
                                                    long p_stride, const float* __restrict__ part, int chunks,
                                                    __half* __restrict__ out, float acc) {
  extern __shared__ float s_m[];                                   // [G][2] group means of this image, then [rows][2 G] scratch
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int cpg = C / G;
  {
    // group means of this image from the per-chunk partials: all threads load (chunks x 2 G values, independent loads in
    // flight together -- a per-value serial loop over the chunks costs a chain of L2 latencies), fixed-order combination
    const int nv = 2 * G, rows = max(1, (int)blockDim.x / nv);
    float* sc = s_m + nv;                                          // [rows][nv]
    const int i = tid % nv, r = tid / nv;
    if (r < rows) {
      float s = 0.f;
      const float* q = part + (long)b * chunks * nv + i;
      for (int k = r; k < chunks; k += rows) s += q[(long)k * nv];
      sc[r * nv + i] = s;
    }
    __syncthreads();
    if (tid < nv) {
      double s = 0.0;
      for (int rr = 0; rr < rows; ++rr) s += (double)sc[rr * nv + tid];
      s_m[tid] = (float)(s / ((double)HW * cpg));
    }
  }
  __syncthreads();
  const int CV = Cblk / 4;
  if (tid >= CV * lanes) return;
  const int cv = tid % CV, lane = tid / CV;
  const int c0 = blockIdx.z * Cblk + cv * 4;
  const long ps = (long)(b / k_slot) * p_stride;
  float mu[4], rs[4], ga[4], be[4], m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int g = (c0 + e) / cpg;
    mu[e] = mean[ps + g]; rs[e] = rstd[ps + g];
    ga[e] = gamma[c0 + e]; be[e] = beta_[c0 + e];
    m1[e] = s_m[2 * g]; m2[e] = s_m[2 * g + 1];
  }
  const int p0 = chunk * ppb, p1 = min(HW, p0 + ppb);
  const float* xq = xp + ps + (long)(p0 + lane) * C + c0;
  const long toff = ((long)b * HW + p0 + lane) * C + c0;
  const __half* tq = t + toff;
  __half* oq = out + toff;
  const long step = (long)lanes * C;
  // U pixels per trip, ALL loads first: the loop body otherwise runs load -> math -> store -> next load (the store to `out` and
  // the optional read of `out` keep the compiler from hoisting the next loads), one memory latency per pixel and thread -- ncu r2:
  // 0.15 of the HBM peak at 31 % issue utilisation
  constexpr int U = 4;
  for (int pix = p0 + lane; pix < p1; pix += lanes * U, xq += U * step, tq += U * step, oq += U * step) {
    float4 xv[U]; uint2 tu[U], pu[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pix + u * lanes < p1) {
        xv[u] = __ldg(reinterpret_cast<const float4*>(xq + u * step));
        tu[u] = __ldg(reinterpret_cast<const uint2*>(tq + u * step));
        if (ACC) pu[u] = *reinterpret_cast<const uint2*>(oq + u * step);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pix + u * lanes < p1) {
        const float2 t0 = __half22float2(*reinterpret_cast<const __half2*>(&tu[u].x)), t1 = __half22float2(*reinterpret_cast<const __half2*>(&tu[u].y));
        const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ts[4] = {t0.x, t0.y, t1.x, t1.y};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = (xs[e] - mu[e]) * rs[e];
          const float f = silu ? ga[e] * silu_d(fmaf(ga[e], xh, be[e])) : ga[e];
          o[e] = MODE == 0 ? f * rs[e] * (ts[e] - m1[e] - xh * m2[e]) : rs[e] * (ts[e] * f - m1[e] - xh * m2[e]);
        }
        if (ACC) {
          const float2 q0 = __half22float2(*reinterpret_cast<const __half2*>(&pu[u].x)), q1 = __half22float2(*reinterpret_cast<const __half2*>(&pu[u].y));
          o[0] += acc * q0.x; o[1] += acc * q0.y; o[2] += acc * q1.x; o[3] += acc * q1.y;
        }
        uint2 ou;
        *reinterpret_cast<__half2*>(&ou.x) = __floats2half2_rn(o[0], o[1]);
        *reinterpret_cast<__half2*>(&ou.y) = __floats2half2_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(oq + u * step) = ou;
      }
  }
}

// LayerNorm linearisation, one warp per token row held in registers (NV vectors of 8 channels per lane; NV = 0: any C, the
// row is read twice).  MODE 0 (JVP): out = gamma rstd (t - m1 - xhat m2); MODE 1 (VJP): g = t gamma; out = rstd (g - m1 - xhat m2)
template <int MODE, int NV>
__global__ void __launch_bounds__(256) ln16_k(const float* __restrict__ xp, const float* __restrict__ mean,
                                              const float* __restrict__ rstd, const float* __restrict__ gamma, long rows_p, int C,
                                              const __half* __restrict__ t, long rows, __half* __restrict__ out, float acc,
                                              int k_slot, long p_stride) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  // (32-bit row arithmetic: the launcher checks rows < 2^31; two 64-bit divisions per token were a third of this kernel's instructions)
  const unsigned rp32 = unsigned(rows_p);
  for (long r = warp; r < rows; r += nw) {
    const unsigned img = unsigned(r) / rp32;
    const long rp = unsigned(r) - img * rp32;
    const long ps = long(img / unsigned(k_slot)) * p_stride;
    const float* xr = xp + ps + rp * C;
    const __half* tr = t + r * C;
    __half* orow = out + r * C;
    const float m = mean[ps + rp], rs = rstd[ps + rp];
    float s1 = 0.f, s2 = 0.f;
    if constexpr (NV > 0) {
      float xh[NV][8], tv[NV][8];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 8;
        if (c < C) {
          f8_load(xr + c, xh[i]);
          h8_unpack(*reinterpret_cast<const uint4*>(tr + c), tv[i]);
          float gq[8];
          if (MODE == 1) f8_load(gamma + c, gq);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            xh[i][e] = (xh[i][e] - m) * rs;
            if (MODE == 1) tv[i][e] *= gq[e];
            s1 += tv[i][e]; s2 = fmaf(xh[i][e], tv[i][e], s2);
          }
        }
      }
      const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 8;
        if (c < C) {
          float o[8], gq[8];
          if (MODE == 0) f8_load(gamma + c, gq);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = MODE == 0 ? gq[e] * rs * (tv[i][e] - m1 - xh[i][e] * m2) : rs * (tv[i][e] - m1 - xh[i][e] * m2);
          if (acc != 0.f) {
            float pv[8];
            h8_unpack(*reinterpret_cast<const uint4*>(orow + c), pv);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += acc * pv[e];
          }
          *reinterpret_cast<uint4*>(orow + c) = h8_pack(o);
        }
      }
    } else {
      for (int c = lane * 8; c < C; c += 256) {
        float xv[8], tv[8], gq[8];
        f8_load(xr + c, xv);
        h8_unpack(*reinterpret_cast<const uint4*>(tr + c), tv);
        if (MODE == 1) f8_load(gamma + c, gq);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - m) * rs;
          const float u = MODE == 1 ? tv[e] * gq[e] : tv[e];
          s1 += u; s2 = fmaf(xh, u, s2);
        }
      }
      const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
      for (int c = lane * 8; c < C; c += 256) {
        float xv[8], tv[8], gq[8], o[8];
        f8_load(xr + c, xv);
        h8_unpack(*reinterpret_cast<const uint4*>(tr + c), tv);
        f8_load(gamma + c, gq);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - m) * rs;
          o[e] = MODE == 0 ? gq[e] * rs * (tv[e] - m1 - xh * m2) : rs * (tv[e] * gq[e] - m1 - xh * m2);
        }
        if (acc != 0.f) {
          float pv[8];
          h8_unpack(*reinterpret_cast<const uint4*>(orow + c), pv);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += acc * pv[e];
        }
        *reinterpret_cast<uint4*>(orow + c) = h8_pack(o);
      }
    }
  }
}

// LayerNorm linearisation for C = 64 NV4 exactly (C = 320: NV4 = 5, C = 640: NV4 = 10): HALF a warp per token, NV4 vectors of FOUR
// channels per lane.  The eight-channel vectors of ln16_k leave 8 of 40 lane-vectors at C = 320 to a second, 25 %-occupied pass;
// here every lane is busy and a warp keeps two tokens in flight.  Same arithmetic and summation tree per token as ln16_k up to the
// order of the partial sums.
template <int MODE, int NV4>
__global__ void __launch_bounds__(256) ln16h_k(const float* __restrict__ xp, const float* __restrict__ mean,
                                               const float* __restrict__ rstd, const float* __restrict__ gamma, long rows_p,
                                               const __half* __restrict__ t, long rows, __half* __restrict__ out, float acc,
                                               int k_slot, long p_stride) {
  constexpr int C = 64 * NV4;
  const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.f / (float)C;
  const unsigned rp32 = unsigned(rows_p);
  for (long r0 = 2 * warp; r0 < rows; r0 += 2 * nw) {
    const long r = r0 + half;
    const bool ok = r < rows;
    const unsigned img = ok ? unsigned(r) / rp32 : 0u;
    const long rp = ok ? long(unsigned(r) - img * rp32) : 0;
    const long ps = long(img / unsigned(k_slot)) * p_stride;
    const float* xr = xp + ps + rp * C;
    const __half* tr = t + (ok ? r : 0) * C;
    __half* orow = out + (ok ? r : 0) * C;
    const float m = mean[ps + rp], rs = rstd[ps + rp];
    float xh[NV4][4], tv[NV4][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = (hl + 16 * i) * 4;
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const uint2 tu = ok ? *reinterpret_cast<const uint2*>(tr + c) : make_uint2(0u, 0u);
      const float2 ta = __half22float2(*reinterpret_cast<const __half2*>(&tu.x)), tb = __half22float2(*reinterpret_cast<const __half2*>(&tu.y));
      xh[i][0] = (xv.x - m) * rs; xh[i][1] = (xv.y - m) * rs; xh[i][2] = (xv.z - m) * rs; xh[i][3] = (xv.w - m) * rs;
      tv[i][0] = ta.x; tv[i][1] = ta.y; tv[i][2] = tb.x; tv[i][3] = tb.y;
      if (MODE == 1) {
        const float4 gq = *reinterpret_cast<const float4*>(gamma + c);
        tv[i][0] *= gq.x; tv[i][1] *= gq.y; tv[i][2] *= gq.z; tv[i][3] *= gq.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { s1 += tv[i][e]; s2 = fmaf(xh[i][e], tv[i][e], s2); }
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {                       // over the 16 lanes of the token
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * invC, m2 = s2 * invC;
    if (ok) {
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int c = (hl + 16 * i) * 4;
        float o[4];
        if (MODE == 0) {
          const float4 gq = *reinterpret_cast<const float4*>(gamma + c);
          const float g4[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = g4[e] * rs * (tv[i][e] - m1 - xh[i][e] * m2);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = rs * (tv[i][e] - m1 - xh[i][e] * m2);
        }
        if (acc != 0.f) {
          const uint2 pu = *reinterpret_cast<const uint2*>(orow + c);
          const float2 pa = __half22float2(*reinterpret_cast<const __half2*>(&pu.x)), pb = __half22float2(*reinterpret_cast<const __half2*>(&pu.y));
          o[0] += acc * pa.x; o[1] += acc * pa.y; o[2] += acc * pb.x; o[3] += acc * pb.y;
        }
        uint2 ou;
        *reinterpret_cast<__half2*>(&ou.x) = __floats2half2_rn(o[0], o[1]);
        *reinterpret_cast<__half2*>(&ou.y) = __floats2half2_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(orow + c) = ou;
      }
    }
  }
}

// dst[r][c] = src[r][c] + beta * dst[r][c] over halves, 8 per access
__global__ void copy2d16_k(__half* __restrict__ dst, long ldd, const __half* __restrict__ src, long lds, long rows, int cols8,
                           float beta) {
  const long total = rows * cols8;
  const bool small = total < (1L << 31);                       // 32-bit index arithmetic where it fits (a 64-bit division per access otherwise)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r; int c;
    if (small) { const unsigned q = unsigned(i) / unsigned(cols8); r = q; c = int(unsigned(i) - q * unsigned(cols8)) * 8; }
    else { r = i / cols8; c = int(i % cols8) * 8; }
    uint4 v = *reinterpret_cast<const uint4*>(src + r * lds + c);
    uint4* d = reinterpret_cast<uint4*>(dst + r * ldd + c);
    if (beta != 0.f) {
      float a[8], o[8];
      h8_unpack(v, a); h8_unpack(*d, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += beta * o[e];
      v = h8_pack(a);
    }
    *d = v;
  }
}

__global__ void col2im16_k(const uint4* __restrict__ col, int nb, int H, int W, int C8, int pad, int Ho, int Wo,
                           uint4* __restrict__ gx, float beta) {
  const long total = (long)nb * H * W * C8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C8); t /= C8;
    const int ix = int(t % W); t /= W;
    const int iy = int(t % H); t /= H;
    const int b = int(t);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
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
        float v[8];
        h8_unpack(col[((((long)b * Ho + oy) * Wo + ox) * 9 + ky * 3 + kx) * C8 + c], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += v[e];
      }
    }
    if (beta != 0.f) {
      float o[8];
      h8_unpack(gx[i], o);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += beta * o[e];
    }
    gx[i] = h8_pack(a);
  }
}

__global__ void upsample_vjp16_k(const uint4* __restrict__ gy, int nb, int H, int W, int C8, uint4* __restrict__ gx, float beta) {
  const long total = (long)nb * H * W * C8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = int(t % C8); t /= C8;
    const int x = int(t % W); t /= W;
    const int y = int(t % H); t /= H;
    const int b = int(t);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        h8_unpack(gy[(((long)b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C8 + c], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += v[e];
      }
    if (beta != 0.f) {
      float o[8];
      h8_unpack(gx[i], o);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += beta * o[e];
    }
    gx[i] = h8_pack(a);
  }
}

__global__ void to_f32_k(float* __restrict__ dst, const __half* __restrict__ src, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    h8_unpack(reinterpret_cast<const uint4*>(src)[i], v);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

}  // namespace

size_t gn_tmp_floats(int HW, int C, int G, int nb) {
  if (C % 4 || C % G) return 0;
  const GnGeom16 g = gn_geom16(HW, C, G, nb);
  if (!g.nz) return 0;
  return (size_t)nb * g.chunks * G * 2 + 16;
}

const char* gn_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, const float* beta, int HW, int C,
                   int G, int silu, const __half* t, int nb, int mode, __half* out, float acc, float* tmp, int k_slot,
                   long p_stride, cudaStream_t st) {
  if (C % 4 || C % G) return "groupnorm: C must be a multiple of 4 and of the group count";
  if (k_slot < 1) k_slot = nb;
  const GnGeom16 g = gn_geom16(HW, C, G, nb);
  if (!g.nz) return "groupnorm: channel / group geometry not supported by the fp16-tangent kernel";
  dim3 grid(g.chunks, nb, g.nz);
  const size_t sh1 = (size_t)g.lanes * g.Cblk * sizeof(float2);
  const size_t sh2 = (size_t)G * 2 * sizeof(float) * (1 + std::max(1, g.threads / (2 * G)));
#define PB_GN16_APPLY(M_, A_) gn16_apply_k<M_, A_><<<grid, g.threads, sh2, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.Cblk, g.lanes, g.ppb, \
                                                                          k_slot, p_stride, tmp, g.chunks, out, acc)
  if (mode == 0) {
    gn16_sums_k<0><<<grid, g.threads, sh1, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.Cblk, g.lanes, g.ppb, k_slot, p_stride, tmp);
    if (acc != 0.f) PB_GN16_APPLY(0, true); else PB_GN16_APPLY(0, false);
  } else {
    gn16_sums_k<1><<<grid, g.threads, sh1, st>>>(xp, mean, rstd, gamma, beta, HW, C, G, silu, t, g.Cblk, g.lanes, g.ppb, k_slot, p_stride, tmp);
    if (acc != 0.f) PB_GN16_APPLY(1, true); else PB_GN16_APPLY(1, false);
  }
#undef PB_GN16_APPLY
  return last_err();
}

const char* ln_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, long rows_p, int C, const __half* t,
                   int nb, int mode, __half* out, float acc, int k_slot, long p_stride, cudaStream_t st) {
  if (C % 8) return "layernorm: C must be a multiple of 8 for fp16 tangents";
  if (k_slot < 1) k_slot = nb;
  const long rows = rows_p * nb;
  if (rows >= (1L << 31)) return "layernorm: too many rows for 32-bit indexing";
  static const bool no_half = getenv("PB_LN_HALFWARP") && atoi(getenv("PB_LN_HALFWARP")) == 0;     // A/B switch
  if (!no_half && (C == 320 || C == 640)) {                  // half a warp per token (see ln16h_k)
    const unsigned gridh = grid_for((rows + 1) / 2 * 32, 256, 8);
#define PB_LN16H(M_, NV4_) ln16h_k<M_, NV4_><<<gridh, 256, 0, st>>>(xp, mean, rstd, gamma, rows_p, t, rows, out, acc, k_slot, p_stride)
    if (C == 320) { if (mode == 0) PB_LN16H(0, 5); else PB_LN16H(1, 5); }
    else { if (mode == 0) PB_LN16H(0, 10); else PB_LN16H(1, 10); }
#undef PB_LN16H
    return last_err();
  }
  const unsigned grid = grid_for(rows * 32, 256, 8);
  const int nv = (C + 255) / 256;
#define PB_LN16(M_, NV_) ln16_k<M_, NV_><<<grid, 256, 0, st>>>(xp, mean, rstd, gamma, rows_p, C, t, rows, out, acc, k_slot, p_stride)
#define PB_LN16_MODE(M_)                                                                                    \
  do {                                                                                                      \
    if (nv == 1) PB_LN16(M_, 1); else if (nv == 2) PB_LN16(M_, 2); else if (nv == 3) PB_LN16(M_, 3);        \
    else if (nv <= 5) PB_LN16(M_, 5); else PB_LN16(M_, 0);                                                  \
  } while (0)
  if (mode == 0) PB_LN16_MODE(0); else PB_LN16_MODE(1);
#undef PB_LN16
#undef PB_LN16_MODE
  return last_err();
}

const char* copy2d(__half* dst, long ldd, const __half* src, long lds, long rows, int cols, float beta, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return nullptr;
  if (cols % 8 || ldd % 8 || lds % 8 || ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15))
    return "copy2d: fp16 slices must be multiples of 8 columns with 16-byte aligned rows";
  copy2d16_k<<<grid_for(rows * (cols / 8), 256, 16), 256, 0, st>>>(dst, ldd, src, lds, rows, cols / 8, beta);
  return last_err();
}

const char* col2im_s2(const __half* col, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, __half* gx, float beta,
                      cudaStream_t st) {
  if (C % 8) return "col2im: C must be a multiple of 8 for fp16 tangents";
  const long total = (long)nb * H * W * (C / 8);
  col2im16_k<<<grid_for(total, 256, 16), 256, 0, st>>>(reinterpret_cast<const uint4*>(col), nb, H, W, C / 8, pad_lo, Ho, Wo,
                                                       reinterpret_cast<uint4*>(gx), beta);
  return last_err();
}

const char* upsample2x_vjp(const __half* gy, int nb, int H, int W, int C, __half* gx, float beta, cudaStream_t st) {
  if (C % 8) return "upsample_vjp: C must be a multiple of 8 for fp16 tangents";
  const long total = (long)nb * H * W * (C / 8);
  upsample_vjp16_k<<<grid_for(total, 256, 16), 256, 0, st>>>(reinterpret_cast<const uint4*>(gy), nb, H, W, C / 8,
                                                             reinterpret_cast<uint4*>(gx), beta);
  return last_err();
}

const char* to_f32(float* dst, const __half* src, size_t n, cudaStream_t st) {
  if (n % 8 || (reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) return "to_f32: n % 8 and alignment";
  to_f32_k<<<grid_for((long)(n / 8), 256, 16), 256, 0, st>>>(dst, src, n / 8);
  return last_err();
}

}  // namespace pb16
