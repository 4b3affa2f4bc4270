// Fused attention-linearisation kernel for sm_100a (tcgen05 + TMEM + TMA): the tangent / cotangent of
// softmax attention without ever materialising the N x N score tangent in HBM.
//
// Per (tangent b, head h) and 128-row tile, looping over 64-column steps j of the score matrix (2-deep pipeline):
//     S   = alpha1 * sum_seg A_seg[rows] . B_seg[cols_j]^T                (tcgen05.mma kind::tf32, K = head dim, -> TMEM)
//     T   = Pm[rows, cols_j] o (S - delta)                                 (CUDA cores: tcgen05.ld; the P tile arrives by TMA in
//                                                                          the swizzled A-operand layout and is overwritten in place)
//     Acc += T . C1[cols_j]                                                (tcgen05.mma, A = T in swizzled smem, -> TMEM)
//     r   += rowsum(T)
// and finally  D[rows] = alpha2 * Acc - r o O[rows] + beta * R[rows].
// With (A, B, Pm, delta, C1) chosen by the engine this is
//   JVP   : dO  = [P o dS] V - rowsum(P o dS) o O  (+ P dV from a plain GEMM via R),   dS = (dQ K^T + Q dK^T)/sqrt(d)
//   VJP-A : Qbar = [P   o (Obar V^T  - delta_row)] K / sqrt(d)
//   VJP-B : Kbar = [P^T o (V Obar^T  - delta_col)] Q / sqrt(d)             (delta = rowsum(Obar o O))
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread) + TMEM allocator, warps 2..9 compute (each TMEM lane
// quarter is shared by two warps that split the 128 columns).  S is double-buffered in TMEM so the tensor core computes
// S(j+1) while the CUDA cores turn S(j) into T(j).  See pb_kernels.h (PbAttnLin) for the exact semantics.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>

#include "pb_kernels.h"
#include "pb_tc.cuh"

namespace pbgemm {
const char* encode4(CUtensorMap* m, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3], const uint32_t box[4]);
const char* encode_plain(CUtensorMap* m, const float* base, int rows, int K, long ld, long sh, int nh, long sb, int nb, int box_rows,
                         int* hmul, int* bmul, uint32_t* bytes);
}

namespace pbattn {
using namespace pbtc;

constexpr int TM = 128;                // score-tile rows per CTA
constexpr int TN = 64;                 // score-tile columns per pipeline step
constexpr int BK = 32;                 // fp32 per 128-byte swizzle row
constexpr int NST = 2;                 // pipeline depth of every streamed operand
constexpr int A_KT = TM * BK * 4;      // [128 x 32] k-block tile: 16 KB
constexpr int B_KT = TN * BK * 4;      // [ 64 x 32] k-block tile:  8 KB
constexpr int NTHREADS = 320;

struct alignas(64) Params {
  CUtensorMap mapA[2], mapB[2], mapC, mapP;
  int a_bmul[2], a_hmul[2], b_bmul[2], b_hmul[2];
  uint32_t a_bytes, b_bytes, c_bytes, p_bytes;   // bytes per barrier phase
  int nseg, kbd, dpad, d;              // k-blocks over the head dim, accumulator width (multiple of 16)
  int Mr, Nc, nb, nh;
  float alpha1, alpha2, beta;
  const float* delta; int delta_mode;
  int want_rsum; const float* O; long ldo;
  float* D; long ldd, sDb;
  const float* R; long ldr, sRb;
  int round_tf32;
};

__host__ __device__ inline int smem_a_bytes(int nseg, int kbd) { return nseg * kbd * A_KT; }
__host__ __device__ inline int smem_b_bytes(int nseg, int kbd) { return nseg * kbd * B_KT; }
__host__ __device__ inline int smem_c_bytes(int dpad) { return (TN / BK) * dpad * BK * 4; }
constexpr int PT_BYTES = (TN / BK) * A_KT;       // P in / T out, in place: 32 KB per stage

// smem: [A resident][B x NST][C1 x NST][P/T x NST][rsum 2 x 128 floats][barriers]
__global__ void __launch_bounds__(NTHREADS, 1) attn_lin_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bbytes = smem_b_bytes(p.nseg, p.kbd), cbytes = smem_c_bytes(p.dpad);
  uint8_t* sA = smem;
  uint8_t* sB = sA + smem_a_bytes(p.nseg, p.kbd);
  uint8_t* sC = sB + NST * bbytes;
  uint8_t* sPT = sC + NST * cbytes;
  float* s_rsum = reinterpret_cast<float*>(sPT + NST * PT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rsum + 2 * TM);
  uint64_t* a_full = bars + 0;
  uint64_t* acc_full = bars + 1;
  uint64_t* b_full = bars + 2;       // each of the following: [NST]
  uint64_t* b_empty = bars + 4;
  uint64_t* c_full = bars + 6;
  uint64_t* c_empty = bars + 8;
  uint64_t* p_full = bars + 10;
  uint64_t* pt_empty = bars + 12;
  uint64_t* s_full = bars + 14;
  uint64_t* s_free = bars + 16;
  uint64_t* t_full = bars + 18;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * TM;
  // tangents fastest: the nb CTAs that share one head's P rows run back to back, so P comes from L2 for all but the first
  const int bat_b = blockIdx.y % p.nb, bat_h = blockIdx.y / p.nb;
  const int nj = (p.Nc + TN - 1) / TN;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1); mbar_init(acc_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 1);
      mbar_init(&p_full[i], 1); mbar_init(&pt_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 8);
      mbar_init(&t_full[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  const uint32_t tm_acc = tmem_base + NST * TN;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, p.a_bytes);
      for (int s = 0; s < p.nseg; ++s)
        for (int kb = 0; kb < p.kbd; ++kb)
          tma_load_4d(sA + (s * p.kbd + kb) * A_KT, &p.mapA[s], a_full, kb * BK, r0, bat_h * p.a_hmul[s], bat_b * p.a_bmul[s]);
      for (int j = 0; j < nj; ++j) {
        const int c0 = j * TN, st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&b_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&b_full[st], p.b_bytes);
        for (int s = 0; s < p.nseg; ++s)
          for (int kb = 0; kb < p.kbd; ++kb)
            tma_load_4d(sB + st * bbytes + (s * p.kbd + kb) * B_KT, &p.mapB[s], &b_full[st], kb * BK, c0, bat_h * p.b_hmul[s],
                        bat_b * p.b_bmul[s]);
        mbar_wait(&pt_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&p_full[st], p.p_bytes);
        for (int kb = 0; kb < TN / BK; ++kb)
          tma_load_4d(sPT + st * PT_BYTES + kb * A_KT, &p.mapP, &p_full[st], c0 + kb * BK, r0, bat_h, 0);
        mbar_wait(&c_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&c_full[st], p.c_bytes);
        for (int kb = 0; kb < TN / BK; ++kb)
          tma_load_4d(sC + st * cbytes + kb * p.dpad * BK * 4, &p.mapC, &c_full[st], c0 + kb * BK, 0, bat_h, 0);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);
      const uint32_t idesc_a = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(p.dpad >> 3) << 17) | (uint32_t(TM >> 4) << 24);
      auto do_acc = [&](int jj) {
        const int st = jj & 1;
        const uint32_t ph = (jj >> 1) & 1;
        mbar_wait(&c_full[st], ph);
        mbar_wait(&t_full[st], ph);
        tcgen05_fence_after();
        for (int kb = 0; kb < TN / BK; ++kb) {
          const uint64_t adesc = make_smem_desc(smem_u32(sPT + st * PT_BYTES + kb * A_KT));
          const uint64_t bdesc = make_smem_desc(smem_u32(sC + st * cbytes + kb * p.dpad * BK * 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32(tm_acc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_a, (jj | kb | k) ? 1u : 0u);
        }
        tcgen05_commit(&c_empty[st]);
        tcgen05_commit(&pt_empty[st]);
      };
      mbar_wait(a_full, 0);
      for (int j = 0; j < nj; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&b_full[st], ph);
        mbar_wait(&s_free[st], ph ^ 1);
        tcgen05_fence_after();
        bool first = true;
        for (int s = 0; s < p.nseg; ++s)
          for (int kb = 0; kb < p.kbd; ++kb) {
            const uint64_t adesc = make_smem_desc(smem_u32(sA + (s * p.kbd + kb) * A_KT));
            const uint64_t bdesc = make_smem_desc(smem_u32(sB + st * bbytes + (s * p.kbd + kb) * B_KT));
            const int nk = min(4, (p.d - kb * BK + 7) / 8);     // columns past d are TMA zero fill: skip those MMAs
            for (int k = 0; k < nk; ++k) {
              mma_tf32(tmem_base + st * TN, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_s, first ? 0u : 1u);
              first = false;
            }
          }
        tcgen05_commit(&b_empty[st]);
        tcgen05_commit(&s_full[st]);
        if (j > 0) do_acc(j - 1);
      }
      do_acc(nj - 1);
      tcgen05_commit(acc_full);
    }
  } else {
    // =========================== compute warps ===========================
    const int cw = warp - 2;                  // 0..7
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
    const int hh = cw >> 2;                   // which 32-column k-block of the 64-column step this warp owns
    const int row = q * 32 + lane;            // tile row
    const int r = r0 + row;
    const bool row_ok = r < p.Mr;
    const float* dbase = p.delta ? p.delta + ((long)bat_b * p.nh + bat_h) * (p.delta_mode == 1 ? p.Mr : p.Nc) : nullptr;
    const float drow = (dbase && p.delta_mode == 1 && row_ok) ? dbase[r] : 0.f;
    const uint32_t tm_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(hh * 32);
    float rsum = 0.f;
    for (int j = 0; j < nj; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int cbase = j * TN + hh * 32;
      float dcol[32];
      if (p.delta_mode == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) dcol[i] = (cbase + i < p.Nc) ? dbase[cbase + i] : 0.f;
      }
      uint8_t* tb = sPT + st * PT_BYTES + hh * A_KT + row * 128;
      mbar_wait(&p_full[st], ph);
      float4 pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = *reinterpret_cast<const float4*>(tb + ((i ^ (row & 7)) << 4));
      mbar_wait(&s_full[st], ph);
      tcgen05_fence_after();
      uint32_t sv[32];
      tmem_ld32(tm_row + st * TN, sv);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[st]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float pp[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dl = p.delta_mode == 2 ? dcol[i * 4 + e] : drow;
          const float t = pp[e] * (p.alpha1 * __uint_as_float(sv[i * 4 + e]) - dl);
          o[e] = rna_tf32(t);
          rsum += o[e];
        }
        *reinterpret_cast<float4*>(tb + ((i ^ (row & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[st]);
    }
    // ---- epilogue: D = alpha2 * Acc - rsum o O + beta * R ----
    s_rsum[hh * TM + row] = rsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hh == 0) {
      const float rs = p.want_rsum ? s_rsum[row] + s_rsum[TM + row] : 0.f;
      mbar_wait(acc_full, 0);
      tcgen05_fence_after();
      float* dptr = p.D + (long)bat_b * p.sDb + (long)(row_ok ? r : 0) * p.ldd + bat_h * p.d;
      const float* rptr = p.R ? p.R + (long)bat_b * p.sRb + (long)(row_ok ? r : 0) * p.ldr + bat_h * p.d : nullptr;
      const float* optr = (p.want_rsum && p.O) ? p.O + (long)(row_ok ? r : 0) * p.ldo + bat_h * p.d : nullptr;
      for (int c16 = 0; c16 < p.dpad; c16 += 16) {
        uint32_t v[16];
        tmem_ld16(tm_acc + (uint32_t(q * 32) << 16) + uint32_t(c16), v);
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 16; g += 4) {
          const int n = c16 + g;
          if (n >= p.d) break;                         // d is a multiple of 4
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = p.alpha2 * __uint_as_float(v[g + e]);
          if (optr) { const float4 ov = *reinterpret_cast<const float4*>(optr + n); o[0] -= rs * ov.x; o[1] -= rs * ov.y; o[2] -= rs * ov.z; o[3] -= rs * ov.w; }
          if (rptr) { const float4 rv = *reinterpret_cast<const float4*>(rptr + n); o[0] += p.beta * rv.x; o[1] += p.beta * rv.y; o[2] += p.beta * rv.z; o[3] += p.beta * rv.w; }
          if (p.round_tf32) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = rna_tf32(o[e]);
          }
          *reinterpret_cast<float4*>(dptr + n) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace pbattn

PBK pbk_attn_lin_supported(int d, int Mr, int Nc) {
  if (d % 4 || d < 8 || d > 64) return "attn_lin: head dim must be a multiple of 4 in [8, 64]";
  if (Mr < 1 || Nc < 1) return "attn_lin: empty problem";
  return nullptr;
}

PBK pbk_attn_lin(const PbAttnLin* ap, pb_stream st) {
  using namespace pbattn;
  const PbAttnLin& a = *ap;
  if (const char* e = pbk_attn_lin_supported(a.d, a.Mr, a.Nc)) return e;
  if (a.nseg < 1 || a.nseg > 2) return "attn_lin: 1 or 2 segments";
  if ((long)a.nb * a.nh > 65535) return "attn_lin: batch too large";
  Params p;
  memset(&p, 0, sizeof p);
  p.nseg = a.nseg; p.d = a.d; p.kbd = (a.d + BK - 1) / BK; p.dpad = (a.d + 15) / 16 * 16;
  p.Mr = a.Mr; p.Nc = a.Nc; p.nb = a.nb; p.nh = a.nh;
  p.alpha1 = a.alpha1; p.alpha2 = a.alpha2; p.beta = a.R ? a.beta : 0.f;
  p.delta = a.delta; p.delta_mode = a.delta ? a.delta_mode : 0;
  p.want_rsum = a.want_rsum; p.O = a.O; p.ldo = a.ldo;
  p.D = a.D; p.ldd = a.ldd; p.sDb = a.sDb; p.R = a.R; p.ldr = a.ldr; p.sRb = a.sRb; p.round_tf32 = a.round_tf32;
  if ((a.ldd % 4) || (a.R && a.ldr % 4) || (a.O && a.ldo % 4) || (a.ldp % 4) ||
      ((reinterpret_cast<uintptr_t>(a.D) | reinterpret_cast<uintptr_t>(a.R) | reinterpret_cast<uintptr_t>(a.O) |
        reinterpret_cast<uintptr_t>(a.Pm)) & 15))
    return "attn_lin: D/R/O/P must be 16-byte aligned with ld % 4 == 0";
  uint32_t abytes = 0, bbytes = 0;
  for (int s = 0; s < a.nseg; ++s) {
    const PbGemmSeg& sg = a.seg[s];
    uint32_t ab, bb;
    if (const char* e = pbgemm::encode_plain(&p.mapA[s], static_cast<const float*>(sg.A), a.Mr, a.d, sg.lda, sg.sAh, a.nh, sg.sAb, a.nb, TM, &p.a_hmul[s],
                                              &p.a_bmul[s], &ab)) return e;
    if (const char* e = pbgemm::encode_plain(&p.mapB[s], static_cast<const float*>(sg.B), a.Nc, a.d, sg.ldb, sg.sBh, a.nh, sg.sBb, a.nb, TN, &p.b_hmul[s],
                                              &p.b_bmul[s], &bb)) return e;
    abytes += ab * p.kbd; bbytes += bb * p.kbd;
  }
  p.a_bytes = abytes; p.b_bytes = bbytes;
  {
    // C1: [nh][d][ldc], K-major over the score columns; box = [32 columns] x [d rows]
    uint64_t dims[4] = {uint64_t(a.Nc), uint64_t(a.d), uint64_t(a.nh), 1};
    uint64_t stb[3] = {uint64_t(a.ldc) * 4, uint64_t(a.sCh) * 4, uint64_t(a.sCh) * 4 * a.nh};
    uint32_t box[4] = {uint32_t(BK), uint32_t(a.d), 1, 1};
    if (const char* e = pbgemm::encode4(&p.mapC, a.C1, dims, stb, box)) return e;
    p.c_bytes = uint32_t(TN / BK) * uint32_t(a.d) * BK * 4;
  }
  {
    // Pm: [nh][Mr][ldp]; box = [32 columns] x [128 rows]
    int hm, bm; uint32_t pb;
    if (const char* e = pbgemm::encode_plain(&p.mapP, a.Pm, a.Mr, a.Nc, a.ldp, a.sPh, a.nh, 0, 1, TM, &hm, &bm, &pb)) return e;
    p.p_bytes = pb * (TN / BK);
  }
  const int smem = smem_a_bytes(p.nseg, p.kbd) + NST * (smem_b_bytes(p.nseg, p.kbd) + smem_c_bytes(p.dpad) + PT_BYTES) + 2 * TM * 4 +
                   256 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_lin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  if (smem > 227 * 1024) return "attn_lin: shared memory budget exceeded";
  dim3 grid((a.Mr + TM - 1) / TM, a.nb * a.nh);
  attn_lin_kernel<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(st)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
