// tcgen05 / TMA TF32 GEMM for sm_100a: the contraction engine behind every conv / linear /
// attention product of the pullback hot path (primal, JVP and VJP passes).
//
// One CTA computes one 128 x BN output tile:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 4D, 128B-swizzled K-major tiles, OOB zero fill
//                 supplies conv padding, K/M/N tails and attention-head tails)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, fp32 accum in TMEM)
//   warps 2..5  : epilogue (tcgen05.ld 32x32b -> alpha/bias/residual -> 128-byte row segments to HBM)
// smem ring of STAGES x (A 16 KB + B BN*128 B) with full/empty mbarriers; MMA completion is
// signalled with tcgen05.commit.  See pb_gemm.h for the operation this implements.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <algorithm>

#include "pb_gemm.h"
#include "pb_tc.cuh"

namespace pbgemm {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = 128 bytes = one swizzle row
constexpr int NTHREADS = 192;

struct alignas(64) Params {
  CUtensorMap mapA[2];
  CUtensorMap mapB[2];
  int kblocks[2];                      // ceil(K_seg / 32)
  uint32_t tx_bytes[2];                // bytes one stage's A+B boxes deliver (boxes are clamped to the tensor)
  int a_bmul[2], a_hmul[2], b_bmul[2], b_hmul[2];
  int nseg, taps, conv_ctot;
  int M, N, nb, nh;
  int conv, H, W, bw, bh, bb, tiles_w, tiles_h;
  int raster_b, mt, nt;
  float* D; const float* R; const float* bias;
  long ldd, sDb, sDh, ldr, sRb, sRh;
  float alpha, beta;
  int round_tf32;
};

using namespace pbtc;

template <int BN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;         // + barriers + 1024B alignment slack
};

template <int BN, int STAGES, int OCC>
__global__ void __launch_bounds__(NTHREADS, OCC) gemm_tf32_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = Smem<BN, STAGES>;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----
  int n0 = blockIdx.x * BN;
  int m0 = 0, bat_b = 0, bat_h = 0;          // plain mode
  int cx0 = 0, cy0 = 0, cb0 = 0;             // conv mode tile origin
  if (p.conv) {
    int t = blockIdx.y;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    cx0 = tw * p.bw; cy0 = th * p.bh; cb0 = t * p.bb;
  } else if (p.raster_b) {
    // 1-D launch, tangent index fastest: the nb CTAs that read the same broadcast A tile (attention probabilities)
    // are adjacent in launch order, so A comes from HBM once and from L2 nb - 1 times
    int t = blockIdx.x;
    bat_b = t % p.nb; t /= p.nb;
    n0 = (t % p.nt) * BN; t /= p.nt;
    m0 = (t % p.mt) * BM;
    bat_h = t / p.mt;
  } else {
    m0 = blockIdx.y * BM;
    bat_h = blockIdx.z % p.nh;
    bat_b = blockIdx.z / p.nh;
  }
  const int total_kb = p.taps * (p.kblocks[0] + (p.nseg > 1 ? p.kblocks[1] : 0));

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int dy = p.conv ? tap / 3 - 1 : 0;
        const int dx = p.conv ? tap % 3 - 1 : 0;
        for (int s = 0; s < p.nseg; ++s) {
          for (int kb = 0; kb < p.kblocks[s]; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::STAGE_BYTES;
            uint8_t* sb = sa + S::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes[s]);
            if (p.conv) {
              tma_load_4d(sa, &p.mapA[s], &full_bar[stage], kb * BK, cx0 + dx, cy0 + dy, cb0);
              tma_load_4d(sb, &p.mapB[s], &full_bar[stage], tap * p.conv_ctot + kb * BK, n0, 0, 0);
            } else {
              tma_load_4d(sa, &p.mapA[s], &full_bar[stage], kb * BK, m0, bat_h * p.a_hmul[s],
                          bat_b * p.a_bmul[s]);
              tma_load_4d(sb, &p.mapB[s], &full_bar[stage], kb * BK, n0, bat_h * p.b_hmul[s],
                          bat_b * p.b_bmul[s]);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) |
                                 (uint32_t(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bdesc = make_smem_desc(sa + S::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          // advance 8 tf32 = 32 bytes along K inside the swizzle row: +2 in the 16-byte address field
          mma_tf32(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        tcgen05_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tcgen05_commit(tmem_full_bar);
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;                   // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;              // tile row
    bool row_ok;
    long d_off, r_off;
    if (p.conv) {
      const int w = r % p.bw;
      const int hh = (r / p.bw) % p.bh;
      const int bb = r / (p.bw * p.bh);
      const int x = cx0 + w, y = cy0 + hh, b = cb0 + bb;
      row_ok = (x < p.W) && (y < p.H) && (b < p.nb);
      const long pix = (static_cast<long>(b) * p.H + y) * p.W + x;
      d_off = pix * p.ldd;
      r_off = pix * p.ldr;
    } else {
      row_ok = (m0 + r) < p.M;
      d_off = bat_b * p.sDb + bat_h * p.sDh + static_cast<long>(m0 + r) * p.ldd;
      r_off = bat_b * p.sRb + bat_h * p.sRh + static_cast<long>(m0 + r) * p.ldr;
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const float alpha = p.alpha, beta = p.beta;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nc = n0 + c * 32;
      if (nc >= p.N) break;                   // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
      if (!row_ok) continue;
      float* dptr = p.D + d_off + nc;
      const float* rptr = p.R ? p.R + r_off + nc : nullptr;
      const bool full = (nc + 32 <= p.N);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = alpha * __uint_as_float(v[j + e]);
        if (full) {
          if (p.bias) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + nc + j);
            o[0] += bv.x; o[1] += bv.y; o[2] += bv.z; o[3] += bv.w;
          }
          if (rptr) {
            const float4 rv = *reinterpret_cast<const float4*>(rptr + j);
            o[0] += beta * rv.x; o[1] += beta * rv.y; o[2] += beta * rv.z; o[3] += beta * rv.w;
          }
          if (p.round_tf32) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t t;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(o[e]));
              o[e] = __uint_as_float(t);
            }
          }
          *reinterpret_cast<float4*>(dptr + j) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (nc + j + e < p.N) {
              float t = o[e];
              if (p.bias) t += p.bias[nc + j + e];
              if (rptr) t += beta * rptr[j + e];
              if (p.round_tf32) {
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(t));
                t = __uint_as_float(u);
              }
              dptr[j + e] = t;
            }
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  });
  return fn;
}

static int g_tmap_dtype_tf32 = 0;   // 0: FLOAT32 (MMA truncates), 1: TFLOAT32 tensor-map type

const char* encode4(CUtensorMap* m, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                           const uint32_t box[4]) {
  EncodeFn fn = get_encode();
  if (!fn) return "cuTensorMapEncodeTiled entry point not available";
  cuuint64_t gd[4]; cuuint64_t gs[3]; cuuint32_t bx[4]; cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 3; ++i) gs[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return "tensor base not 16-byte aligned";
  for (int i = 0; i < 3; ++i)
    if (gs[i] % 16 != 0 || gs[i] == 0) return "tensor stride not a positive multiple of 16 bytes";
  CUresult r = fn(m, g_tmap_dtype_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                  const_cast<float*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu box=%u,%u,%u,%u",
             int(r), (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
             (unsigned long long)gd[3], (unsigned long long)gs[0], (unsigned long long)gs[1],
             (unsigned long long)gs[2], bx[0], bx[1], bx[2], bx[3]);
    return buf;
  }
  return nullptr;
}

// plain operand [rows][K] with (h, b) batch strides; stride 0 => broadcast (extent 1)
const char* encode_plain(CUtensorMap* m, const float* base, int rows, int K, long ld, long sh, int nh, long sb,
                                int nb, int box_rows, int* hmul, int* bmul, uint32_t* bytes) {
  box_rows = std::min(box_rows, rows);
  *bytes = uint32_t(box_rows) * BK * 4;
  *hmul = (sh != 0 && nh > 1) ? 1 : 0;
  *bmul = (sb != 0 && nb > 1) ? 1 : 0;
  uint64_t dims[4] = {uint64_t(K), uint64_t(rows), uint64_t(*hmul ? nh : 1), uint64_t(*bmul ? nb : 1)};
  uint64_t st[3] = {uint64_t(ld) * 4, uint64_t(*hmul ? sh : ld) * 4, uint64_t(*bmul ? sb : ld) * 4};
  uint32_t box[4] = {uint32_t(BK), uint32_t(box_rows), 1, 1};
  return encode4(m, base, dims, st, box);
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

template <int BN, int STAGES, int OCC>
static const char* launch_t(const Params& p, dim3 grid, cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  static_assert(S::TOTAL * OCC <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  gemm_tf32_kernel<BN, STAGES, OCC><<<grid, NTHREADS, S::TOTAL, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace pbgemm

extern "C" __attribute__((visibility("default"))) void pb_gemm_set_tmap_tf32(int on) { pbgemm::g_tmap_dtype_tf32 = on; }

// tuning hooks (scripts/bench_gemm.py): force a tile width (0 = heuristic) / CTAs per SM (1 or 2)
static int g_force_bn = 0, g_occ = 2, g_use160 = 0;
extern "C" __attribute__((visibility("default"))) void pb_gemm_tune(int force_bn, int occ, int use160) {
  g_force_bn = force_bn; g_occ = occ; g_use160 = use160;
}

// Returns nullptr on success, else a static error string.  Stream-ordered, no host sync.
const char* pb_gemm_launch(const PbGemm& g, cudaStream_t st) {
  using namespace pbgemm;
  if (g.M <= 0 || g.N <= 0) return "gemm: empty problem";
  Params p;
  memset(&p, 0, sizeof p);
  p.nseg = g.nseg; p.M = g.M; p.N = g.N; p.nb = g.nb; p.nh = g.nh;
  p.conv = g.conv; p.H = g.H; p.W = g.W;
  p.D = g.D; p.R = g.R; p.bias = g.bias;
  p.ldd = g.ldd; p.sDb = g.sDb; p.sDh = g.sDh; p.ldr = g.ldr; p.sRb = g.sRb; p.sRh = g.sRh;
  p.alpha = g.alpha; p.beta = g.R ? g.beta : 0.f; p.round_tf32 = g.round_tf32;
  if ((g.ldd % 4) || (g.R && (g.ldr % 4)) || (reinterpret_cast<uintptr_t>(g.D) & 15) ||
      (g.R && (reinterpret_cast<uintptr_t>(g.R) & 15)) || (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)))
    return "gemm: D/R/bias must be 16-byte aligned with ld % 4 == 0";

  // tile width: 160 when it divides N (every SD channel count is a multiple of 160: no padded columns), 64 for narrow
  // outputs or when wider tiles cannot fill the machine, else 128
  int BN = 128;
  long mt = g.conv ? 0 : (long)((g.M + BM - 1) / BM) * g.nb * g.nh;
  if (g.conv) {
    p.bw = std::min(pow2_floor(g.W), 128);
    p.bh = std::min(pow2_floor(g.H), 128 / p.bw);
    p.bb = std::min(128 / (p.bw * p.bh), g.nb);
    p.tiles_w = (g.W + p.bw - 1) / p.bw;
    p.tiles_h = (g.H + p.bh - 1) / p.bh;
    mt = (long)p.tiles_w * p.tiles_h * ((g.nb + p.bb - 1) / p.bb);
    p.taps = 9;
  } else {
    p.taps = 1;
  }
  if (g_use160 && g.N % 160 == 0 && mt * (g.N / 160) >= 148) BN = 160;
  else if (g.N <= 64 || mt * ((g.N + 127) / 128) < 148) BN = 64;
  if (g_force_bn) BN = g_force_bn;

  int ktot = 0;
  for (int s = 0; s < g.nseg; ++s) {
    const PbGemmSeg& sg = g.seg[s];
    if (sg.K <= 0) return "gemm: empty K segment";
    p.kblocks[s] = (sg.K + BK - 1) / BK;
    ktot += p.kblocks[s];
    const char* err;
    if (g.conv) {
      if (g.nseg != 1) return "gemm: conv mode takes one segment";
      if (sg.K % BK) return "gemm: conv channels must be a multiple of 32";
      p.conv_ctot = sg.K;
      uint64_t dims[4] = {uint64_t(sg.K), uint64_t(g.W), uint64_t(g.H), uint64_t(g.nb)};
      uint64_t stb[3] = {uint64_t(sg.lda) * 4, uint64_t(sg.lda) * 4 * g.W, uint64_t(sg.lda) * 4 * g.W * g.H};
      uint32_t box[4] = {uint32_t(BK), uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bb)};
      err = encode4(&p.mapA[s], sg.A, dims, stb, box);
      if (err) return err;
      int hm, bm; uint32_t bbytes;
      err = encode_plain(&p.mapB[s], sg.B, g.N, 9 * sg.K, sg.ldb, 0, 1, 0, 1, BN, &hm, &bm, &bbytes);
      if (err) return err;
      p.tx_bytes[s] = uint32_t(p.bw * p.bh * p.bb) * BK * 4 + bbytes;
    } else {
      uint32_t abytes, bbytes;
      err = encode_plain(&p.mapA[s], sg.A, g.M, sg.K, sg.lda, sg.sAh, g.nh, sg.sAb, g.nb, BM, &p.a_hmul[s],
                         &p.a_bmul[s], &abytes);
      if (err) return err;
      err = encode_plain(&p.mapB[s], sg.B, g.N, sg.K, sg.ldb, sg.sBh, g.nh, sg.sBb, g.nb, BN, &p.b_hmul[s],
                         &p.b_bmul[s], &bbytes);
      if (err) return err;
      p.tx_bytes[s] = abytes + bbytes;
    }
  }
  ktot *= p.taps;
  dim3 grid((g.N + BN - 1) / BN, g.conv ? (unsigned)mt : (unsigned)((g.M + BM - 1) / BM),
            g.conv ? 1 : (unsigned)(g.nb * g.nh));
  if (!g.conv && g.nb > 1 && g.seg[0].sAb == 0 && (g.nseg == 1 || g.seg[1].sAb == 0)) {
    p.raster_b = 1; p.nt = (int)grid.x; p.mt = (int)grid.y;
    grid = dim3(grid.x * grid.y * grid.z, 1, 1);
  }
  const bool shallow = ktot <= 6;
  const bool occ2 = g_occ >= 2;
  if (BN == 160) {
    if (shallow) return launch_t<160, 2, 2>(p, grid, st);
    return occ2 ? launch_t<160, 3, 2>(p, grid, st) : launch_t<160, 6, 1>(p, grid, st);
  }
  if (BN == 128) {
    if (shallow) return launch_t<128, 2, 2>(p, grid, st);
    return occ2 ? launch_t<128, 3, 2>(p, grid, st) : launch_t<128, 6, 1>(p, grid, st);
  }
  if (shallow) return launch_t<64, 2, 2>(p, grid, st);
  return occ2 ? launch_t<64, 4, 2>(p, grid, st) : launch_t<64, 8, 1>(p, grid, st);
}
