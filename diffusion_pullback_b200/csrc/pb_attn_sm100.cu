// Fused attention-linearisation kernel for sm_100a (tcgen05 + TMEM + TMA): the tangent / cotangent of
// softmax attention without ever materialising the N x N score tangent in HBM, streaming the probability
// matrix exactly once per (tangent, head) for every product that contracts with it.
//
// Per (tangent b, head h) and 128-row tile, looping over 32-column steps j of the score matrix:
//     S   = alpha1 * sum_seg A_seg[rows] . B_seg[cols_j]^T                (tcgen05.mma kind::tf32, K = head dim, -> TMEM ring of 4)
//     Acc2 += Pm[rows, cols_j] . C2[cols_j]                                (optional: the P tile as it arrived by TMA)
//     T   = Pm[rows, cols_j] o (S - delta)                                 (CUDA cores: tcgen05.ld; the P tile sits in the swizzled
//                                                                          A-operand layout and is overwritten in place)
//     Acc += T . C1[cols_j]                                                (tcgen05.mma, A = T in swizzled smem, -> TMEM)
//     r   += rowsum(T)
// and finally  D[rows] = alpha2 * Acc - r o O[rows] + beta * R[rows]  (and D2[rows] = Acc2; without D2, Acc2 is Acc).
// With (A, B, Pm, delta, C1, C2) chosen by the engine this is
//   JVP   : dO  = [P o dS] V - rowsum(P o dS) o O + P dV,                  dS = (dQ K^T + Q dK^T)/sqrt(d)
//   VJP-A : Qbar = [P   o (Obar V^T  - delta_row)] K / sqrt(d)
//   VJP-B : Kbar = [P^T o (V Obar^T  - delta_col)] Q / sqrt(d),  Vbar = P^T Obar        (delta = rowsum(Obar o O))
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + issuer of the score MMAs, warp 10 issuer of the accumulating
// MMAs, warps 2..9 compute in two groups of four that take alternate column steps (each warp owns one TMEM lane quarter =
// 32 tile rows, all 32 columns of its step).
// Pipeline: four independent smem rings -- S operands B (3 stages), C2 (3), C1 (4-5: a C1 tile lives until the accumulate
// MMA LAG = 2 steps later has retired) and the in-place P/T tiles (every remaining 16 KB) -- all fed by one producer
// thread that polls the rings' "empty" barriers and refills whichever stage its consumer has released, so that every
// operand of step j is requested as many steps ahead as its ring is deep; S itself sits in a TMEM ring of 4; the
// accumulate MMA trails the S MMA by two steps so that the tensor core never waits for the CUDA cores.
// Head-dim tail: a head dim of 32 m + 8 (or + 16) keeps its last k-block in a 32-byte (64-byte) swizzled tile instead of
// a zero-padded 128-byte one (SD-1.5's d = 40: A 40 KB instead of 64 KB, which is what buys the deep P ring).
// See pb_kernels.h (PbAttnLin) for the exact semantics.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pb_host_util.h"
#include "pb_kernels.h"
#include "pb_tc.cuh"

namespace pbgemm {
const char* encode4x(CUtensorMap* m, const void* base, int f16, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4], int swizzle_bytes);
const char* encode_plainx(CUtensorMap* m, const void* base, int f16, int rows, int K, long ld, long sh, int nh, long sb,
                          int nb, int box_cols, int box_rows, int swizzle_bytes, int* hmul, int* bmul, uint32_t* bytes);
}

namespace pbattn16 {   // pb_attn16_sm100.cu: the column-batched kernel of the all-fp16 plan
const char* launch(const PbAttnLin& a, cudaStream_t st, bool* handled);
}

namespace pbattn {
using namespace pbtc;

constexpr int TM = 128;                // score-tile rows per CTA
constexpr int TN = 32;                 // score-tile columns per pipeline step = one 128-byte k-block of the T . C1 product
constexpr int BK = 32;                 // fp32 per 128-byte swizzle row
constexpr int NS = 4;                  // S ring in TMEM
constexpr int LAG = 2;                 // the accumulate MMA of step j is issued with the S MMA of step j + LAG
constexpr int MAX_NB = 3;              // B ring and C2 ring: 3 stages, 2 when shared memory is tight (head dim > 64)
constexpr int MAX_NC1 = 6, MAX_NPT = 8;
constexpr int PT_BYTES = TM * BK * 4;  // P in / T out, in place: 16 KB per stage
constexpr int NTHREADS = 352;          // warp 0 TMA, 1 S issue, 2..9 compute, 10 accumulate issue
constexpr int TMEM_COLS = 512;         // S ring [0, 128), Acc [128, 224), Acc2 [224, 320)
constexpr int ACC2_OFF = 96;           // head dim <= 96

struct alignas(64) Params {
  CUtensorMap mapA[2], mapB[2];        // full 32-float k-blocks of the S operands
  CUtensorMap mapAt[2], mapBt[2];      // the compact tail k-block (8 or 16 floats), if any
  CUtensorMap mapC, mapC2, mapP;
  int a_bmul[2], a_hmul[2], b_bmul[2], b_hmul[2];
  // problem slots: tangent b belongs to problem b / k_slot; a segment operand with a_slot / b_slot set is PRIMAL and its batch
  // coordinate is the problem index, as are the batch coordinates of the P and C1 maps (ps_mul = 1) and the O pointer
  int k_slot, a_slot[2], b_slot[2], ps_mul; long o_stride;
  uint32_t a_bytes, b_bytes, c_bytes, p_bytes;   // bytes per barrier phase
  int nseg, d, dpad;                   // accumulator width dpad = d rounded up to 16
  int kfull, tail;                     // S contraction: kfull 128-byte k-blocks + a tail of `tail` 32-byte k-steps (0, 1, 2)
  int nk_last;                         // MMAs (32-byte k-steps) of the last full k-block (3 when the head dim ends inside it)
  int bk;                              // S-operand elements per 128-byte k-block (32 fp32, 64 fp16)
  int a_seg_bytes, b_seg_bytes;        // smem bytes of one segment's A tile set / B tile set
  int b_stage_bytes, c_tile_bytes;     // one ring stage of B (all segments) / one C tile
  int nb_st, nc1_st, npt_st;           // ring depths of B / C2, C1 and P/T
  int has_c2, sep_acc2;
  int Mr, Nc, nb, nh;
  float alpha1, alpha2, beta;
  const float* delta; int delta_mode;
  int want_rsum; const float* O; long ldo;
  float* D; long ldd, sDb;
  const float* R; long ldr, sRb;
  float* D2; long ldd2, sD2b;
  int round_tf32;
  float inv_pscale;                    // 1 / p_scale of the (possibly pre-scaled) probability matrix
};

// K-major swizzled smem matrix descriptor; swizzle span 128 / 64 / 32 bytes, 8-row groups 8 * span apart
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, int span) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * span) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(span == 128 ? 2 : span == 64 ? 4 : 6) << 61;
  return d;
}

// position in a ring of n stages: stage index + phase bit, advanced without division (the role loops are single-warp
// instruction streams: every integer division by a runtime ring depth costs ~100 cycles of dependent latency)
struct Ring {
  int idx; uint32_t ph;
  __device__ __forceinline__ void next(int n) { if (++idx == n) { idx = 0; ph ^= 1u; } }
  __device__ __forceinline__ void next2(int n) { idx += 2; if (idx >= n) { idx -= n; ph ^= 1u; } }   // n >= 2
};

// smem: [A resident][B ring][C1 ring][C2 ring][P/T ring][rsum 2 x 128 floats][column deltas 8 x 64 floats][barriers]
// Template parameters fix the contraction shape so that the single-warp issue loops unroll into straight-line code
// (NSEG segments, KFULL full k-blocks, NTAIL tail k-steps of 8 floats, C2M: 0 no second product, 1 folded into Acc,
// 2 separate accumulator); NSEG < 0 is the generic kernel that reads the shape from Params.
// P16: the probability matrix (pre-scaled by p_scale = Nc so that it sits in fp16's normal range), C1 and C2 are fp16 and
// the step is 64 columns wide: P / T tiles are still 128 rows x 128 bytes in the same swizzle, the two accumulating
// products run as kind::f16, and the P tile -- the largest stream of the kernel -- is half the bytes per column.
// M16: 0 all fp32 (TF32 MMAs); 1 P16 above; 2 = 1 + the S operands (A, B segments) are fp16 too (kind::f16 score MMAs over
// 64-element k-blocks) and D / D2 are written as halves: the all-fp16 tangent plan of the engine.
// NKL: MMAs of the last full k-block (4, or 3 when the head dim ends inside it: fp16 head dim 40 = 80 of 128 bytes).
template <int NSEG, int KFULL, int NTAIL, int C2M, int M16, int NKL>
__global__ void __launch_bounds__(NTHREADS, 1) attn_lin_kernel(const __grid_constant__ Params p) {
  constexpr bool GEN = NSEG < 0;
  constexpr bool P16 = M16 >= 1, S16 = M16 == 2;
  constexpr int TNc = P16 ? 64 : 32;            // score columns per step
  const int nseg = GEN ? p.nseg : NSEG;
  const int kfull = GEN ? p.kfull : KFULL;
  const int tail = GEN ? p.tail : NTAIL;        // 32-byte k-steps of the compact tail tile
  const bool has_c2 = GEN ? (p.has_c2 != 0) : (C2M != 0);
  const bool sep_acc2 = GEN ? (p.sep_acc2 != 0) : (C2M == 2);
  const int a_seg_bytes = GEN ? p.a_seg_bytes : KFULL * TM * 128 + TM * NTAIL * 32;
  const int b_seg_bytes = GEN ? p.b_seg_bytes : KFULL * TNc * 128 + TNc * NTAIL * 32;
  const int b_stage_bytes = GEN ? p.b_stage_bytes : (NSEG > 0 ? NSEG : 1) * (KFULL * TNc * 128 + TNc * NTAIL * 32);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NB = p.nb_st, NC1 = p.nc1_st, NPT = p.npt_st;
  uint8_t* sA = smem;
  uint8_t* sB = sA + nseg * a_seg_bytes;
  uint8_t* sC = sB + NB * b_stage_bytes;
  uint8_t* sC2 = sC + NC1 * p.c_tile_bytes;
  uint8_t* sPT = sC2 + (has_c2 ? NB * p.c_tile_bytes : 0);
  float* s_rsum = reinterpret_cast<float*>(sPT + NPT * PT_BYTES);
  float* s_dcol = s_rsum + 2 * TM;                       // [8 compute warps][64]: the step's column deltas (delta_mode 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dcol + 8 * 64);
  uint64_t* a_full = bars + 0;
  uint64_t* acc_full = bars + 1;
  uint64_t* b_full = bars + 2;                 // [MAX_NB]
  uint64_t* b_empty = b_full + MAX_NB;
  uint64_t* c2_full = b_empty + MAX_NB;        // [MAX_NB]
  uint64_t* c2_empty = c2_full + MAX_NB;
  uint64_t* c_full = c2_empty + MAX_NB;        // [MAX_NC1]
  uint64_t* c_empty = c_full + MAX_NC1;
  uint64_t* s_full = c_empty + MAX_NC1;        // [NS]
  uint64_t* s_free = s_full + NS;
  uint64_t* p_full = s_free + NS;              // [MAX_NPT]
  uint64_t* pt_empty = p_full + MAX_NPT;
  uint64_t* t_full = pt_empty + MAX_NPT;
  uint64_t* p_used = t_full + MAX_NPT;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(p_used + MAX_NPT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * TM;
  // tangents fastest: the nb CTAs that share one head's P rows run back to back, so P comes from L2 for all but the first
  const int bat_b = blockIdx.y % p.nb, bat_h = blockIdx.y / p.nb;
  const int bat_s = bat_b / p.k_slot;          // problem slot of this tangent (0 with one problem)
  const int nj = (p.Nc + TNc - 1) / TNc;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1); mbar_init(acc_full, 1);
    for (int i = 0; i < MAX_NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&c2_full[i], 1); mbar_init(&c2_empty[i], 1); }
    for (int i = 0; i < MAX_NC1; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
    for (int i = 0; i < MAX_NPT; ++i) { mbar_init(&p_full[i], 1); mbar_init(&pt_empty[i], 1); mbar_init(&t_full[i], 4); mbar_init(&p_used[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (p.Nc < TNc) {
    // fewer score columns than one step: the TMA box of a B tile is only Nc rows, the rows below keep whatever the previous
    // kernel left in shared memory, and 0 (the zero-filled P columns) times a NaN bit pattern is NaN -- zero the ring once
    uint4* z = reinterpret_cast<uint4*>(sB);
    const int n16 = NB * b_stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  const uint32_t tm_acc = tmem_base + NS * TNc;
  const uint32_t tm_acc2 = sep_acc2 ? tm_acc + ACC2_OFF : tm_acc;
  const int tail_span = tail * 32;                             // bytes per row of the tail tile (32 or 64)
  const int BKe = S16 ? 64 : BK;                               // S-operand elements per 128-byte k-block
  const int a_tail_off = kfull * TM * 128, b_tail_off = kfull * TNc * 128;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, p.a_bytes);
      for (int s = 0; s < nseg; ++s) {
        uint8_t* dst = sA + s * a_seg_bytes;
        const int ab = p.a_slot[s] ? bat_s : bat_b * p.a_bmul[s];
        for (int kb = 0; kb < kfull; ++kb)
          tma_load_4d(dst + kb * TM * 128, &p.mapA[s], a_full, kb * BKe, r0, bat_h * p.a_hmul[s], ab);
        if (tail) tma_load_4d(dst + a_tail_off, &p.mapAt[s], a_full, kfull * BKe, r0, bat_h * p.a_hmul[s], ab);
      }
      // Every ring is refilled as soon as ITS consumer releases a stage: the S-operand rings are released by the score
      // MMAs, the P/T and C1 rings two steps later by the accumulate MMAs, so one thread polls the three "empty" barriers
      // instead of blocking on them in a fixed order (a blocked P wait would hold back the B tiles the score warp needs).
      Ring rp{0, 0}, rb{0, 0}, rc{0, 0};
      int mp = 0, mb = 0, mc = 0;
      const int bh[2] = {bat_h * p.b_hmul[0], bat_h * p.b_hmul[1]};
      const int bb[2] = {p.b_slot[0] ? bat_s : bat_b * p.b_bmul[0], p.b_slot[1] ? bat_s : bat_b * p.b_bmul[1]};
      const int ps = bat_s * p.ps_mul;
      long long t0 = clock64();
      unsigned spins = 0;
      while (mp < nj || mb < nj || mc < nj) {
        bool any = false;
        if (mb < nj && mbar_test_wait(&b_empty[rb.idx], rb.ph ^ 1) && (!has_c2 || mbar_test_wait(&c2_empty[rb.idx], rb.ph ^ 1))) {
          const int c0 = mb * TNc, st = rb.idx;
          mbar_arrive_expect_tx(&b_full[st], p.b_bytes);
#pragma unroll
          for (int s = 0; s < nseg; ++s) {
            uint8_t* dst = sB + st * b_stage_bytes + s * b_seg_bytes;
            for (int kb = 0; kb < kfull; ++kb) tma_load_4d(dst + kb * TNc * 128, &p.mapB[s], &b_full[st], kb * BKe, c0, bh[s], bb[s]);
            if (tail) tma_load_4d(dst + b_tail_off, &p.mapBt[s], &b_full[st], kfull * BKe, c0, bh[s], bb[s]);
          }
          if (has_c2) {
            mbar_arrive_expect_tx(&c2_full[st], p.c_bytes);
            tma_load_4d(sC2 + st * p.c_tile_bytes, &p.mapC2, &c2_full[st], c0, 0, bat_h, bat_b);
          }
          rb.next(NB); ++mb; any = true;
        }
        if (mp < nj && mbar_test_wait(&pt_empty[rp.idx], rp.ph ^ 1)) {
          mbar_arrive_expect_tx(&p_full[rp.idx], p.p_bytes);
          tma_load_4d(sPT + rp.idx * PT_BYTES, &p.mapP, &p_full[rp.idx], mp * TNc, r0, bat_h, ps);
          rp.next(NPT); ++mp; any = true;
        }
        if (mc < nj && mbar_test_wait(&c_empty[rc.idx], rc.ph ^ 1)) {
          mbar_arrive_expect_tx(&c_full[rc.idx], p.c_bytes);
          tma_load_4d(sC + rc.idx * p.c_tile_bytes, &p.mapC, &c_full[rc.idx], mc * TNc, 0, bat_h, ps);
          rc.next(NC1); ++mc; any = true;
        }
        if (any) { spins = 0; }
        else if ((++spins & 0xfff) == 0 && clock64() - t0 > 8000000000LL) {     // a protocol bug must trap, never hang the GPU
          printf("pb_attn: producer timeout block (%d,%d)\n", blockIdx.x, blockIdx.y);
          __trap();
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // =========================== MMA issuers ===========================
    // warp 1 issues the score products S(j), warp 10 the products that accumulate over the steps (P . C2, T . C1): two
    // single-warp instruction streams instead of one.  The whole warp walks its loop; one elected lane issues (elect_one).
    const uint32_t idesc_s = (1u << 4) | (S16 ? 0u : (2u << 7) | (2u << 10)) | (uint32_t(TNc >> 3) << 17) | (uint32_t(TM >> 4) << 24);
    const uint32_t idesc_a = (1u << 4) | (P16 ? 0u : (2u << 7) | (2u << 10)) | (uint32_t(p.dpad >> 3) << 17) | (uint32_t(TM >> 4) << 24);
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sm_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t uA = sm_u, uB = uA + nseg * a_seg_bytes, uC = uB + NB * b_stage_bytes, uC2 = uC + NC1 * p.c_tile_bytes;
    const uint32_t uPT = uC2 + (has_c2 ? NB * p.c_tile_bytes : 0);
    // descriptors: desc(addr + off) = desc(addr) + (off >> 4) (all of smem fits the 14-bit address field)
    const uint32_t c_tile16 = p.c_tile_bytes >> 4;
    if (warp == 1) {
      const uint64_t dA = make_smem_desc(uA), dB = make_smem_desc(uB);
      const uint64_t dAt = make_desc_sw(uA + a_tail_off, tail_span), dBt = make_desc_sw(uB + b_tail_off, tail_span);
      const uint32_t a_seg16 = a_seg_bytes >> 4, b_seg16 = b_seg_bytes >> 4, b_stage16 = b_stage_bytes >> 4;
      const int nk_last = GEN ? p.nk_last : NKL;                 // columns past d are TMA zero fill: skip those MMAs
      const int ntail = tail;
      auto mma_s = [&](uint32_t d_s, uint64_t ad, uint64_t bd, uint32_t on) {
        if constexpr (S16) mma_f16(d_s, ad, bd, idesc_s, on); else mma_tf32(d_s, ad, bd, idesc_s, on);
      };
      Ring rb{0, 0}, rs{0, 0};                                   // S operands / S in TMEM
      mbar_wait(a_full, 0);
      for (int j = 0; j < nj; ++j) {
        mbar_wait(&b_full[rb.idx], rb.ph);
        mbar_wait(&s_free[rs.idx], rs.ph ^ 1);
        tcgen05_fence_after();
        if (elect_one()) {
          uint32_t on = 0;
          const uint32_t d_s = tm_u + rs.idx * TNc;
#pragma unroll
          for (int s = 0; s < nseg; ++s) {
            const uint64_t a0 = dA + uint64_t(s * a_seg16), b0 = dB + uint64_t(rb.idx * b_stage16 + s * b_seg16);
#pragma unroll
            for (int kb = 0; kb < kfull; ++kb) {
              const uint64_t adesc = a0 + uint64_t(kb * (TM * 128 >> 4)), bdesc = b0 + uint64_t(kb * (TNc * 128 >> 4));
              const int nk = (kb == kfull - 1) ? nk_last : 4;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk) { mma_s(d_s, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), on); on = 1; }
            }
            if (ntail) {
              const uint64_t adesc = dAt + uint64_t(s * a_seg16), bdesc = dBt + uint64_t(rb.idx * b_stage16 + s * b_seg16);
#pragma unroll
              for (int k = 0; k < 2; ++k)
                if (k < ntail) { mma_s(d_s, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), on); on = 1; }
            }
          }
          tcgen05_commit(&b_empty[rb.idx]);
          tcgen05_commit(&s_full[rs.idx]);
        }
        __syncwarp();
        rb.next(NB); rs.next(NS);
      }
    } else {
      const uint64_t dC = make_smem_desc(uC), dC2 = make_smem_desc(uC2), dPT = make_smem_desc(uPT);
      const uint32_t u_acc = tm_u + NS * TNc, u_acc2 = sep_acc2 ? u_acc + ACC2_OFF : u_acc;
      uint32_t acc_on = 0, acc2_on = 0;                          // 0 until the accumulator has been written once
      Ring rc2{0, 0}, rp{0, 0}, ra_pt{0, 0}, ra_c{0, 0};         // C2 / P for the C2 product / accumulate: T, C1
      auto do_acc = [&]() {                                      // Acc += T(jj) . C1(jj), jj = the accumulate rings' position
        mbar_wait(&c_full[ra_c.idx], ra_c.ph);
        mbar_wait(&t_full[ra_pt.idx], ra_pt.ph);
        tcgen05_fence_after();
        const uint64_t adesc = dPT + uint64_t(ra_pt.idx * (PT_BYTES >> 4));
        const uint64_t bdesc = dC + uint64_t(ra_c.idx * c_tile16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if constexpr (P16) mma_f16(u_acc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_a, acc_on | uint32_t(k));
            else mma_tf32(u_acc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_a, acc_on | uint32_t(k));
          }
          tcgen05_commit(&c_empty[ra_c.idx]);
          tcgen05_commit(&pt_empty[ra_pt.idx]);
        }
        __syncwarp();
        acc_on = 1;
        ra_c.next(NC1); ra_pt.next(NPT);
      };
      for (int j = 0; j < nj; ++j) {
        if (has_c2) {                                            // Acc2 += P(j) . C2(j), before the CUDA cores overwrite P(j)
          mbar_wait(&c2_full[rc2.idx], rc2.ph);
          mbar_wait(&p_full[rp.idx], rp.ph);
          tcgen05_fence_after();
          const uint64_t adesc = dPT + uint64_t(rp.idx * (PT_BYTES >> 4));
          const uint64_t bdesc = dC2 + uint64_t(rc2.idx * c_tile16);
          const uint32_t on = sep_acc2 ? acc2_on : acc_on;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (P16) mma_f16(u_acc2, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_a, on | uint32_t(k));
              else mma_tf32(u_acc2, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_a, on | uint32_t(k));
            }
            tcgen05_commit(&c2_empty[rc2.idx]);
            tcgen05_commit(&p_used[rp.idx]);
          }
          __syncwarp();
          if (sep_acc2) acc2_on = 1; else acc_on = 1;
          rp.next(NPT); rc2.next(NB);
        }
        if (j >= LAG) do_acc();
      }
      for (int jj = max(0, nj - LAG); jj < nj; ++jj) do_acc();
      if (elect_one()) tcgen05_commit(acc_full);
      __syncwarp();
    }
  } else {
    // =========================== compute warps ===========================
    const int grp = (warp - 2) >> 2;          // 0 / 1: takes the even / odd column steps
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
    const int row = q * 32 + lane;            // tile row
    const int r = r0 + row;
    const bool row_ok = r < p.Mr;
    const float* dbase = p.delta ? p.delta + ((long)bat_b * p.nh + bat_h) * (p.delta_mode == 1 ? p.Mr : p.Nc) : nullptr;
    const float drow = (dbase && p.delta_mode == 1 && row_ok) ? dbase[r] : 0.f;
    const uint32_t tm_row = tmem_base + (uint32_t(q * 32) << 16);
    float rsum = 0.f;
    Ring rp{grp, 0}, rs{grp, 0};              // this group's position in the P/T ring and the S ring (advance by 2)
    const uint32_t swz = uint32_t(row & 7);
    for (int j = grp; j < nj; j += 2) {
      const int st = rp.idx, ss = rs.idx;
      const uint32_t ph = rp.ph;
      const int cbase = j * TNc;
      uint8_t* tb = sPT + st * PT_BYTES + row * 128;
      if constexpr (!P16) {
        float dcol[32];
        if (p.delta_mode == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) dcol[i] = (cbase + i < p.Nc) ? __ldg(dbase + cbase + i) : 0.f;
        }
        mbar_wait(&p_full[st], ph);
        float4 pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] = *reinterpret_cast<const float4*>(tb + ((uint32_t(i) ^ swz) << 4));
        mbar_wait(&s_full[ss], rs.ph);
        tcgen05_fence_after();
        uint32_t sv[32];
        tmem_ld32(tm_row + ss * TNc, sv);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[ss]);
        float4 tv[8];
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
          tv[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (has_c2) mbar_wait(&p_used[st], ph);                  // the P . C2 MMA has consumed the tile
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(tb + ((uint32_t(i) ^ swz) << 4)) = tv[i];
      } else {
        // 64 columns per step: the row is 8 chunks of 8 halves (scaled probabilities), T goes back as halves
        float* sd = s_dcol + (warp - 2) * 64;
        if (p.delta_mode == 2) {                                 // one coalesced load per warp, read back as broadcasts
          const int c0 = cbase + lane, c1 = c0 + 32;
          sd[lane] = c0 < p.Nc ? __ldg(dbase + c0) : 0.f;
          sd[32 + lane] = c1 < p.Nc ? __ldg(dbase + c1) : 0.f;
          __syncwarp();
        }
        mbar_wait(&p_full[st], ph);
        uint4 pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pv[i] = *reinterpret_cast<const uint4*>(tb + ((uint32_t(i) ^ swz) << 4));
        mbar_wait(&s_full[ss], rs.ph);
        tcgen05_fence_after();
        uint4 tv[8];
#pragma unroll
        for (int hv = 0; hv < 2; ++hv) {
          uint32_t sv[32];
          tmem_ld32(tm_row + ss * TNc + hv * 32, sv);
          if (hv == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[ss]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ci = hv * 4 + i;
            const __half2* ph2 = reinterpret_cast<const __half2*>(&pv[ci]);
            __half2 oh[4];
            float dl[8];
            if (p.delta_mode == 2) {
              const float4 da = *reinterpret_cast<const float4*>(sd + ci * 8), db = *reinterpret_cast<const float4*>(sd + ci * 8 + 4);
              dl[0] = da.x; dl[1] = da.y; dl[2] = da.z; dl[3] = da.w; dl[4] = db.x; dl[5] = db.y; dl[6] = db.z; dl[7] = db.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) dl[e] = drow;
            }
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const float2 pf = __half22float2(ph2[e2]);
              const float d0 = dl[e2 * 2], d1 = dl[e2 * 2 + 1];
              float t0 = pf.x * (p.alpha1 * __uint_as_float(sv[i * 8 + e2 * 2]) - d0);
              float t1 = pf.y * (p.alpha1 * __uint_as_float(sv[i * 8 + e2 * 2 + 1]) - d1);
              t0 = fminf(fmaxf(t0, -65504.f), 65504.f);         // saturate instead of producing inf
              t1 = fminf(fmaxf(t1, -65504.f), 65504.f);
              oh[e2] = __floats2half2_rn(t0, t1);
              const float2 back = __half22float2(oh[e2]);        // the row sum of what the tensor core will see
              rsum += back.x + back.y;
            }
            tv[ci] = *reinterpret_cast<const uint4*>(oh);
          }
        }
        if (has_c2) mbar_wait(&p_used[st], ph);                  // the P . C2 MMA has consumed the tile
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(tb + ((uint32_t(i) ^ swz) << 4)) = tv[i];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[st]);
      rp.next2(NPT); rs.next2(NS);
    }
    // ---- epilogue: group 0: D = alpha2 * Acc - rsum o O + beta * R ; group 1: D2 = Acc2 ----
    s_rsum[grp * TM + row] = rsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0 || sep_acc2) {
      const float rs = (grp == 0 && p.want_rsum) ? (s_rsum[row] + s_rsum[TM + row]) * p.inv_pscale : 0.f;
      mbar_wait(acc_full, 0);
      tcgen05_fence_after();
      const long rr = row_ok ? r : 0;
      float* dptr; const float* rptr = nullptr; const float* optr = nullptr;
      float alpha;
      uint32_t tm;
      long doff;                                     // element offset of this row's head slice in D / D2
      if (grp == 0) {
        doff = (long)bat_b * p.sDb + rr * p.ldd + bat_h * p.d;
        dptr = p.D + doff;
        if (p.R) rptr = p.R + (long)bat_b * p.sRb + rr * p.ldr + bat_h * p.d;
        if (p.want_rsum && p.O) optr = p.O + bat_s * p.o_stride + rr * p.ldo + bat_h * p.d;
        alpha = p.alpha2 * p.inv_pscale; tm = tm_acc;
      } else {
        doff = (long)bat_b * p.sD2b + rr * p.ldd2 + bat_h * p.d;
        dptr = p.D2 + doff;
        alpha = p.inv_pscale; tm = tm_acc2;
      }
      __half* hptr = reinterpret_cast<__half*>(grp == 0 ? p.D : p.D2) + doff;      // S16: the outputs hold halves
      for (int c16 = 0; c16 < p.dpad; c16 += 16) {
        uint32_t v[16];
        tmem_ld16(tm + (uint32_t(q * 32) << 16) + uint32_t(c16), v);
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 16; g += 4) {
          const int n = c16 + g;
          if (n >= p.d) break;                         // d is a multiple of 4
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = alpha * __uint_as_float(v[g + e]);
          if (optr) { const float4 ov = *reinterpret_cast<const float4*>(optr + n); o[0] -= rs * ov.x; o[1] -= rs * ov.y; o[2] -= rs * ov.z; o[3] -= rs * ov.w; }
          if (rptr) { const float4 rv = *reinterpret_cast<const float4*>(rptr + n); o[0] += p.beta * rv.x; o[1] += p.beta * rv.y; o[2] += p.beta * rv.z; o[3] += p.beta * rv.w; }
          if (p.round_tf32 && !S16) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = rna_tf32(o[e]);
          }
          if constexpr (S16) {
            uint2 hv;
            *reinterpret_cast<__half2*>(&hv.x) = __floats2half2_rn(o[0], o[1]);
            *reinterpret_cast<__half2*>(&hv.y) = __floats2half2_rn(o[2], o[3]);
            *reinterpret_cast<uint2*>(hptr + n) = hv;
          } else {
            *reinterpret_cast<float4*>(dptr + n) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace pbattn

PBK pbk_attn_lin_supported(int d, int Mr, int Nc) {
  if (d % 4 || d < 8 || d > 96) return "attn_lin: head dim must be a multiple of 4 in [8, 96]";
  if (Mr < 1 || Nc < 1) return "attn_lin: empty problem";
  return nullptr;
}

static const char* pbk_attn_lin_v1(const PbAttnLin* ap, pb_stream st);
PBK pbk_attn_lin(const PbAttnLin* ap, pb_stream st) {
  const PbAttnLin& a = *ap;
  if (const char* e = pbk_attn_lin_supported(a.d, a.Mr, a.Nc)) return e;
  if (a.nseg < 1 || a.nseg > 2) return "attn_lin: 1 or 2 segments";
  if ((long)a.nb * a.nh > 65535) return "attn_lin: batch too large";
  if (a.D2 && !a.C2) return "attn_lin: D2 needs C2";
  {
    // all-fp16 plan, head dim <= 64, the engine's four roles: the column-batched kernel (P and the primal tiles loaded once per
    // group of tangent columns); everything else: the per-column kernel below
    bool handled = false;
    if (const char* e = pbattn16::launch(a, static_cast<cudaStream_t>(st), &handled)) return e;
    static const bool cmp = getenv("PB_ATTN_CMP") != nullptr;      // debugging: run the per-column kernel too and compare D / D2
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (handled && cmp) cudaStreamIsCapturing(static_cast<cudaStream_t>(st), &cs);
    if (handled && cmp && cs == cudaStreamCaptureStatusNone) {
      static bool inside = false;
      if (!inside) {
        inside = true;
        cudaStream_t s_ = static_cast<cudaStream_t>(st);
        cudaStreamSynchronize(s_);
        const size_t n1 = (size_t)a.nb * a.sDb, n2 = a.D2 ? (size_t)a.nb * a.sD2b : 0;
        std::vector<__half> v2(n1), v2b(n2), v1(n1), v1b(n2);
        cudaMemcpy(v2.data(), a.D, n1 * 2, cudaMemcpyDeviceToHost);
        if (n2) cudaMemcpy(v2b.data(), a.D2, n2 * 2, cudaMemcpyDeviceToHost);
        setenv("PB_ATTN_V1_ONCE", "1", 1);
        PbAttnLin a1 = a;
        const char* e1 = pbk_attn_lin_v1(&a1, st);
        cudaStreamSynchronize(s_);
        cudaMemcpy(v1.data(), a.D, n1 * 2, cudaMemcpyDeviceToHost);
        if (n2) cudaMemcpy(v1b.data(), a.D2, n2 * 2, cudaMemcpyDeviceToHost);
        auto diff = [&](const std::vector<__half>& x, const std::vector<__half>& y, long ld, long sb, const char* nm) {
          double num = 0, den = 0; long bad = 0, firstbad = -1;
          for (int b = 0; b < a.nb; ++b)
            for (long r = 0; r < a.Mr; ++r)
              for (int c = 0; c < a.nh * a.d; ++c) {
                const long i = b * sb + r * ld + c;
                const float p_ = __half2float(x[i]), q_ = __half2float(y[i]);
                if (!(p_ == p_) || fabsf(p_ - q_) > 1e-2f * (fabsf(q_) + 1e-3f)) { if (firstbad < 0) firstbad = i; ++bad; }
                num += (double)(p_ - q_) * (p_ - q_); den += (double)q_ * q_;
              }
          fprintf(stderr, "pb_attn cmp %s: Mr %d Nc %d d %d nb %d nh %d nseg %d c2 %d dm %d rsum %d k_slot %d | rel %.3e bad %ld first %ld (b %ld r %ld c %ld) %s\n", nm,
                  a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg, a.C2 ? (a.D2 ? 2 : 1) : 0, a.delta ? a.delta_mode : 0, a.want_rsum, a.k_slot,
                  den > 0 ? sqrt(num / den) : -1.0, bad, firstbad, firstbad < 0 ? -1 : firstbad / sb, firstbad < 0 ? -1 : (firstbad % sb) / ld,
                  firstbad < 0 ? -1 : (firstbad % sb) % ld, e1 ? e1 : "");
        };
        diff(v2, v1, a.ldd, a.sDb, "D ");
        if (n2) diff(v2b, v1b, a.ldd2, a.sD2b, "D2");
        inside = false;
        return nullptr;
      }
    }
    static const bool sync_after = getenv("PB_ATTN_SYNC") != nullptr;
    if (handled && sync_after) cudaStreamSynchronize(static_cast<cudaStream_t>(st));
    if (handled) return nullptr;
  }
  return pbk_attn_lin_v1(ap, st);
}

// the per-column kernel of this file
static const char* pbk_attn_lin_v1(const PbAttnLin* ap, pb_stream st) {
  using namespace pbattn;
  const PbAttnLin& a = *ap;
  const int s16 = a.s16 ? 1 : 0;                // the S operands (segments) and the outputs D / D2 hold halves too
  if (s16 && !a.p16) return "attn_lin: fp16 S operands need the fp16 probability path";
  if (s16 && a.R) return "attn_lin: no residual with fp16 outputs";
  const int p16 = a.p16 ? 1 : 0;                // Pm (pre-scaled by p_scale), C1, C2 hold halves; 64-column steps
  const int TNh = p16 ? 64 : TN;                // score columns per step
  const long ces = p16 ? 2 : 4;                 // element size of Pm / C1 / C2
  const int cq = p16 ? 8 : 4;                   // elements per 16 bytes
  if (p16 && !(a.p_scale > 0.f)) return "attn_lin: p_scale must be positive";
  Params p;
  memset(&p, 0, sizeof p);
  p.nseg = a.nseg; p.d = a.d; p.dpad = (a.d + 15) / 16 * 16;
  // S contraction over the head dim in 32-byte k-steps (8 fp32 / 16 halves = one MMA): full 128-byte k-blocks of 4 steps + a
  // compact tail of 1 or 2 steps (32- / 64-byte swizzle); a 3-step remainder is a full block whose 4th MMA is skipped
  const int ses = s16 ? 2 : 4;                  // S-operand element size
  const int ksteps = (a.d * ses + 31) / 32, rem = ksteps % 4;
  p.bk = 128 / ses;
  p.kfull = ksteps / 4 + (rem == 3 ? 1 : 0);
  p.tail = rem == 3 ? 0 : rem;
  p.nk_last = rem == 3 ? 3 : 4;
  if (s16 && (a.d % 8)) return "attn_lin: fp16 S operands need a head dim that is a multiple of 8";
  p.Mr = a.Mr; p.Nc = a.Nc; p.nb = a.nb; p.nh = a.nh;
  p.alpha1 = a.alpha1; p.alpha2 = a.alpha2; p.beta = a.R ? a.beta : 0.f;
  p.delta = a.delta; p.delta_mode = a.delta ? a.delta_mode : 0;
  p.want_rsum = a.want_rsum; p.O = a.O; p.ldo = a.ldo;
  p.D = a.D; p.ldd = a.ldd; p.sDb = a.sDb; p.R = a.R; p.ldr = a.ldr; p.sRb = a.sRb; p.round_tf32 = a.round_tf32;
  p.has_c2 = a.C2 ? 1 : 0; p.sep_acc2 = a.D2 ? 1 : 0; p.D2 = a.D2; p.ldd2 = a.ldd2; p.sD2b = a.sD2b;
  p.inv_pscale = p16 ? 1.f / a.p_scale : 1.f;
  // problem slots: nslots problems of k_slot tangents each; primal operands are strided by p_stride bytes per problem
  const int k_slot = (a.k_slot > 0 && a.k_slot < a.nb) ? a.k_slot : a.nb;
  const int nslots = a.nb / k_slot;
  if (a.nb % k_slot) return "attn_lin: the tangent batch must be a whole number of problem slots";
  if (nslots > 1 && (a.p_stride <= 0 || a.p_stride % 16)) return "attn_lin: p_stride must be a positive multiple of 16 bytes";
  p.k_slot = k_slot; p.ps_mul = nslots > 1 ? 1 : 0; p.o_stride = nslots > 1 ? a.p_stride / 4 : 0;
  const int dq = s16 ? 8 : 4;                   // elements per 16 bytes of D / D2
  if ((a.ldd % dq) || (a.R && a.ldr % 4) || (a.O && a.ldo % 4) || (a.ldp % cq) || (a.ldc % cq) || (a.D2 && a.ldd2 % dq) ||
      ((reinterpret_cast<uintptr_t>(a.D) | reinterpret_cast<uintptr_t>(a.R) | reinterpret_cast<uintptr_t>(a.O) |
        reinterpret_cast<uintptr_t>(a.Pm) | reinterpret_cast<uintptr_t>(a.D2)) & 15))
    return "attn_lin: D/D2/R/O/P must be 16-byte aligned with rows that are multiples of 16 bytes";
  const int tail_span = p.tail * 32;
  const int tail_el = tail_span / ses;          // elements per row of the tail tile
  p.a_seg_bytes = p.kfull * TM * 128 + TM * tail_span;
  p.b_seg_bytes = p.kfull * TNh * 128 + TNh * tail_span;
  p.b_stage_bytes = p.nseg * p.b_seg_bytes;
  p.c_tile_bytes = p.dpad * 128;
  uint32_t abytes = 0, bbytes = 0;
  for (int s = 0; s < a.nseg; ++s) {
    PbGemmSeg sg = a.seg[s];
    uint32_t ab = 0, bb = 0;
    // a primal (batch-broadcast) operand with several problem slots: its batch dimension is the problem index
    int nbA = a.nb, nbB = a.nb;
    if (nslots > 1 && sg.sAb == 0) { sg.sAb = a.p_stride / ses; nbA = nslots; p.a_slot[s] = 1; }
    if (nslots > 1 && sg.sBb == 0) { sg.sBb = a.p_stride / ses; nbB = nslots; p.b_slot[s] = 1; }
    if (p.kfull) {
      if (const char* e = pbgemm::encode_plainx(&p.mapA[s], sg.A, s16, a.Mr, a.d, sg.lda, sg.sAh, a.nh, sg.sAb, nbA, p.bk, TM, 128,
                                                &p.a_hmul[s], &p.a_bmul[s], &ab)) return e;
      if (const char* e = pbgemm::encode_plainx(&p.mapB[s], sg.B, s16, a.Nc, a.d, sg.ldb, sg.sBh, a.nh, sg.sBb, nbB, p.bk, TNh, 128,
                                                &p.b_hmul[s], &p.b_bmul[s], &bb)) return e;
      abytes += ab * p.kfull; bbytes += bb * p.kfull;
    }
    if (p.tail) {
      if (const char* e = pbgemm::encode_plainx(&p.mapAt[s], sg.A, s16, a.Mr, a.d, sg.lda, sg.sAh, a.nh, sg.sAb, nbA, tail_el, TM,
                                                tail_span, &p.a_hmul[s], &p.a_bmul[s], &ab)) return e;
      if (const char* e = pbgemm::encode_plainx(&p.mapBt[s], sg.B, s16, a.Nc, a.d, sg.ldb, sg.sBh, a.nh, sg.sBb, nbB, tail_el, TNh,
                                                tail_span, &p.b_hmul[s], &p.b_bmul[s], &bb)) return e;
      abytes += ab; bbytes += bb;
    }
  }
  p.a_bytes = abytes; p.b_bytes = bbytes;
  {
    // C1: [nh][d][ldc], K-major over the score columns; box = [one step of columns = 128 bytes] x [d rows]
    uint64_t dims[4] = {uint64_t(a.Nc), uint64_t(a.d), uint64_t(a.nh), uint64_t(nslots)};
    uint64_t stb[3] = {uint64_t(a.ldc) * ces, uint64_t(a.sCh) * ces, nslots > 1 ? uint64_t(a.p_stride) : uint64_t(a.sCh) * ces * a.nh};
    uint32_t box[4] = {uint32_t(TNh), uint32_t(a.d), 1, 1};
    if (const char* e = pbgemm::encode4x(&p.mapC, a.C1, p16, dims, stb, box, 128)) return e;
    p.c_bytes = uint32_t(a.d) * 128;
  }
  if (a.C2) {
    // C2: [nb][nh][d][ldc2]
    if ((a.ldc2 % cq) || (reinterpret_cast<uintptr_t>(a.C2) & 15)) return "attn_lin: C2 must be 16-byte aligned with rows that are multiples of 16 bytes";
    uint64_t dims[4] = {uint64_t(a.Nc), uint64_t(a.d), uint64_t(a.nh), uint64_t(a.nb)};
    uint64_t stb[3] = {uint64_t(a.ldc2) * ces, uint64_t(a.nh > 1 ? a.sC2h : a.ldc2) * ces, uint64_t(a.nb > 1 ? a.sC2b : a.ldc2) * ces};
    uint32_t box[4] = {uint32_t(TNh), uint32_t(a.d), 1, 1};
    if (const char* e = pbgemm::encode4x(&p.mapC2, a.C2, p16, dims, stb, box, 128)) return e;
  }
  {
    // Pm: [nh][Mr][ldp]; box = [one step of columns = 128 bytes] x [128 rows]
    int hm, bm; uint32_t pb;
    if (const char* e = pbgemm::encode_plainx(&p.mapP, a.Pm, p16, a.Mr, a.Nc, a.ldp, a.sPh, a.nh, nslots > 1 ? a.p_stride / ces : 0, nslots,
                                              TNh, TM, 128, &hm, &bm, &pb)) return e;
    p.p_bytes = pb;
  }
  // ring depths from the shared-memory budget: B and C2 rings of 3 stages, C1 ring of 5 (4 when tight), every remaining
  // 16 KB goes to the in-place P/T ring; big head dims (80: SD-1.x 32x32 layers) fall back to 2 / 3 stages
  const int budget = 227 * 1024 - 1024 - (2 * TM * 4 + 8 * 64 * 4 + 1024);
  int nbst = MAX_NB, nc1 = 5, npt = 0;
  auto fit = [&]() {
    const int fixed = p.nseg * p.a_seg_bytes + nbst * p.b_stage_bytes + (p.has_c2 ? nbst * p.c_tile_bytes : 0);
    npt = std::min((budget - fixed - nc1 * p.c_tile_bytes) / PT_BYTES, MAX_NPT);
    return fixed + nc1 * p.c_tile_bytes;
  };
  int used = fit();
  if (npt < 5) { nc1 = 4; used = fit(); }
  if (npt < LAG + 1) { nbst = 2; nc1 = LAG + 1; used = fit(); }
  if (npt < LAG + 1) return "attn_lin: shared memory budget exceeded";
  p.nb_st = nbst;
  p.nc1_st = nc1; p.npt_st = npt;
  const int smem = used + npt * PT_BYTES + 2 * TM * 4 + 8 * 64 * 4 + 1024 + 1024;
  if (smem > 227 * 1024) return "attn_lin: shared memory budget exceeded";
  dim3 grid((a.Mr + TM - 1) / TM, a.nb * a.nh);
  const int c2m = a.C2 ? (a.D2 ? 2 : 1) : 0;
  void (*kern)(Params) = s16 ? attn_lin_kernel<-1, -1, -1, -1, 2, 4> : p16 ? attn_lin_kernel<-1, -1, -1, -1, 1, 4> : attn_lin_kernel<-1, -1, -1, -1, 0, 4>;
  // shape-specialised instantiations (straight-line issue loops): (segments, full k-blocks, tail k-steps, C2 mode, MMAs of the last block)
#define PB_ATTN_CASE(NSEG_, KF_, NT_, C2M_)                                                                        \
  if (!p16 && p.nk_last == 4 && p.nseg == NSEG_ && p.kfull == KF_ && p.tail == NT_ && c2m == C2M_)                 \
    kern = attn_lin_kernel<NSEG_, KF_, NT_, C2M_, 0, 4>;
  PB_ATTN_CASE(2, 1, 1, 1) PB_ATTN_CASE(1, 1, 1, 0) PB_ATTN_CASE(1, 1, 1, 2)      // fp32 head dim 40 (SD-1.x 64x64 layers): JVP, VJP-A, VJP-B
  PB_ATTN_CASE(2, 2, 0, 1) PB_ATTN_CASE(1, 2, 0, 0) PB_ATTN_CASE(1, 2, 0, 2)      // fp32 head dim 64 (SD-2.x)
  PB_ATTN_CASE(2, 2, 2, 1) PB_ATTN_CASE(1, 2, 2, 0) PB_ATTN_CASE(1, 2, 2, 2)      // fp32 head dim 80 (SD-1.x 32x32 layers)
#undef PB_ATTN_CASE
#define PB_ATTN_CASE16(NSEG_, KF_, NT_, C2M_, NKL_)                                                                \
  if (s16 && p.nk_last == NKL_ && p.nseg == NSEG_ && p.kfull == KF_ && p.tail == NT_ && c2m == C2M_)               \
    kern = attn_lin_kernel<NSEG_, KF_, NT_, C2M_, 2, NKL_>;
  PB_ATTN_CASE16(2, 1, 0, 1, 3) PB_ATTN_CASE16(1, 1, 0, 0, 3) PB_ATTN_CASE16(1, 1, 0, 2, 3)   // fp16 head dim 40: 80 of 128 bytes
  PB_ATTN_CASE16(2, 1, 0, 1, 4) PB_ATTN_CASE16(1, 1, 0, 0, 4) PB_ATTN_CASE16(1, 1, 0, 2, 4)   // fp16 head dim 64
  PB_ATTN_CASE16(2, 1, 1, 1, 4) PB_ATTN_CASE16(1, 1, 1, 0, 4) PB_ATTN_CASE16(1, 1, 1, 2, 4)   // fp16 head dim 80: 128 + 32 bytes
#undef PB_ATTN_CASE16
  if (const char* err = pbhost::optin_smem(kern, 227 * 1024)) return err;   // once per (device, instantiation)
  kern<<<grid, NTHREADS, smem, static_cast<cudaStream_t>(st)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
