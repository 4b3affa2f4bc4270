// Column-batched fused attention linearisation for the all-fp16 tangent plan (sm_100a: tcgen05 + TMEM + TMA).
//
// Same contract as pb_attn_sm100.cu (PbAttnLin with p16 = s16 = 1, see pb_kernels.h), different work decomposition: a CTA owns
// (head, 128-row tile, problem slot) and a GROUP of KC tangent columns, and walks the score columns in steps of 64.  Everything
// that does not depend on the tangent -- the probability tile (16 KB per step, the largest stream), the primal score operand
// (K or V rows) and the primal C1 tile (V^T, K^T or Q^T) -- is loaded ONCE per step and used by all KC columns; only the
// per-tangent tiles (dK, dV^T / Obar, Obar^T) are loaded per column.  The kernel of pb_attn_sm100.cu loads all of it per
// column and is paced by the L2 -> SM path (48 KB per column-step at d = 40, 27 B/clk/SM of the ~42 the L2 delivers; ncu
// r2c: tensor pipe 24 %, issue 44 %, the top stalls are the TMA-fed barriers); here a column-step needs 15 KB (JVP, KC = 5),
// 5.6 KB (VJP-A) or 19 KB (VJP-B, KC = 3).
//
// Per substep u = (step j, column c):
//     S_c   = sum_seg A_seg . B_seg^T                       (tcgen05.mma kind::f16 over the head dim -> TMEM ring)
//     Acc2_c (or Acc_c) += P(j) . C2_c(j)                   (optional)
//     T_c   = P(j) o (alpha1 S_c - delta_c)                  (CUDA cores: P row held in registers for the KC columns of the step,
//                                                            S from TMEM, T to a swizzled smem ring as halves)
//     Acc_c += T_c . C1(j)
// The row sums of T that the JVP needs come out of the tensor core: row d of every C1 stage is set to ones once (rows >= d
// of a C tile are never written by TMA), so accumulator column d is rowsum(T) of exactly the halves the MMA consumed.
// Head dim d <= 64; the score operands sit in compact k-chunks (d = 40: a 64-byte-swizzled tile of 32 halves + a
// 32-byte-swizzled tile of 16, 96 bytes per row instead of 128) -- shared memory is what bounds KC.
// Warp roles: 0 TMA producer (one thread polling the ring "empty" barriers), 1 score MMAs + TMEM allocator, 2..9 compute
// (two groups of four on alternate substeps; warp % 4 = TMEM lane quarter), 10 accumulating MMAs.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pb_host_util.h"
#include "pb_kernels.h"
#include "pb_tc.cuh"

namespace pbgemm {
extern int g_tmap_promo256;
const char* encode4x(CUtensorMap* m, const void* base, int f16, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4], int swizzle_bytes);
const char* encode_plainx(CUtensorMap* m, const void* base, int f16, int rows, int K, long ld, long sh, int nh, long sb,
                          int nb, int box_cols, int box_rows, int swizzle_bytes, int* hmul, int* bmul, uint32_t* bytes);
}

namespace pbattn16 {
using namespace pbtc;

constexpr int TM = 128, TN = 64;
constexpr int NTHREADS = 448;          // warp 0 TMA (A, shared stages), 13 TMA (per-column stages), 1 / 11 score MMAs, 2..9 compute, 10 / 12 accumulating MMAs
constexpr int TMEM_COLS = 512;
constexpr int MAX_KC = 8, MAX_NS = 4, MAX_NSH = 4, MAX_NPC = 6, MAX_NT = 4;
constexpr int PT_BYTES = TM * 128;     // P tile / T tile: 128 rows x 64 halves, 128-byte swizzle

// k-chunks of the score operands over the head dim: (swizzle span bytes, MMAs of 16 halves)
template <int KCFG> struct KCfg;
template <> struct KCfg<0> { static constexpr int SP0 = 32, N0 = 1, SP1 = 0, N1 = 0; };     // d <= 16
template <> struct KCfg<1> { static constexpr int SP0 = 64, N0 = 2, SP1 = 0, N1 = 0; };     // d <= 32
template <> struct KCfg<2> { static constexpr int SP0 = 64, N0 = 2, SP1 = 32, N1 = 1; };    // d <= 48
template <> struct KCfg<3> { static constexpr int SP0 = 128, N0 = 4, SP1 = 0, N1 = 0; };    // d <= 64

struct alignas(64) Params {
  CUtensorMap mapA[2][2], mapB[2][2];  // [segment][k-chunk]
  CUtensorMap mapC[2], mapC2[2], mapP; // C tiles per cluster rank (a CTA of a pair holds half of the rows)
  int a_pc[2], b_pc[2];                // operand is per tangent column (else primal: shared by the columns of the CTA)
  int a_hmul[2], b_hmul[2], a_bmul[2], b_bmul[2], a_smul[2], b_smul[2];   // head / tangent / slot coordinate multipliers
  int a_tile0[2], b_idx[2];            // first resident A tile of the segment; tile index of its B inside a stage
  int nA, nbsh, nbpc;                  // resident A tiles (at kc_max); shared / per-column B tiles per stage
  uint32_t a_bytes[2], sh_bytes[2], pc_bytes[2];   // stage bytes per cluster rank
  int pc_cnt;                          // consumers that release a per-column stage (score warp, accumulate warp)
  int ctile, accw, accw_tot;           // bytes of a C tile (clr rows); accumulator columns per tangent (x2 with D2)
  int clr;                             // rows of a C tile a CTA holds: accw, or accw / 2 in a CTA pair
  int ns, nsh, npc, nt;                // ring depths: S (TMEM), shared stage, per-column stage, T
  int k_slot, nslots, ngrp, kc_base, kc_rem, kc_max, ps_mul;
  int d, Mr, Nc, nh;
  float alpha1, alpha2, inv_pscale;
  const float* delta;
  int want_rsum; const float* O; long ldo, o_stride;
  __half* D; long ldd, sDb;
  __half* D2; long ldd2, sD2b;
  long long* trace;                    // PB_ATTN_TRACE: per-substep event clocks of CTA (0, 0), [512][16]
};

// K-major swizzled smem matrix descriptor (swizzle span 128 / 64 / 32 bytes, 8-row groups 8 * span apart), split into the
// low word (14-bit start address in 16-byte units + LBO) and the constant high word (SBO, version, layout)
__host__ __device__ constexpr uint32_t desc_hi(int span) {
  return uint32_t((8 * span) >> 4) | (1u << 14) | (uint32_t(span == 128 ? 2 : span == 64 ? 4 : 6) << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return (uint64_t(hi) << 32) | lo; }

// peer handshakes of the CTA-pair variant: remote arrive with the default semantics and a plain wait, as CUTLASS' ClusterBarrier
#define PEER_ARRIVE mbar_arrive_cluster
#define PEER_WAIT mbar_wait
#define TR(u_, e_) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (u_) < 512) p.trace[(u_) * 16 + (e_)] = clock64(); } while (0)

struct Ring {
  int idx; uint32_t ph;
  __device__ __forceinline__ void next(int n) { if (++idx == n) { idx = 0; ph ^= 1u; } }
  __device__ __forceinline__ void next2(int n) { idx += 2; if (idx >= n) { idx -= n; ph ^= 1u; } }   // n >= 2
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
      : "memory");
}
// two floats -> packed halves, round to nearest, saturating at +-65504 instead of producing inf
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <bool PAIR>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (PAIR) mma_f16_2sm(d, a, b, idesc, acc); else mma_f16(d, a, b, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (PAIR) tcgen05_commit_2sm(bar); else tcgen05_commit(bar);
}

// NSEG score segments, KCFG k-chunk layout, C2M: 0 no second product, 1 folded into Acc, 2 separate accumulator / output,
// DM delta mode: 0 none, 1 per row, 2 per column
// PAIR: two CTAs of a cluster (row tiles 2i, 2i + 1 of the same head / slot / column group, on the two SMs of a TPC) run every
// product as ONE tcgen05.mma.cta_group::2 of 256 rows issued by the leader (cluster rank 0).  A one-CTA MMA of these shapes
// costs 67-72 clocks (N = 48 / 64, A from shared memory: ~38 clocks per instruction on top of the math), a pair MMA 41-43 for
// both SMs' work (scripts/mma_rate.cu); with 14 MMAs per substep the one-CTA kernel sits at its instruction roofline.  Each
// CTA keeps its own rows (A tiles, P, T, accumulators, epilogue) and HALF of every B operand (32 of the 64 K / dK rows of a
// step, half of the C tile rows).  Protocol: the peer CTA runs the same warp programs with the same local barriers; where the
// leader's issuing warp would issue, the peer's twin has done the same waits on ITS barriers and arrives on a leader
// barrier (peer_s / peer_c2 / peer_t, indexed like the ring stage the product uses), which the leader waits for as well; the
// leader's commits arrive on the barriers of both CTAs (multicast), so the peer's rings advance exactly as the leader's.
template <int NSEG, int KCFG, int C2M, int DM, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1) attn16_kernel(const __grid_constant__ Params p) {
  using KF = KCfg<KCFG>;
  constexpr int SP0 = KF::SP0, N0 = KF::N0, SP1 = KF::SP1, N1 = KF::N1;
  constexpr int BROWS = PAIR ? TN / 2 : TN;                 // rows of a score-operand B tile this CTA holds
  constexpr int A_TILE = TM * (SP0 + SP1), B_TILE = BROWS * (SP0 + SP1);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr bool HAS_C2 = C2M != 0, SEP = C2M == 2;
  constexpr int LAG = HAS_C2 ? 2 : 0;          // the T . C1 product of substep u is issued with the P . C2 product of u + LAG

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.ns, NSH = p.nsh, NPC = p.npc, NT = p.nt;
  const bool has_pc = p.pc_cnt > 0;
  uint8_t* sA = smem;
  uint8_t* sBsh = sA + p.nA * A_TILE;
  uint8_t* sBpc = sBsh + NSH * p.nbsh * B_TILE;
  uint8_t* sC1 = sBpc + NPC * p.nbpc * B_TILE;
  uint8_t* sC2 = sC1 + NSH * p.ctile;
  uint8_t* sP = sC2 + (HAS_C2 ? NPC * p.ctile : 0);
  uint8_t* sT = sP + NSH * PT_BYTES;
  float* s_dcol = reinterpret_cast<float*>(sT + NT * PT_BYTES);      // [8 compute warps][64] column deltas of the substep
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dcol + 8 * 64);
  uint64_t* a_full = bars + 0;
  uint64_t* acc_full = bars + 1;
  uint64_t* sh_full = bars + 2;                 // [MAX_NSH]
  uint64_t* sh_empty = sh_full + MAX_NSH;
  uint64_t* pc_full = sh_empty + MAX_NSH;       // [MAX_NPC]
  uint64_t* pc_empty = pc_full + MAX_NPC;
  uint64_t* s_full = pc_empty + MAX_NPC;        // [MAX_NS]
  uint64_t* s_free = s_full + MAX_NS;
  uint64_t* t_full = s_free + MAX_NS;           // [MAX_NT]
  uint64_t* t_empty = t_full + MAX_NT;
  uint64_t* acc_zeroed = t_empty + MAX_NT;
  uint64_t* peer_a = acc_zeroed + 1;            // CTA pair, on the leader: the peer's A tiles landed
  uint64_t* peer_z = peer_a + 1;                // the peer's accumulators are zeroed (its 8 compute warps)
  uint64_t* peer_s = peer_z + 1;                // [MAX_NS]  the peer is ready for the score product into this S stage
  uint64_t* peer_c2 = peer_s + MAX_NS;          // [MAX_NPC] ... for the P . C2 product of this per-column stage
  uint64_t* peer_t = peer_c2 + MAX_NPC;         // [MAX_NT]  ... for the T . C1 product of this T stage
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(peer_t + MAX_NT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * TM;
  int y = blockIdx.y;
  const int gi = y % p.ngrp; y /= p.ngrp;
  const int slot = y % p.nslots;
  const int bat_h = y / p.nslots;
  const int KC = p.kc_base + (gi < p.kc_rem ? 1 : 0);                 // tangent columns of this CTA
  const int b0 = slot * p.k_slot + gi * p.kc_base + min(gi, p.kc_rem);   // its first tangent
  const int nj = (p.Nc + TN - 1) / TN;
  const int nu = nj * KC;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1); mbar_init(acc_full, 2); mbar_init(acc_zeroed, 8);
    for (int i = 0; i < MAX_NSH; ++i) { mbar_init(&sh_full[i], 1); mbar_init(&sh_empty[i], 12); }
    for (int i = 0; i < MAX_NPC; ++i) { mbar_init(&pc_full[i], 1); mbar_init(&pc_empty[i], max(p.pc_cnt, 1)); }
    for (int i = 0; i < MAX_NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
    for (int i = 0; i < MAX_NT; ++i) { mbar_init(&t_full[i], 4); mbar_init(&t_empty[i], 1); mbar_init(&peer_t[i], 1); }
    mbar_init(peer_a, 1); mbar_init(peer_z, 8);
    for (int i = 0; i < MAX_NS; ++i) mbar_init(&peer_s[i], 1);
    for (int i = 0; i < MAX_NPC; ++i) mbar_init(&peer_c2[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // rows d .. accw-1 of every C tile stage are outside the TMA box: zero them once; with want_rsum row d of the C1 stages is
    // ones, so that accumulator column d collects rowsum(T).  (A row of equal 16-byte chunks is swizzle-invariant.)
    // (a CTA of a pair holds rows rank * clr .. of the tile: its constant rows start where the TMA box ends)
    const int lr = p.clr;
    const int first = min(lr, max(0, p.d - int(rank) * lr));
    const int xrows = lr - first;
    const int per_stage = xrows * 8;                                   // 16-byte chunks
    const int n1 = NSH * per_stage, n2 = HAS_C2 ? NPC * per_stage : 0;
    const uint32_t one2 = 0x3C003C00u;
    for (int i = threadIdx.x; per_stage > 0 && i < n1 + n2; i += NTHREADS) {
      const bool c1 = i < n1;
      const int k = c1 ? i : i - n1;
      const int stg = k / per_stage, rem = k % per_stage;
      uint8_t* base = (c1 ? sC1 : sC2) + stg * p.ctile + first * 128 + rem * 16;
      const uint32_t v = (c1 && p.want_rsum && int(rank) * lr + first + (rem >> 3) == p.d) ? one2 : 0u;
      *reinterpret_cast<uint4*>(base) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();       // both CTAs' barriers are initialised before either signals the other's
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // =========================== TMA producer: resident A tiles, then the shared stages ===========================
    // The whole warp walks the loop and ONE ELECTED lane issues: inside an `if (lane == 0)` region the compiler cannot prove the
    // TMA operands warp-uniform and wraps every UTMALDG in an ELECT / R2UR.BROADCAST waterfall (~150 clocks each; a thread that
    // fed both rings that way issued a stage every ~850 clocks and paced the CTA, traced with PB_ATTN_TRACE).
    {
      uint32_t abytes = 0;
#pragma unroll
      for (int s = 0; s < NSEG; ++s) abytes += p.a_bytes[s] * (p.a_pc[s] ? KC : 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(a_full, abytes);
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
          const int ncol = p.a_pc[s] ? KC : 1;
          for (int c = 0; c < ncol; ++c) {
            uint8_t* dst = sA + (p.a_tile0[s] + c) * A_TILE;
            const int bc = p.a_pc[s] ? (b0 + c) * p.a_bmul[s] : slot * p.a_smul[s];
            tma_load_4d(dst, &p.mapA[s][0], a_full, 0, r0, bat_h * p.a_hmul[s], bc);
            if (N1) tma_load_4d(dst + TM * SP0, &p.mapA[s][1], a_full, SP0 / 2, r0, bat_h * p.a_hmul[s], bc);
          }
        }
      }
      __syncwarp();
      Ring rsh{0, 0};
      const int ps = slot * p.ps_mul;
      int hb[NSEG], sb[NSEG]; uint32_t ob[NSEG]; bool shb[NSEG];
#pragma unroll
      for (int s = 0; s < NSEG; ++s) {
        hb[s] = bat_h * p.b_hmul[s]; sb[s] = slot * p.b_smul[s]; ob[s] = uint32_t(p.b_idx[s]) * B_TILE; shb[s] = !p.b_pc[s];
      }
      const uint32_t sh_bytes = p.sh_bytes[rank], ctile = p.ctile, shB = uint32_t(p.nbsh) * B_TILE;
      const int brow = int(rank) * BROWS, crow = int(rank) * p.clr;
      for (int msh = 0; msh < nj; ++msh) {
        mbar_wait(&sh_empty[rsh.idx], rsh.ph ^ 1);
        const int st = rsh.idx;
        if (elect_one()) {
          TR(msh * KC, 11);
          mbar_arrive_expect_tx(&sh_full[st], sh_bytes);
          tma_load_4d(sP + st * PT_BYTES, &p.mapP, &sh_full[st], msh * TN, r0, bat_h, ps);
#pragma unroll
          for (int s = 0; s < NSEG; ++s)
            if (shb[s]) {
              uint8_t* dst = sBsh + st * shB + ob[s];
              tma_load_4d(dst, &p.mapB[s][0], &sh_full[st], 0, msh * TN + brow, hb[s], sb[s]);
              if (N1) tma_load_4d(dst + BROWS * SP0, &p.mapB[s][1], &sh_full[st], SP0 / 2, msh * TN + brow, hb[s], sb[s]);
            }
          tma_load_4d(sC1 + st * ctile, &p.mapC[rank], &sh_full[st], msh * TN, crow, bat_h, ps);
        }
        __syncwarp();
        rsh.next(NSH);
      }
    }
  } else if (warp == 13) {
    // =========================== TMA producer of the per-column stages ===========================
    if (has_pc) {
      Ring rpc{0, 0};
      int hb[NSEG], bm[NSEG]; uint32_t ob[NSEG]; bool pcb[NSEG];
#pragma unroll
      for (int s = 0; s < NSEG; ++s) {
        hb[s] = bat_h * p.b_hmul[s]; bm[s] = p.b_bmul[s]; ob[s] = uint32_t(p.b_idx[s]) * B_TILE; pcb[s] = p.b_pc[s] != 0;
      }
      const uint32_t pc_bytes = p.pc_bytes[rank], ctile = p.ctile, pcB = uint32_t(p.nbpc) * B_TILE;
      const int brow = int(rank) * BROWS, crow = int(rank) * p.clr;
      int pj = 0, pcol = 0;
      for (int mpc = 0; mpc < nu; ++mpc) {
        mbar_wait(&pc_empty[rpc.idx], rpc.ph ^ 1);
        const int st = rpc.idx;
        if (elect_one()) {
          TR(mpc, 10);
          mbar_arrive_expect_tx(&pc_full[st], pc_bytes);
#pragma unroll
          for (int s = 0; s < NSEG; ++s)
            if (pcb[s]) {
              uint8_t* dst = sBpc + st * pcB + ob[s];
              const int bc = (b0 + pcol) * bm[s];
              tma_load_4d(dst, &p.mapB[s][0], &pc_full[st], 0, pj * TN + brow, hb[s], bc);
              if (N1) tma_load_4d(dst + BROWS * SP0, &p.mapB[s][1], &pc_full[st], SP0 / 2, pj * TN + brow, hb[s], bc);
            }
          if (HAS_C2) tma_load_4d(sC2 + st * ctile, &p.mapC2[rank], &pc_full[st], pj * TN, crow, bat_h, b0 + pcol);
        }
        __syncwarp();
        rpc.next(NPC);
        if (++pcol == KC) { pcol = 0; ++pj; }
      }
    }
  } else if (warp == 1 || warp == 11) {
    // =========================== score MMAs ===========================
    // Two issuing warps take alternate substeps.  A single warp's instruction stream (three barrier probes, elect, descriptor
    // arithmetic in the uniform datapath, six MMAs, three commits: ~1300 clocks per substep, traced with PB_ATTN_TRACE) paced the
    // whole CTA while the tensor pipe was 27 % busy; everything loop-invariant is hoisted and a descriptor is one 32-bit add
    // away: desc = {lo + offset16, HI(span)} -- only the 14-bit address field of the low word ever changes.
    const int w = warp == 1 ? 0 : 1;
    const uint32_t idesc_s = (1u << 4) | (uint32_t(TN >> 3) << 17) | (uint32_t((PAIR ? 2 * TM : TM) >> 4) << 24);
    const uint32_t peer_s_l = PAIR ? mapa_u32(smem_u32(peer_s), 0) : 0u, peer_a_l = PAIR ? mapa_u32(smem_u32(peer_a), 0) : 0u;
    constexpr uint32_t HI0 = desc_hi(SP0), HI1 = desc_hi(SP1 ? SP1 : 32);
    const uint32_t uA = smem_u32(sA), uBsh = smem_u32(sBsh), uBpc = smem_u32(sBpc);
    uint32_t a_lo[NSEG], a_cs[NSEG], b_lo[NSEG], b_st[NSEG], b_sel[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      a_lo[s] = desc_lo(uA) + uint32_t(p.a_tile0[s]) * (A_TILE >> 4);
      a_cs[s] = p.a_pc[s] ? (A_TILE >> 4) : 0;
      b_sel[s] = p.b_pc[s] ? 1u : 0u;
      b_lo[s] = desc_lo(b_sel[s] ? uBpc : uBsh) + uint32_t(p.b_idx[s]) * (B_TILE >> 4);
      b_st[s] = uint32_t(b_sel[s] ? p.nbpc : p.nbsh) * (B_TILE >> 4);
    }
    const bool wait_pc = p.nbpc != 0;
    Ring rsh{0, 0}, rpc{w, 0}, rs{w, 0};
    int j = w / KC, c = w % KC;
    int jr = 0, jdone = -1, jw = -1;          // step rsh points at / last step released by a commit / last step waited for
    mbar_wait(a_full, 0);
    if constexpr (PAIR) {
      if (leader) PEER_WAIT(peer_a, 0);
      else if (w == 0 && lane == 0) PEER_ARRIVE(peer_a_l);
    }
    for (int u = w; u < nu; u += 2) {
      while (jr < j) {                         // steps this warp had no substep in are released by a plain arrive
        if (jdone != jr) { mbar_wait(&sh_full[rsh.idx], rsh.ph); if (lane == 0) mbar_arrive(&sh_empty[rsh.idx]); }
        rsh.next(NSH); ++jr;
      }
      // the per-column tile is the operand that arrives last (its ring is the shallow one): every other wait comes first, and
      // its stage is the first thing released after the MMAs
      mbar_wait(&s_free[rs.idx], rs.ph ^ 1);
      if (jw != j) { mbar_wait(&sh_full[rsh.idx], rsh.ph); jw = j; }
      if (lane == 0) TR(u, 1);
      if (wait_pc) mbar_wait(&pc_full[rpc.idx], rpc.ph);
      if (lane == 0) TR(u, 0);
      if constexpr (PAIR) {                    // the peer's twin has done the same waits on its own barriers
        if (leader) PEER_WAIT(&peer_s[rs.idx], rs.ph);
        else if (lane == 0) PEER_ARRIVE(peer_s_l + uint32_t(rs.idx) * 8u);
      }
      tcgen05_fence_after();
      const bool last = c + 2 >= KC;           // this warp's last substep of step j
      if (leader && elect_one()) {
        const uint32_t d_s = tmem_base + rs.idx * TN;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
          const uint32_t al = a_lo[s] + uint32_t(c) * a_cs[s];
          const uint32_t bl = b_lo[s] + uint32_t(b_sel[s] ? rpc.idx : rsh.idx) * b_st[s];
#pragma unroll
          for (int k = 0; k < N0; ++k) mma<PAIR>(d_s, mk_desc(al + 2 * k, HI0), mk_desc(bl + 2 * k, HI0), idesc_s, (s | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < N1; ++k)
            mma<PAIR>(d_s, mk_desc(al + (TM * SP0 >> 4) + 2 * k, HI1), mk_desc(bl + (BROWS * SP0 >> 4) + 2 * k, HI1), idesc_s, 1u);
        }
        if (wait_pc) commit<PAIR>(&pc_empty[rpc.idx]);
        commit<PAIR>(&s_full[rs.idx]);
        if (last) commit<PAIR>(&sh_empty[rsh.idx]);
      }
      __syncwarp();
      if (lane == 0) TR(u, 2);
      if (last) jdone = j;
      if (wait_pc) rpc.next2(NPC);
      rs.next2(NS);
      c += 2;
      while (c >= KC) { c -= KC; ++j; }
    }
    while (jr < nj) {
      if (jdone != jr) { mbar_wait(&sh_full[rsh.idx], rsh.ph); if (lane == 0) mbar_arrive(&sh_empty[rsh.idx]); }
      rsh.next(NSH); ++jr;
    }
  } else if (warp == 10 || warp == 12) {
    // =========================== accumulating MMAs ===========================
    // Two issuing warps on alternate substeps: P . C2_c of substep u, then T_c . C1 of substep u - LAG.  The accumulators are
    // zeroed by the compute warps up front (acc_zeroed) and every product accumulates, so the two warps need no ordering
    // between them ("first write" vs "accumulate" on the same columns).
    const int w = warp == 10 ? 0 : 1;
    const uint32_t idesc_a = (1u << 4) | (uint32_t(p.accw >> 3) << 17) | (uint32_t((PAIR ? 2 * TM : TM) >> 4) << 24);
    const uint32_t peer_c2_l = PAIR ? mapa_u32(smem_u32(peer_c2), 0) : 0u, peer_t_l = PAIR ? mapa_u32(smem_u32(peer_t), 0) : 0u;
    constexpr uint32_t HI = desc_hi(128);
    const uint32_t u_acc = tmem_base + NS * TN;
    const uint32_t lC1 = desc_lo(smem_u32(sC1)), lC2 = desc_lo(smem_u32(sC2)), lP = desc_lo(smem_u32(sP)), lT = desc_lo(smem_u32(sT));
    const uint32_t ct16 = p.ctile >> 4;
    const uint32_t accw_tot = p.accw_tot, acc2_off = SEP ? p.accw : 0;
    Ring rshp{0, 0}, rsha{0, 0}, rpc{w, 0}, rt{w, 0};
    int jp = w / KC, cp = w % KC, jrp = 0, jwp = -1;              // P . C2 cursor and its view of the shared-stage ring
    int ja = w / KC, ca = w % KC, jra = 0, jwa = -1, jdone = -1;  // T . C1 cursor (this one releases the shared stages)
    mbar_wait(acc_zeroed, 0);
    if constexpr (PAIR) {
      if (leader) PEER_WAIT(peer_z, 0);
    }
    tcgen05_fence_after();
    for (int u = w; u < nu + LAG; u += 2) {
      if (u >= LAG) {                             // steps the T . C1 cursor skips are released first: the P . C2 product below
        while (jra < ja) {                        // may be waiting for a stage that only this arrive frees
          if (jdone != jra) { mbar_wait(&sh_full[rsha.idx], rsha.ph); if (lane == 0) mbar_arrive(&sh_empty[rsha.idx]); }
          rsha.next(NSH); ++jra;
        }
      }
      if (HAS_C2 && u < nu) {                     // Acc2_c (or Acc_c) += P(j) . C2_c(j)
        while (jrp < jp) { rshp.next(NSH); ++jrp; }
        if (jwp != jp) { mbar_wait(&sh_full[rshp.idx], rshp.ph); jwp = jp; }
        mbar_wait(&pc_full[rpc.idx], rpc.ph);
        if (lane == 0) TR(u, 3);
        if constexpr (PAIR) {
          if (leader) PEER_WAIT(&peer_c2[rpc.idx], rpc.ph);
          else if (lane == 0) PEER_ARRIVE(peer_c2_l + uint32_t(rpc.idx) * 8u);
        }
        tcgen05_fence_after();
        const uint32_t al = lP + uint32_t(rshp.idx) * (PT_BYTES >> 4);
        const uint32_t bl = lC2 + uint32_t(rpc.idx) * ct16;
        const uint32_t tacc = u_acc + uint32_t(cp) * accw_tot + acc2_off;
        if (leader && elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma<PAIR>(tacc, mk_desc(al + 2 * k, HI), mk_desc(bl + 2 * k, HI), idesc_a, 1u);
          commit<PAIR>(&pc_empty[rpc.idx]);
        }
        __syncwarp();
        rpc.next2(NPC);
        cp += 2;
        while (cp >= KC) { cp -= KC; ++jp; }
      }
      if (u >= LAG) {                             // Acc_c += T_c . C1(j) of substep u - LAG
        if (jwa != ja) { mbar_wait(&sh_full[rsha.idx], rsha.ph); jwa = ja; }
        mbar_wait(&t_full[rt.idx], rt.ph);
        if (lane == 0) TR(u - LAG, 4);
        if constexpr (PAIR) {
          if (leader) PEER_WAIT(&peer_t[rt.idx], rt.ph);
          else if (lane == 0) PEER_ARRIVE(peer_t_l + uint32_t(rt.idx) * 8u);
        }
        tcgen05_fence_after();
        const uint32_t al = lT + uint32_t(rt.idx) * (PT_BYTES >> 4);
        const uint32_t bl = lC1 + uint32_t(rsha.idx) * ct16;
        const uint32_t tacc = u_acc + uint32_t(ca) * accw_tot;
        const bool last = ca + 2 >= KC;
        if (leader && elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma<PAIR>(tacc, mk_desc(al + 2 * k, HI), mk_desc(bl + 2 * k, HI), idesc_a, 1u);
          commit<PAIR>(&t_empty[rt.idx]);
          if (last) commit<PAIR>(&sh_empty[rsha.idx]);
        }
        __syncwarp();
        if (last) jdone = ja;
        rt.next2(NT);
        ca += 2;
        while (ca >= KC) { ca -= KC; ++ja; }
      }
    }
    while (jra < nj) {
      if (jdone != jra) { mbar_wait(&sh_full[rsha.idx], rsha.ph); if (lane == 0) mbar_arrive(&sh_empty[rsha.idx]); }
      rsha.next(NSH); ++jra;
    }
    if (leader && elect_one()) commit<PAIR>(acc_full);
    __syncwarp();
  } else {
    // =========================== compute warps ===========================
    const int cw = warp - 2;
    const int grp = cw >> 2;                  // 0 / 1: takes the even / odd substeps
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
    const int row = q * 32 + lane;
    const int r = r0 + row;
    const bool row_ok = r < p.Mr;
    const uint32_t tm_row = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t swz = uint32_t(row & 7);
    const uint32_t uP = smem_u32(sP) + row * 128, uT = smem_u32(sT) + row * 128;
    float* sd = s_dcol + cw * 64;
    {
      // zero the accumulators of this warp's TMEM lane quarter (the two groups split the 16-column chunks)
      const int nchunk = (KC * p.accw_tot) >> 4;
      for (int ch = grp; ch < nchunk; ch += 2) tmem_st16_zero(tm_row + NS * TN + ch * 16);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_zeroed);
        if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(peer_z), 0));
      }
    }
    uint4 pv[8];
    Ring rsh{0, 0}, rs{grp, 0}, rt{grp, 0};
    int jn = 0;                               // next step whose shared stage this warp has not released yet
    int c = grp % KC, j = grp / KC;
    // the deltas of a substep are fetched one substep ahead (a global load at the top of the substep sat on the critical path
    // of every compute warp: VJP-B ran 25 % behind the JVP for less tensor work)
    float dnext0 = 0.f, dnext1 = 0.f;
    auto fetch_delta = [&](int jj, int cc) {
      const int bb = b0 + cc;
      if (DM == 1) {
        dnext0 = row_ok ? __ldg(p.delta + ((long)bb * p.nh + bat_h) * p.Mr + r) : 0.f;
      } else if (DM == 2) {
        const float* dbase = p.delta + ((long)bb * p.nh + bat_h) * p.Nc;
        const int c0 = jj * TN + lane, c1 = c0 + 32;
        dnext0 = c0 < p.Nc ? __ldg(dbase + c0) : 0.f;
        dnext1 = c1 < p.Nc ? __ldg(dbase + c1) : 0.f;
      }
    };
    if (grp < nu) fetch_delta(j, c);
    for (int u = grp; u < nu; u += 2) {
      if (q == 0 && lane == 0) TR(u, 9);
      const float drow = dnext0;
      if (DM == 2) {                          // one coalesced load per warp (issued a substep ago), read back as broadcasts
        sd[lane] = dnext0;
        sd[32 + lane] = dnext1;
        __syncwarp();
      }
      {
        int cn = c + 2, jx = j;
        while (cn >= KC) { cn -= KC; ++jx; }
        if (u + 2 < nu) fetch_delta(jx, cn);
      }
      while (jn <= j) {                       // every compute warp releases every shared stage, used or not
        mbar_wait(&sh_full[rsh.idx], rsh.ph);
        if (jn == j) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = lds128(uP + rsh.idx * PT_BYTES + ((uint32_t(i) ^ swz) << 4));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh_empty[rsh.idx]);
        rsh.next(NSH); ++jn;
      }
      mbar_wait(&s_full[rs.idx], rs.ph);
      if (q == 0 && lane == 0) TR(u, 5);
      tcgen05_fence_after();
#pragma unroll
      for (int hv = 0; hv < 2; ++hv) {
        uint32_t sv[32];
        tmem_ld32(tm_row + rs.idx * TN + hv * 32, sv);
        if (hv == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[rs.idx]);
        }
        uint4 tv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ci = hv * 4 + i;
          const __half2* ph2 = reinterpret_cast<const __half2*>(&pv[ci]);
          float dl[8];
          if (DM == 2) {
            const float4 da = *reinterpret_cast<const float4*>(sd + ci * 8), db = *reinterpret_cast<const float4*>(sd + ci * 8 + 4);
            dl[0] = da.x; dl[1] = da.y; dl[2] = da.z; dl[3] = da.w; dl[4] = db.x; dl[5] = db.y; dl[6] = db.z; dl[7] = db.w;
          }
          uint32_t oh[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 pf = __half22float2(ph2[e2]);
            const float s0 = __uint_as_float(sv[i * 8 + e2 * 2]), s1 = __uint_as_float(sv[i * 8 + e2 * 2 + 1]);
            float t0, t1;
            if (DM == 0) { t0 = pf.x * (p.alpha1 * s0); t1 = pf.y * (p.alpha1 * s1); }
            else if (DM == 1) { t0 = pf.x * fmaf(p.alpha1, s0, -drow); t1 = pf.y * fmaf(p.alpha1, s1, -drow); }
            else { t0 = pf.x * fmaf(p.alpha1, s0, -dl[e2 * 2]); t1 = pf.y * fmaf(p.alpha1, s1, -dl[e2 * 2 + 1]); }
            oh[e2] = pack_h2_sat(t0, t1);
          }
          tv[i] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        }
        if (hv == 0) {                          // the T stage is written half by half (registers): its previous reader must be done
          if (q == 0 && lane == 0) TR(u, 6);
          mbar_wait(&t_empty[rt.idx], rt.ph ^ 1);
          if (q == 0 && lane == 0) TR(u, 7);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(uT + rt.idx * PT_BYTES + ((uint32_t(hv * 4 + i) ^ swz) << 4), tv[i]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[rt.idx]);
      if (q == 0 && lane == 0) TR(u, 8);
      rs.next2(NS); rt.next2(NT);
      c += 2;
      while (c >= KC) { c -= KC; ++j; }
    }
    while (jn < nj) {                         // steps in which this group had no substep
      mbar_wait(&sh_full[rsh.idx], rsh.ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh_empty[rsh.idx]);
      rsh.next(NSH); ++jn;
    }
    // ---- epilogue: columns c = grp, grp + 2, ...:  D = alpha2 * Acc - rowsum(T) o O ;  D2 = Acc2 ----
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const long rr = row_ok ? r : 0;
    const int dm16 = p.d & 15, dc16 = p.d & ~15;
    for (int cc = grp; cc < KC; cc += 2) {
      const int b = b0 + cc;
      const uint32_t tacc = tm_row + NS * TN + cc * p.accw_tot;
      float rsum = 0.f;
      if (p.want_rsum) {
        uint32_t v[16];
        tmem_ld16(tacc + dc16, v);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (e == dm16) rsum = __uint_as_float(v[e]) * p.inv_pscale;
      }
      const float* optr = (p.want_rsum && p.O) ? p.O + slot * p.o_stride + rr * p.ldo + bat_h * p.d : nullptr;
#pragma unroll 1
      for (int half = 0; half < (SEP ? 2 : 1); ++half) {
        __half* hptr = (half == 0 ? p.D + (long)b * p.sDb + rr * p.ldd : p.D2 + (long)b * p.sD2b + rr * p.ldd2) + bat_h * p.d;
        const float alpha = (half == 0 ? p.alpha2 : 1.f) * p.inv_pscale;
        const uint32_t ta = tacc + (half ? p.accw : 0);
        for (int c16 = 0; c16 < p.d; c16 += 16) {
          uint32_t v[16];
          tmem_ld16(ta + c16, v);
#pragma unroll
          for (int g = 0; g < 16; g += 8) {
            const int n = c16 + g;
            if (n >= p.d || !row_ok) break;              // d is a multiple of 8
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = alpha * __uint_as_float(v[g + e]);
            if (half == 0 && optr) {
              const float4 oa = *reinterpret_cast<const float4*>(optr + n), ob = *reinterpret_cast<const float4*>(optr + n + 4);
              o[0] -= rsum * oa.x; o[1] -= rsum * oa.y; o[2] -= rsum * oa.z; o[3] -= rsum * oa.w;
              o[4] -= rsum * ob.x; o[5] -= rsum * ob.y; o[6] -= rsum * ob.z; o[7] -= rsum * ob.w;
            }
            uint4 hv;
            *reinterpret_cast<__half2*>(&hv.x) = __floats2half2_rn(o[0], o[1]);
            *reinterpret_cast<__half2*>(&hv.y) = __floats2half2_rn(o[2], o[3]);
            *reinterpret_cast<__half2*>(&hv.z) = __floats2half2_rn(o[4], o[5]);
            *reinterpret_cast<__half2*>(&hv.w) = __floats2half2_rn(o[6], o[7]);
            *reinterpret_cast<uint4*>(hptr + n) = hv;
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();       // no CTA leaves while its peer may still signal its barriers or read its tiles
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static long long* g_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int pbk_attn16_trace_read(long long* host, int n) {   // debugging aid (scripts/trace_attn.py)
  if (!g_trace) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host, g_trace, sizeof(long long) * std::min(n, 512 * 16), cudaMemcpyDeviceToHost);
  return std::min(n, 512 * 16);
}

// nullptr: launched; *handled = false: geometry / role outside this kernel (the caller falls back to pb_attn_sm100.cu)
const char* launch(const PbAttnLin& a, cudaStream_t st, bool* handled) {
  *handled = false;
  static const bool off = getenv("PB_ATTN_V1") != nullptr;             // A/B switch: the per-column kernel only
  static const int kc_cap = getenv("PB_ATTN_KC") ? std::max(1, atoi(getenv("PB_ATTN_KC"))) : MAX_KC;
  if (off || !a.s16 || !a.p16 || a.R) return nullptr;
  if (a.d % 8 || a.d < 8 || a.d > 64 || a.Nc < TN || a.Mr < TM) return nullptr;
  const int c2m = a.C2 ? (a.D2 ? 2 : 1) : 0;
  const int dm = a.delta ? a.delta_mode : 0;
  int role = -1;                                                        // 0 JVP, 1 cross JVP / no delta, 2 VJP-A, 3 VJP-B
  if (a.nseg == 2 && c2m == 1 && dm == 0) role = 0;
  if (a.nseg == 1 && c2m == 0 && dm == 0) role = 1;
  if (a.nseg == 1 && c2m == 0 && dm == 1) role = 2;
  if (a.nseg == 1 && c2m == 2 && dm == 2) role = 3;
  if (role < 0) return nullptr;
  if (!(a.p_scale > 0.f)) return "attn_lin: p_scale must be positive";
  const int kcfg = a.d <= 16 ? 0 : a.d <= 32 ? 1 : a.d <= 48 ? 2 : 3;
  const int sp0 = kcfg == 0 ? 32 : kcfg == 3 ? 128 : 64, sp1 = kcfg == 2 ? 32 : 0;
  // CTA pairs (see the kernel): an even number of row tiles, and a C tile whose second half still starts inside the d rows
  // OPT-IN (PB_ATTN_PAIR=1): correct (the GPU tests pass with it on) but not faster -- 4096-token JVP 1.78 ms against 1.66 ms
  // one-CTA, although its MMAs cost 586 instead of 971 clocks per substep: the dependency loop s_free -> score MMA -> compute
  // -> T . C1 -> s_free (two S stages per parity class, TMEM is full) already takes ~890 clocks per substep in the one-CTA
  // kernel, and every product of the pair adds a cross-CTA handshake to it
  static const int pair_env = getenv("PB_ATTN_PAIR") ? atoi(getenv("PB_ATTN_PAIR")) : 0;
  const int accw0 = (a.d + (a.want_rsum ? 1 : 0) + 15) / 16 * 16;
  const bool pair = pair_env && ((a.Mr + TM - 1) / TM) % 2 == 0 && a.d > accw0 / 2 && a.Nc % 8 == 0;
  const int brows = pair ? TN / 2 : TN;
  const int a_tile = TM * (sp0 + sp1), b_tile = brows * (sp0 + sp1);

  Params p;
  memset(&p, 0, sizeof p);
  p.d = a.d; p.Mr = a.Mr; p.Nc = a.Nc; p.nh = a.nh;
  p.alpha1 = a.alpha1; p.alpha2 = a.alpha2; p.inv_pscale = 1.f / a.p_scale;
  p.delta = a.delta;
  p.want_rsum = a.want_rsum ? 1 : 0; p.O = a.O; p.ldo = a.ldo;
  p.D = reinterpret_cast<__half*>(a.D); p.ldd = a.ldd; p.sDb = a.sDb;
  p.D2 = reinterpret_cast<__half*>(a.D2); p.ldd2 = a.ldd2; p.sD2b = a.sD2b;
  const int k_slot = (a.k_slot > 0 && a.k_slot < a.nb) ? a.k_slot : a.nb;
  const int nslots = a.nb / k_slot;
  if (a.nb % k_slot) return "attn_lin: the tangent batch must be a whole number of problem slots";
  if (nslots > 1 && (a.p_stride <= 0 || a.p_stride % 16)) return "attn_lin: p_stride must be a positive multiple of 16 bytes";
  p.k_slot = k_slot; p.nslots = nslots; p.ps_mul = nslots > 1 ? 1 : 0; p.o_stride = nslots > 1 ? a.p_stride / 4 : 0;
  if ((a.ldd % 8) || (a.O && a.ldo % 4) || (a.ldp % 8) || (a.ldc % 8) || (a.D2 && a.ldd2 % 8) || (a.C2 && a.ldc2 % 8) ||
      ((reinterpret_cast<uintptr_t>(a.D) | reinterpret_cast<uintptr_t>(a.O) | reinterpret_cast<uintptr_t>(a.Pm) |
        reinterpret_cast<uintptr_t>(a.D2) | reinterpret_cast<uintptr_t>(a.C2)) & 15))
    return "attn_lin: D/D2/O/P/C2 must be 16-byte aligned with rows that are multiples of 16 bytes";
  p.accw = (a.d + (p.want_rsum ? 1 : 0) + 15) / 16 * 16;
  p.accw_tot = p.accw * (c2m == 2 ? 2 : 1);
  p.clr = pair ? p.accw / 2 : p.accw;
  p.ctile = p.clr * 128;

  // operands: per tangent column (batch stride != 0) or primal
  int nA_fixed = 0, nA_pc = 0;
  uint32_t sh_b = 0, pc_b = 0;                                          // stage bytes common to both ranks (C tiles added below)
  for (int s = 0; s < a.nseg; ++s) {
    p.a_pc[s] = a.seg[s].sAb != 0 ? 1 : 0;
    p.b_pc[s] = a.seg[s].sBb != 0 ? 1 : 0;
    if (p.a_pc[s]) ++nA_pc; else ++nA_fixed;
    p.b_idx[s] = p.b_pc[s] ? p.nbpc++ : p.nbsh++;
  }
  for (int s = 0; s < a.nseg; ++s) {
    PbGemmSeg sg = a.seg[s];
    int nbA = a.nb, nbB = a.nb;
    if (!p.a_pc[s] && nslots > 1) { sg.sAb = a.p_stride / 2; nbA = nslots; }     // primal operand: batch axis = problem slot
    if (!p.b_pc[s] && nslots > 1) { sg.sBb = a.p_stride / 2; nbB = nslots; }
    uint32_t ab = 0, bb = 0, t = 0;
    int hm, bm;
    const int el0 = sp0 / 2, el1 = sp1 / 2;
    if (const char* e = pbgemm::encode_plainx(&p.mapA[s][0], sg.A, 1, a.Mr, a.d, sg.lda, sg.sAh, a.nh, sg.sAb, nbA, el0, TM, sp0, &hm, &bm, &t)) return e;
    ab += t; p.a_hmul[s] = hm; p.a_bmul[s] = p.a_pc[s] ? bm : 0; if (!p.a_pc[s]) p.a_smul[s] = bm;
    if (const char* e = pbgemm::encode_plainx(&p.mapB[s][0], sg.B, 1, a.Nc, a.d, sg.ldb, sg.sBh, a.nh, sg.sBb, nbB, el0, brows, sp0, &hm, &bm, &t)) return e;
    bb += t; p.b_hmul[s] = hm; p.b_bmul[s] = p.b_pc[s] ? bm : 0; if (!p.b_pc[s]) p.b_smul[s] = bm;
    if (sp1) {
      if (const char* e = pbgemm::encode_plainx(&p.mapA[s][1], sg.A, 1, a.Mr, a.d, sg.lda, sg.sAh, a.nh, sg.sAb, nbA, el1, TM, sp1, &hm, &bm, &t)) return e;
      ab += t;
      if (const char* e = pbgemm::encode_plainx(&p.mapB[s][1], sg.B, 1, a.Nc, a.d, sg.ldb, sg.sBh, a.nh, sg.sBb, nbB, el1, brows, sp1, &hm, &bm, &t)) return e;
      bb += t;
    }
    p.a_bytes[s] = ab;
    if (p.b_pc[s]) pc_b += bb; else sh_b += bb;
  }
  {
    // C1: [nh][d][ldc], K-major over the score columns; box = [64 columns = 128 bytes] x [d rows]
    uint64_t dims[4] = {uint64_t(a.Nc), uint64_t(a.d), uint64_t(a.nh), uint64_t(nslots)};
    uint64_t stb[3] = {uint64_t(a.ldc) * 2, uint64_t(a.sCh) * 2, nslots > 1 ? uint64_t(a.p_stride) : uint64_t(a.sCh) * 2 * a.nh};
    for (int rk = 0; rk < (pair ? 2 : 1); ++rk) {                      // rank rk holds rows rk * clr .. of the tile
      const int rows = std::min(p.clr, a.d - rk * p.clr);
      uint32_t box[4] = {uint32_t(TN), uint32_t(rows), 1, 1};
      if (const char* e = pbgemm::encode4x(&p.mapC[rk], a.C1, 1, dims, stb, box, 128)) return e;
      p.sh_bytes[rk] = uint32_t(rows) * 128;
    }
  }
  if (a.C2) {
    uint64_t dims[4] = {uint64_t(a.Nc), uint64_t(a.d), uint64_t(a.nh), uint64_t(a.nb)};
    uint64_t stb[3] = {uint64_t(a.ldc2) * 2, uint64_t(a.nh > 1 ? a.sC2h : a.ldc2) * 2, uint64_t(a.nb > 1 ? a.sC2b : a.ldc2) * 2};
    for (int rk = 0; rk < (pair ? 2 : 1); ++rk) {
      const int rows = std::min(p.clr, a.d - rk * p.clr);
      uint32_t box[4] = {uint32_t(TN), uint32_t(rows), 1, 1};
      if (const char* e = pbgemm::encode4x(&p.mapC2[rk], a.C2, 1, dims, stb, box, 128)) return e;
      p.pc_bytes[rk] = uint32_t(rows) * 128;
    }
  }
  {
    int hm, bm; uint32_t pb;
    static const bool p256 = getenv("PB_ATTN_P256") != nullptr;
    pbgemm::g_tmap_promo256 = p256 ? 1 : 0;
    struct Reset { ~Reset() { pbgemm::g_tmap_promo256 = 0; } } reset;
    if (const char* e = pbgemm::encode_plainx(&p.mapP, a.Pm, 1, a.Mr, a.Nc, a.ldp, a.sPh, a.nh, nslots > 1 ? a.p_stride / 2 : 0, nslots,
                                              TN, TM, 128, &hm, &bm, &pb)) return e;
    sh_b += pb;
  }
  for (int rk = 0; rk < 2; ++rk) { p.sh_bytes[rk] += sh_b; p.pc_bytes[rk] += pc_b; }
  p.pc_cnt = (p.nbpc > 0 ? 1 : 0) + (a.C2 ? 1 : 0);
  const bool has_pc = p.pc_cnt > 0;

  // column groups and ring depths from the shared-memory and TMEM budgets
  const int budget = 227 * 1024;
  const int fixed = 1024 /* alignment */ + 8 * 64 * 4 + 512 /* barriers */;
  auto smem_for = [&](int kc, int nsh, int npc, int nt) {
    return fixed + (nA_fixed + nA_pc * kc) * a_tile + nsh * (p.nbsh * b_tile + p.ctile + PT_BYTES) +
           npc * (p.nbpc * b_tile + (a.C2 ? p.ctile : 0)) + nt * PT_BYTES;
  };
  int ngrp = 0, kc = 0, ns = 0, nsh = 2, npc = 0, nt = 2;
  for (int g = 1; g <= k_slot; ++g) {
    const int k = (k_slot + g - 1) / g;
    if (k > MAX_KC || k > kc_cap) continue;
    // S, T and per-column rings are walked by two parity classes of warps (even / odd substeps): their depths must be EVEN, so
    // that a stage always belongs to one class -- with an odd depth a warp revisits a stage two laps later and a one-bit
    // phase parity cannot tell "two releases ago" from "now" (dead-lock seen at NS = 3; scripts/sim_attn16_protocol.py)
    static const bool ns4 = getenv("PB_ATTN_NS4") != nullptr;
    int s = std::min(MAX_NS, (TMEM_COLS - k * p.accw_tot) / TN) & ~1;
    if (s < (ns4 ? 4 : 2)) continue;
    static const int pc0_env = getenv("PB_ATTN_PC0") ? atoi(getenv("PB_ATTN_PC0")) : 4;      // A/B switch: per-column ring depth the search starts from
    const int pc0 = has_pc ? std::max(2, pc0_env & ~1) : 0;
    // the accumulate warp releases a shared stage LAG = 2 substeps after the P . C2 product of the same substep, and that
    // product may already need the stage NSH steps on: kc_min * (NSH - 1) >= LAG + 1 or the ring dead-locks
    const int kmin = k_slot / g;
    const int sh0 = a.C2 ? 1 + (3 + kmin - 1) / kmin : 2;
    if (sh0 > MAX_NSH || smem_for(k, sh0, pc0, 2) > budget) continue;
    ngrp = g; kc = k; ns = s; npc = pc0; nsh = sh0;
    break;
  }
  if (!ngrp) return nullptr;                                             // does not fit: the per-column kernel takes it
  for (bool grew = true; grew;) {                                        // spend what is left on deeper rings, round-robin
    grew = false;
    if (nt + 2 <= MAX_NT && smem_for(kc, nsh, npc, nt + 2) <= budget) { nt += 2; grew = true; }
    if (has_pc && npc + 2 <= MAX_NPC && smem_for(kc, nsh, npc + 2, nt) <= budget) { npc += 2; grew = true; }
    if (nsh < 3 && smem_for(kc, nsh + 1, npc, nt) <= budget) { ++nsh; grew = true; }
  }
  p.ngrp = ngrp; p.kc_base = k_slot / ngrp; p.kc_rem = k_slot % ngrp; p.kc_max = kc;
  p.ns = ns; p.nsh = nsh; p.npc = npc; p.nt = nt;
  int t0 = 0;
  for (int s = 0; s < a.nseg; ++s) { p.a_tile0[s] = t0; t0 += p.a_pc[s] ? kc : 1; }
  p.nA = t0;
  const int smem = smem_for(kc, nsh, npc, nt);
  const long gy = (long)a.nh * nslots * ngrp;
  if (gy > 65535) return "attn_lin: batch too large";
  dim3 grid((a.Mr + TM - 1) / TM, (unsigned)gy);

  static const bool tracing = getenv("PB_ATTN_TRACE") != nullptr;
  if (tracing) {
    if (!g_trace) cudaMalloc(&g_trace, sizeof(long long) * 512 * 16);
    cudaMemsetAsync(g_trace, 0, sizeof(long long) * 512 * 16, st);
    p.trace = g_trace;
    fprintf(stderr, "pb_attn16: pair %d role %d kc %d ngrp %d ns %d nsh %d npc %d nt %d smem %d\n", int(pair), role, kc, ngrp, ns, nsh, npc, nt, smem);
  }
  void (*kern)(Params) = nullptr;
#define PB_A16_ROLE2(KCFG_, PAIR_)                                                                            \
  kern = role == 0 ? attn16_kernel<2, KCFG_, 1, 0, PAIR_> : role == 1 ? attn16_kernel<1, KCFG_, 0, 0, PAIR_> \
       : role == 2 ? attn16_kernel<1, KCFG_, 0, 1, PAIR_> : attn16_kernel<1, KCFG_, 2, 2, PAIR_>;
#define PB_A16_ROLE(KCFG_) if (pair) { PB_A16_ROLE2(KCFG_, true) } else { PB_A16_ROLE2(KCFG_, false) }
  if (kcfg == 0) { PB_A16_ROLE(0) } else if (kcfg == 1) { PB_A16_ROLE(1) } else if (kcfg == 2) { PB_A16_ROLE(2) } else { PB_A16_ROLE(3) }
#undef PB_A16_ROLE
#undef PB_A16_ROLE2
  if (const char* err = pbhost::optin_smem(kern, 227 * 1024)) return err;
  cudaError_t e;
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
  } else {
    kern<<<grid, NTHREADS, smem, st>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cudaGetErrorString(e);
  *handled = true;
  return nullptr;
}

}  // namespace pbattn16
