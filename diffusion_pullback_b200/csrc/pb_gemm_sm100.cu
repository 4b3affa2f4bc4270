// tcgen05 / TMA GEMM (TF32 or fp16 operands, fp32 accumulation) for sm_100a: the contraction engine behind every
// conv / linear / attention product of the pullback hot path (primal, JVP and VJP passes).
//
// Persistent kernel, one CTA per SM, each CTA walks a static list of work items (128 x BN output tile, K split):
//   warp 0      : TMA producer  (cp.async.bulk.tensor 4D, 128B-swizzled K-major tiles, OOB zero fill
//                 supplies conv padding, K/M/N tails and attention-head tails)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (the whole warp walks the loop, one elected lane issues; kind::tf32
//                 or kind::f16, fp32 accumulators in TMEM, two accumulator stages so the epilogue of item i overlaps the
//                 main loop of item i + 1)
//   warps 2..5  : epilogue group 0: tcgen05.ld 32x32b -> alpha / bias / residual / RNA rounding -> 128B-swizzled smem staging
//                 -> TMA tensor store (the residual tile arrives by TMA into the same staging buffer, 3 chunks ahead)
//   warps 6..9  : epilogue group 1 (fp16-output kernels): the groups take alternate 32-column chunks of every tile
// CTA-pair variant (PAIR, fp16 operands): the two CTAs of a cluster run tiles 2j, 2j + 1 as ONE 256-row
// tcgen05.mma.cta_group::2 issued by the leader; each CTA stages its 128 rows of A and half of the B tile (see the kernel).
// GEGLU variant of the pair kernel: the GEGLU linearisation of the ff1 tangent in the epilogue (PbGemm::gg).
// smem: STAGES x (A 16 KB + B BN*128 B) operand ring with full/empty mbarriers (MMA completion by tcgen05.commit)
// + 4 x 16 KB staging buffers.
// Scheduling: output tiles are dealt round-robin to the CTAs.  The tiles of the last, partial wave (all tiles when there
// are fewer tiles than SMs: small-M weight-streaming layers) are cut along K into `splits` items each so that the wave
// fills the machine (split count from a small cost model: rounds x k-blocks + reduce); such items store their raw partial
// tile to scratch through the same TMA epilogue and
// splitk_reduce_k sums the partials in split order (deterministic) and applies alpha / bias / residual / rounding.
// Operand / output types (template parameters): fp32 operands run as kind::tf32 (32-element k-blocks), fp16 operands as
// kind::f16 (64-element k-blocks, half the L2->smem bytes per flop); accumulation is fp32 in TMEM either way and the
// output / residual is fp32 or fp16 (32-column chunks: 128-byte or 64-byte staging rows).
// See pb_gemm.h for the operation this implements.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <type_traits>

#include "pb_gemm.h"
#include "pb_host_util.h"
#include "pb_tc.cuh"

namespace pbgemm {

constexpr int BM = 128;
constexpr int BK = 32;                 // k-block of the fp32 / TF32 path: 32 fp32 = 128 bytes = one swizzle row
constexpr int BK16 = 64;               // k-block of the fp16 path: 64 halves = 128 bytes
constexpr int NTHREADS = 320;              // warps: 0 TMA producer, 1 MMA issuer, 2-5 / 6-9 epilogue groups
constexpr int EPI_W = 32;              // epilogue chunk: 32 columns = one 128-byte staging row
constexpr int NBUF = 4;                // staging buffers
constexpr int EPI_BYTES = BM * EPI_W * 4;
constexpr int ACC_STRIDE = 256;        // TMEM columns between the two accumulator stages

struct alignas(64) Params {
  CUtensorMap mapA[2];
  CUtensorMap mapB[2];
  CUtensorMap mapD, mapR;
  int kblocks[2];                      // ceil(K_seg / 32)
  uint32_t tx_bytes[2];                // bytes one stage's A+B boxes deliver (boxes are clamped to the tensor)
  int a_bmul[2], a_hmul[2], b_bmul[2], b_hmul[2], d_bmul, d_hmul;
  int a_bdiv[2], b_bdiv[2];            // problem slots: a primal operand's batch coordinate is bat_b / bdiv (0: bat_b * bmul)
  int nseg, taps, conv_ctot;
  int M, N, nb, nh;
  int conv, H, W, bw, bh, bb, tiles_w, tiles_h;
  CUtensorMap mapW;                    // split-K scratch as a [tile][split][128][BN] tensor
  int raster_b, mt, nt, items;
  int pair;                            // CTA-pair kernel: tiles 2j and 2j + 1 are the two 128-row halves of pair-item j
  int full_tiles;                      // tiles [0, full_tiles) are whole items; every later tile is `splits` partial items
  int splits, kb_per_split;            // partial item s covers flat k-blocks [s*kb_per_split, (s+1)*kb_per_split)
  float* ws;
  int epi_tma; uint32_t r_bytes;       // epilogue through smem staging + TMA (else guarded direct stores)
  void* D; const void* R; const float* bias;
  long ldd, sDb, sDh, ldr, sRb, sRh;
  float alpha, beta;
  int round_tf32;
  const float* gg; int gg_F; unsigned gg_rows_p, gg_k_slot; long gg_p_stride;   // GEGLU tangent epilogue (PbGemm::gg)
  long long* trace;                    // PB_GEMM_TRACE: per-item clocks of CTA 0 ([item][8]: producer begin / end, MMA begin / end, epilogue begin / end)
};

using namespace pbtc;

template <int BN, int STAGES, bool PAIR = false>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 4;   // a CTA of a pair stages half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STAGING_OFF + NBUF * EPI_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;         // + barriers + 1024B alignment slack
};

struct Tile {
  int n0, m0, bat_b, bat_h;            // plain mode
  int cx0, cy0, cb0;                   // conv mode tile origin
  int split, tile_id, partial;
};

template <int BN>
__device__ __forceinline__ Tile decode_tile(const Params& p, int tile) {
  Tile t;
  t.m0 = t.bat_b = t.bat_h = t.cx0 = t.cy0 = t.cb0 = t.split = t.partial = 0;
  t.tile_id = tile;
  int v = tile;
  if (p.pair) {
    // CTA-pair kernel (conv mode, or plain mode with nb = nh = 1): tiles 2j and 2j + 1 share the column tile and are row
    // tiles 2m', 2m' + 1; a row tile past the end (odd count) lies outside the tensor: zero-filled loads, clipped stores
    const int rank = v & 1; v >>= 1;
    t.n0 = (v % p.nt) * BN;
    int m = 2 * (v / p.nt) + rank;
    if (p.conv) {
      const int tw = m % p.tiles_w; m /= p.tiles_w;
      const int th = m % p.tiles_h; m /= p.tiles_h;
      t.cx0 = tw * p.bw; t.cy0 = th * p.bh; t.cb0 = m * p.bb;
    } else {
      t.m0 = m * BM;
    }
    return t;
  }
  if (p.raster_b) {
    // tangent index fastest: the nb tiles that read the same broadcast A tile (attention probabilities) run at the same
    // time on neighbouring SMs, so A comes from HBM once and from L2 nb - 1 times
    t.bat_b = v % p.nb; v /= p.nb;
    t.n0 = (v % p.nt) * BN; v /= p.nt;
    t.m0 = (v % p.mt) * BM;
    t.bat_h = v / p.mt;
    return t;
  }
  t.n0 = (v % p.nt) * BN; v /= p.nt;
  int m = v % p.mt; v /= p.mt;
  if (p.conv) {
    const int tw = m % p.tiles_w; m /= p.tiles_w;
    const int th = m % p.tiles_h; m /= p.tiles_h;
    t.cx0 = tw * p.bw; t.cy0 = th * p.bh; t.cb0 = m * p.bb;
  } else {
    t.m0 = m * BM;
    t.bat_h = v % p.nh;
    t.bat_b = v / p.nh;
  }
  return t;
}

template <int BN>
__device__ __forceinline__ Tile decode_item(const Params& p, int item) {
  if (item < p.full_tiles) return decode_tile<BN>(p, item);
  int j = item - p.full_tiles;
  Tile t;
  if (p.pair) {                        // items 2i, 2i + 1: the same split of the two tiles of a pair
    const int rank = j & 1; j >>= 1;
    t = decode_tile<BN>(p, p.full_tiles + 2 * (j / p.splits) + rank);
  } else {
    t = decode_tile<BN>(p, p.full_tiles + j / p.splits);
  }
  t.split = j % p.splits;
  t.partial = 1;
  return t;
}

// row r of a tile -> element offsets of D / R (false: the row lies outside the tensor)
__device__ __forceinline__ bool tile_row(const Params& p, const Tile& t, int r, long* d_off, long* r_off) {
  if (p.conv) {
    const int w = r % p.bw;
    const int hh = (r / p.bw) % p.bh;
    const int bb = r / (p.bw * p.bh);
    const int x = t.cx0 + w, y = t.cy0 + hh, b = t.cb0 + bb;
    const long pix = (static_cast<long>(b) * p.H + y) * p.W + x;
    *d_off = pix * p.ldd;
    *r_off = pix * p.ldr;
    return (x < p.W) && (y < p.H) && (b < p.nb);
  }
  *d_off = t.bat_b * p.sDb + t.bat_h * p.sDh + static_cast<long>(t.m0 + r) * p.ldd;
  *r_off = t.bat_b * p.sRb + t.bat_h * p.sRh + static_cast<long>(t.m0 + r) * p.ldr;
  return (t.m0 + r) < p.M;
}

// 32 accumulator columns of tile row r -> staging row (swizzled 16-byte chunks), fused with alpha / bias / residual
// (the residual chunk is already in the staging buffer) / rounding
template <bool OUT16>
__device__ __forceinline__ void stage_chunk(uint8_t* buf, int r, const uint32_t (&v)[32], float alpha, float beta,
                                            const float* bias, int nc, int N, bool full, bool has_r, int rnd) {
  if constexpr (!OUT16) {
    uint8_t* row = buf + r * 128;
    const uint32_t swz = uint32_t(r & 7);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4* sp = reinterpret_cast<float4*>(row + ((uint32_t(j >> 2) ^ swz) << 4));
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = alpha * __uint_as_float(v[j + e]);
      if (bias) {
        if (full) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + nc + j);
          o[0] += bv.x; o[1] += bv.y; o[2] += bv.z; o[3] += bv.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (nc + j + e < N) o[e] += bias[nc + j + e];
        }
      }
      if (has_r) {
        const float4 rv = *sp;
        o[0] += beta * rv.x; o[1] += beta * rv.y; o[2] += beta * rv.z; o[3] += beta * rv.w;
      }
      if (rnd) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rna_tf32(o[e]);
      }
      *sp = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
    uint8_t* row = buf + r * 64;
    const uint32_t swz = uint32_t((r >> 1) & 3);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4* sp = reinterpret_cast<uint4*>(row + ((uint32_t(j >> 3) ^ swz) << 4));
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = alpha * __uint_as_float(v[j + e]);
      if (bias) {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (full || nc + j + e < N) o[e] += bias[nc + j + e];
      }
      if (has_r) {
        const uint4 rv = *sp;
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(rh[e]);
          o[2 * e] += beta * f.x; o[2 * e + 1] += beta * f.y;
        }
      }
      uint4 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(o[2 * e], o[2 * e + 1]);
      *sp = ov;
    }
  }
}

// PAIR: the CTA-pair variant (fp16 operands).  The two CTAs of a cluster, on the two SMs of a TPC, own row tiles 2m' and
// 2m' + 1 of the same column tile and run them as ONE 256 x BN tcgen05.mma.cta_group::2 issued by the leader (cluster
// rank 0): each CTA stages its own 128 rows of A and HALF of the B tile, so a k-block costs a CTA 16 KB + BN * 64 B of
// shared-memory writes and as many reads instead of 16 KB + BN * 128 B -- the one-CTA 128 x 160 tile needs 230 B per clock
// of shared-memory bandwidth (36 KB written by TMA + 36 KB read by the tensor core per 320-clock k-block) against the 128 B
// per clock an SM has, which is what held the big convolutions at half the tensor peak.  Protocol: both producers load
// into their own stage s and signal the LEADER's full barrier (one expect_tx of both CTAs' bytes); the leader's commits
// arrive on the empty / accumulator-full barriers of BOTH CTAs (multicast); the epilogue warps of both CTAs release an
// accumulator stage on the leader's barrier (8 arrivals).  Epilogue, staging, residual and split-K paths are per CTA and
// unchanged.
// GEGLU (pair kernel, BN = 256, fp16 out): the tangent of ff1 never reaches HBM.  The weight rows are interleaved in blocks of 64
// ([32 a rows | the 32 matching gate rows]), so accumulator chunk 2i holds da and chunk 2i + 1 holds dg of output columns
// n0 / 2 + 32 i ..; an epilogue group combines its chunk PAIRS with the cached factors, out = da * gelu(g) + dg * a gelu'(g)
// (two 128-byte global reads per row and pair), and stores 32 output columns through the usual staging + TMA path.
template <int BN, int STAGES, bool AB16, bool D16, bool PAIR = false, bool GEGLU = false>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const __grid_constant__ Params p) {
  static_assert(!PAIR || AB16, "the CTA-pair variant takes fp16 operands");
  static_assert(!GEGLU || (PAIR && D16 && BN % 64 == 0), "the GEGLU epilogue belongs to the fp16 pair kernel");
  using OutT = typename std::conditional<D16, __half, float>::type;
  constexpr int KB_ELEMS = AB16 ? BK16 : BK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = Smem<BN, STAGES, PAIR>;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  uint8_t* staging = smem + S::STAGING_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* r_bar = tmem_empty_bar + 2;              // [2 * NBUF]: residual barriers of the two epilogue groups
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(r_bar + 2 * NBUF);
  constexpr int NG = D16 ? 2 : 1;                    // epilogue groups (see the epilogue)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // descriptor fetches overlap the prologue: operands by the producer lane, output / residual / scratch by the store thread
  if (threadIdx.x == 0) {
    prefetch_tmap(&p.mapA[0]); prefetch_tmap(&p.mapB[0]);
    if (p.nseg > 1) { prefetch_tmap(&p.mapA[1]); prefetch_tmap(&p.mapB[1]); }
  } else if (threadIdx.x == 64) {
    if (p.epi_tma) prefetch_tmap(&p.mapD);
    if (p.R != nullptr && p.epi_tma) prefetch_tmap(&p.mapR);
    if (p.splits > 1) prefetch_tmap(&p.mapW);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], (PAIR ? 8 : 4) * NG); }
    for (int i = 0; i < 2 * NBUF; ++i) mbar_init(&r_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {        // collective over the pair: warp 1 of both CTAs, the same shared-memory offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_base_smem)),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_base_smem)),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int kb_tap = p.kblocks[0] + (p.nseg > 1 ? p.kblocks[1] : 0);
  const int total_kb = p.taps * kb_tap;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // The whole warp walks the loop and one ELECTED lane issues: inside an `if (lane == 0)` region the compiler cannot prove
    // the TMA operands warp-uniform and wraps every UTMALDG in an ELECT / R2UR.BROADCAST waterfall (47 broadcasts in this
    // kernel's SASS), which made a k-block's two loads cost more issue time than its MMAs take to execute.
    {
      int stage = 0; uint32_t phase = 0;
      const uint32_t full_leader = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0u;   // the leader's full barriers
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Tile t = decode_item<BN>(p, item);
        const int kb_begin = t.partial ? t.split * p.kb_per_split : 0;
        const int kb_end = t.partial ? min(total_kb, kb_begin + p.kb_per_split) : total_kb;
        const int tli = (item - blockIdx.x) / gridDim.x;
        if (p.trace && blockIdx.x == 0 && lane == 0 && tli < 64) p.trace[tli * 8 + 0] = clock64();
        int tap = kb_begin / kb_tap;
        int kbt = kb_begin - tap * kb_tap;                     // k-block inside the tap, advanced without division
        for (int it = kb_begin; it < kb_end; ++it) {
          const int s = kbt < p.kblocks[0] ? 0 : 1;
          const int kb = s ? kbt - p.kblocks[0] : kbt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if constexpr (PAIR) {
            const uint32_t fb = full_leader + uint32_t(stage) * 8u;
            const int nb0 = t.n0 + int(cta_rank) * (BN / 2);
            if (elect_one()) {
              // the leader expects both CTAs' bytes; a peer load that lands first only drives the count negative
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * p.tx_bytes[0]);
              if (p.conv) {
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                tma_load_4d_2sm(sa, &p.mapA[0], fb, kb * KB_ELEMS, t.cx0 + dx, t.cy0 + dy, t.cb0);
                tma_load_4d_2sm(sb, &p.mapB[0], fb, kb * KB_ELEMS, tap, nb0, 0);
              } else {
                tma_load_4d_2sm(sa, &p.mapA[0], fb, kb * KB_ELEMS, t.m0, 0, 0);
                tma_load_4d_2sm(sb, &p.mapB[0], fb, kb * KB_ELEMS, nb0, 0, 0);
              }
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes[s]);
            if (p.conv) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              tma_load_4d(sa, &p.mapA[s], &full_bar[stage], kb * KB_ELEMS, t.cx0 + dx, t.cy0 + dy, t.cb0);
              // filter as a (channel, tap, out-channel) tensor: a k-block running past the channel count is zero-filled
              tma_load_4d(sb, &p.mapB[s], &full_bar[stage], kb * KB_ELEMS, tap, t.n0, 0);
            } else {
              const int ca = p.a_bdiv[s] ? t.bat_b / p.a_bdiv[s] : t.bat_b * p.a_bmul[s];
              const int cb = p.b_bdiv[s] ? t.bat_b / p.b_bdiv[s] : t.bat_b * p.b_bmul[s];
              tma_load_4d(sa, &p.mapA[s], &full_bar[stage], kb * KB_ELEMS, t.m0, t.bat_h * p.a_hmul[s], ca);
              tma_load_4d(sb, &p.mapB[s], &full_bar[stage], kb * KB_ELEMS, t.n0, t.bat_h * p.b_hmul[s], cb);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++kbt == kb_tap) { kbt = 0; ++tap; }
        }
        if (p.trace && blockIdx.x == 0 && lane == 0 && tli < 64) p.trace[tli * 8 + 1] = clock64();
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the loop; one elected lane issues (see elect_one); of a CTA pair only the leader's
    {
      // instruction descriptor: fp32 accumulate; A/B format tf32 (2) or f16 (0); N >> 3, M >> 4 (256 rows over a pair)
      constexpr uint32_t idesc = (1u << 4) | (AB16 ? 0u : (2u << 7) | (2u << 10)) | (uint32_t(BN >> 3) << 17) |
                                 (uint32_t((PAIR ? 2 * BM : BM) >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0; uint32_t phase = 0;
      int li = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++li) {
        const bool partial = item >= p.full_tiles;
        const int kb_begin = partial ? (((item - p.full_tiles) >> (PAIR ? 1 : 0)) % p.splits) * p.kb_per_split : 0;
        const int kb_end = partial ? min(total_kb, kb_begin + p.kb_per_split) : total_kb;
        const int as = li & 1;
        mbar_wait(&tmem_empty_bar[as], ((li >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_u + uint32_t(as * ACC_STRIDE);
        if (p.trace && blockIdx.x == 0 && lane == 0 && li < 64) p.trace[li * 8 + 2] = clock64();
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u + stage * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + S::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 8 tf32 / 16 halves = 32 bytes along K inside the swizzle row: +2 in the 16-byte address field
              if constexpr (PAIR)
                mma_f16_2sm(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, ((kb - kb_begin) | k) ? 1u : 0u);
              else if constexpr (AB16)
                mma_f16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, ((kb - kb_begin) | k) ? 1u : 0u);
              else
                mma_tf32(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, ((kb - kb_begin) | k) ? 1u : 0u);
            }
            if constexpr (PAIR) tcgen05_commit_2sm(&empty_bar[stage]); else tcgen05_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if constexpr (PAIR) tcgen05_commit_2sm(&tmem_full_bar[as]); else tcgen05_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (p.trace && blockIdx.x == 0 && lane == 0 && li < 64) p.trace[li * 8 + 3] = clock64();
      }
    }
  } else if (warp >= 2 && ((warp - 2) >> 2) < NG) {
    // =========================== epilogue ===========================
    // fp16-output kernels run TWO epilogue groups of four warps (warps 2-5 and 6-9; a warp may touch the TMEM lane quarter
    // warp % 4): group g takes the 32-column chunks c = g, g + 2, ... of every tile.  One group needs ~1300 clocks per chunk
    // (TMEM load -> convert -> staging -> proxy fence -> group barrier -> TMA store), 6400 per 128 x 160 tile, which is what
    // paced every GEMM with fewer than ~12 k-blocks per tile (per-item clocks of PB_GEMM_TRACE).  Each group owns half of the
    // staging area, its own named barrier, residual barriers and bulk async-groups.
    const int eg = (warp - 2) >> 2;           // epilogue group
    const int q = warp & 3;                   // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;              // tile row
    // the group's first warp drives its staging TMA traffic through ONE ELECTED lane (elect.sync names the same leader for the
    // same member mask every time, so the bulk async-groups are committed and waited on by one thread); an
    // `if (threadIdx.x == 64)` region made every UTMASTG / UTMALDG of the epilogue an ELECT / R2UR.BROADCAST waterfall
    const bool w0 = ((warp - 2) & 3) == 0;
    const int bar_id = 1 + eg;
    const uint32_t tmem_empty_leader = PAIR ? mapa_u32(smem_u32(tmem_empty_bar), 0) : 0u;
    // staging geometry: fp32-output kernels 4 x 16 KB (one group); fp16-output kernels 4 x 8 KB per group for output tiles
    // (64-byte rows) and 2 x 16 KB per group for the fp32 partial tiles of split items
    uint8_t* gstaging = staging + eg * (NBUF * EPI_BYTES / NG);
    uint64_t* gr_bar = r_bar + eg * NBUF;
    uint32_t gch = 0;                         // chunks this group pushed through its staging ring so far
    uint32_t r_par = 0;                       // bit b: parity of the next residual load into staging buffer b
    int cur_partial = 0;
    constexpr int TAILQ = D16 ? 7 : 3;        // TMA stores clip the inner dimension in 16-byte units
    int li = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++li) {
      const Tile t = decode_item<BN>(p, item);
      const int as = li & 1;
      const int nchunks = min(BN / EPI_W, (p.N - t.n0 + EPI_W - 1) / EPI_W);
      // a ragged last chunk (N not a multiple of 16 bytes) takes the guarded direct path
      const int ntma = t.partial ? nchunks
                                 : !p.epi_tma ? 0 : ((p.N & TAILQ) && t.n0 + nchunks * EPI_W > p.N) ? nchunks - 1 : nchunks;
      const uint32_t acc = tmem_base + uint32_t(as * ACC_STRIDE) + (uint32_t(q * 32) << 16);
      const float alpha = t.partial ? 1.f : p.alpha, beta = p.beta;
      const bool has_r = p.R != nullptr && !t.partial;
      const float* bias = t.partial ? nullptr : p.bias;
      const int rnd = t.partial ? 0 : p.round_tf32;
      const CUtensorMap* dmap = t.partial ? &p.mapW : &p.mapD;
      int c1 = p.conv ? t.cx0 : t.m0, c2 = p.conv ? t.cy0 : t.bat_h * p.d_hmul, c3 = p.conv ? t.cb0 : t.bat_b * p.d_bmul;
      const int rc1 = c1, rc2 = c2, rc3 = c3;
      int dcol0 = t.n0;
      if (t.partial) { c1 = ((t.tile_id - p.full_tiles) * p.splits + t.split) * BM; c2 = c3 = 0; dcol0 = 0; }
      if (D16 && t.partial && !cur_partial) {
        // the 16 KB partial-tile buffers overlay the 8 KB output buffers: the group's pending stores must have been read
        cur_partial = 1;
        if (w0) {
          if (elect_one()) bulk_wait_read<0>();
          __syncwarp();
        }
        named_bar_sync(bar_id, 128);
      }
      const uint32_t nbuf = (D16 && t.partial) ? 2u : uint32_t(NBUF);
      const uint32_t bufsz = (D16 && !t.partial) ? uint32_t(EPI_BYTES / 2) : uint32_t(EPI_BYTES);
      auto prefetch_r = [&](int c, uint32_t g) {   // residual chunk c of this tile -> staging buffer g % NBUF of the group
        const uint32_t b = g % NBUF;
        mbar_arrive_expect_tx(&gr_bar[b], p.r_bytes);
        tma_load_4d(gstaging + b * bufsz, &p.mapR, &gr_bar[b], t.n0 + c * EPI_W, rc1, rc2, rc3);
      };
      long d_off, r_off;
      const bool row_ok = tile_row(p, t, r, &d_off, &r_off);

      if (w0 && has_r) {
        if (elect_one())
          for (int j = 0; j < 3 && eg + NG * j < ntma; ++j) prefetch_r(eg + NG * j, gch + j);
        __syncwarp();
      }
      // GEGLU epilogue: factor-row pointers of the four rows whose pieces this lane fetches (row 8 it + lane / 4 of the warp's 32), and
      // the pieces of the group's first two half-steps issued ahead of the accumulator wait
      const float* gptr[4] = {nullptr, nullptr, nullptr, nullptr}; bool gok[4] = {false, false, false, false};
      float4 gpre1[2][4], gpre2[2][4];
      if constexpr (GEGLU) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const unsigned m = unsigned(t.m0 + q * 32 + it * 8 + (lane >> 2));
          const unsigned img = m / p.gg_rows_p;
          gptr[it] = p.gg + (long)(img / p.gg_k_slot) * p.gg_p_stride + (long)(m - img * p.gg_rows_p) * 2 * p.gg_F;
          gok[it] = int(m) < p.M;
        }
#pragma unroll
        for (int s0 = 0; s0 < 2; ++s0) {
          const int fn = (t.n0 >> 1) + 32 * eg + s0 * 16 + 4 * (lane & 3);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            gpre1[s0][it] = gok[it] ? __ldg(reinterpret_cast<const float4*>(gptr[it] + fn)) : make_float4(0.f, 0.f, 0.f, 0.f);
            gpre2[s0][it] = gok[it] ? __ldg(reinterpret_cast<const float4*>(gptr[it] + p.gg_F + fn)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && li < 64) p.trace[li * 8 + 6] = clock64();
      mbar_wait(&tmem_full_bar[as], (li >> 1) & 1);
      tcgen05_fence_after();
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && li < 64) p.trace[li * 8 + 4] = clock64();
      if (eg >= nchunks) {                        // no chunk of this tile for the group: release the accumulator at once
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + uint32_t(as) * 8u);
          else mbar_arrive(&tmem_empty_bar[as]);
        }
      }

      if constexpr (GEGLU) {
        // half-steps s = (pair, 16-column half) of this group.  The factors of a half-step (64 bytes per row and operand) are
        // loaded COALESCED -- lane l fetches 16-byte piece l % 4 of row 8 it + l / 4 of its warp's 32 rows, four trips per operand --
        // TWO half-steps ahead into registers (the first two were issued before the accumulator wait; carrying them over from the
        // CTA's previous tile instead spilled 256 bytes per thread and was slower), pass through a per-warp
        // scratch tile in the staging area (swizzled like the output rows: conflict-free both ways) and come back as the
        // thread's own row.  One row per thread straight from global memory made every LDG.128 a 32-sector request with
        // ~4000 clocks of latency: 552 us for the 102400 x 2560 x 320 launch against 307 us of the two separate kernels.
        constexpr int NPAIR = BN / 64, NSTEP = 2 * (NPAIR / NG);
        static_assert(NPAIR % NG == 0, "pairs split evenly over the epilogue groups");
        uint8_t* scr1 = gstaging + 2 * (EPI_BYTES / 2);               // buffers 2, 3 of the group: factor scratch (G1, G2)
        uint8_t* scr2 = gstaging + 3 * (EPI_BYTES / 2);
        const uint32_t swz = uint32_t((r >> 1) & 3);
        float4 pf1[3][4], pf2[3][4];                                  // pieces in flight: [step % 3][trip]
#pragma unroll
        for (int it = 0; it < 4; ++it) { pf1[0][it] = gpre1[0][it]; pf2[0][it] = gpre2[0][it]; pf1[1][it] = gpre1[1][it]; pf2[1][it] = gpre2[1][it]; }
#pragma unroll
        for (int s = 0; s < NSTEP; ++s) {
          const int i = eg + NG * (s >> 1), hh = s & 1, cur = s % 3;
          const int f0 = (t.n0 >> 1) + 32 * i;                        // first output column of the pair
          if (s + 2 < NSTEP) {
            const int fn = (t.n0 >> 1) + 32 * (eg + NG * ((s + 2) >> 1)) + ((s + 2) & 1) * 16 + 4 * (lane & 3);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              pf1[(s + 2) % 3][it] = gok[it] ? __ldg(reinterpret_cast<const float4*>(gptr[it] + fn)) : make_float4(0.f, 0.f, 0.f, 0.f);
              pf2[(s + 2) % 3][it] = gok[it] ? __ldg(reinterpret_cast<const float4*>(gptr[it] + p.gg_F + fn)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          // pieces of this half-step -> scratch rows of this warp -> the thread's own row
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = q * 32 + it * 8 + (lane >> 2);
            const uint32_t off = uint32_t(row * 64) + ((uint32_t(lane & 3) ^ uint32_t((row >> 1) & 3)) << 4);
            *reinterpret_cast<float4*>(scr1 + off) = pf1[cur][it];
            *reinterpret_cast<float4*>(scr2 + off) = pf2[cur][it];
          }
          __syncwarp();
          float4 a4[4], c4[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            a4[qq] = *reinterpret_cast<const float4*>(scr1 + r * 64 + ((uint32_t(qq) ^ swz) << 4));
            c4[qq] = *reinterpret_cast<const float4*>(scr2 + r * 64 + ((uint32_t(qq) ^ swz) << 4));
          }
          __syncwarp();
          const uint32_t ob = gch & 1u;                               // output buffers 0, 1 of the group
          uint8_t* sbuf = gstaging + ob * uint32_t(EPI_BYTES / 2);
          uint8_t* srow = sbuf + r * 64;
          uint32_t va[16], vg[16];
          tmem_ld16(acc + uint32_t((2 * i) * EPI_W + hh * 16), va);
          tmem_ld16(acc + uint32_t((2 * i + 1) * EPI_W + hh * 16), vg);
          if (s == NSTEP - 1) {                                       // the group's last read of this accumulator stage
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + uint32_t(as) * 8u);
          }
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {                            // 8 columns = one 16-byte piece of the staging row
            const float4 a0 = a4[2 * qq], a1 = a4[2 * qq + 1], c0 = c4[2 * qq], c1 = c4[2 * qq + 1];
            const float w1[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float w2[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = fmaf(__uint_as_float(va[8 * qq + e]), w1[e], __uint_as_float(vg[8 * qq + e]) * w2[e]);
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(o[2 * e], o[2 * e + 1]);
            *reinterpret_cast<uint4*>(srow + ((uint32_t(hh * 2 + qq) ^ swz) << 4)) = ov;
          }
          if (hh == 1) {
            fence_proxy_async_smem();
            if (w0) {                                                 // two output buffers: the previous pair's store must have been read
              if (elect_one()) bulk_wait_read<0>();
              __syncwarp();
            }
            named_bar_sync(bar_id, 128);
            if (w0) {
              if (elect_one()) {
                tma_store_4d(&p.mapD, sbuf, f0, t.m0, 0, 0);
                bulk_commit();
              }
              __syncwarp();
            }
            ++gch;
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = eg; c < nchunks; c += NG) {
        const int nc = t.n0 + c * EPI_W;
        uint32_t v[32];
        tmem_ld32(acc + uint32_t(c * EPI_W), v);
        if (c + NG >= nchunks) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + uint32_t(as) * 8u);
            else mbar_arrive(&tmem_empty_bar[as]);
          }
        }
        const bool full = (nc + EPI_W <= p.N);
        if (c < ntma) {
          const uint32_t b = gch % nbuf;
          uint8_t* sbuf = gstaging + b * bufsz;
          if (has_r) { mbar_wait(&gr_bar[b], (r_par >> b) & 1u); r_par ^= 1u << b; }
          if (t.partial) stage_chunk<false>(sbuf, r, v, alpha, beta, bias, nc, p.N, full, has_r, rnd);
          else stage_chunk<D16>(sbuf, r, v, alpha, beta, bias, nc, p.N, full, has_r, rnd);
          fence_proxy_async_smem();
          // the NEXT chunk stages into the buffer of the store issued nbuf - 1 chunks before it: the store thread waits for
          // that store's reads BEFORE the group barrier, so that every warp leaving the barrier knows the buffer is free (a
          // wait after the barrier told only the store thread; with two 16 KB buffers the other warps overwrote a partial
          // tile the previous store was still reading).  With a residual the buffer's next user is the TMA residual load
          // issued below by the store thread itself, and the warps wait for its barrier.
          if (w0 && !has_r) {
            if (elect_one()) {
              if (nbuf == 2u) bulk_wait_read<0>(); else bulk_wait_read<NBUF - 2>();
            }
            __syncwarp();
          }
          named_bar_sync(bar_id, 128);
          if (w0) {
            if (elect_one()) {
              tma_store_4d(dmap, sbuf, dcol0 + c * EPI_W, c1, c2, c3);
              bulk_commit();
              // with a residual the buffer of the PREVIOUS store is refilled right here: only the newest store may be pending
              if (has_r) {
                bulk_wait_read<1>();
                if (c + 3 * NG < ntma) prefetch_r(c + 3 * NG, gch + 3);
              }
            }
            __syncwarp();
          }
          ++gch;
        } else if (row_ok) {
          OutT* dptr = static_cast<OutT*>(p.D) + d_off + nc;
          const OutT* rptr = has_r ? static_cast<const OutT*>(p.R) + r_off + nc : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nc + j < p.N) {
              float o = alpha * __uint_as_float(v[j]);
              if (bias) o += bias[nc + j];
              if (rptr) o += beta * static_cast<float>(rptr[j]);
              if (rnd && !D16) o = rna_tf32(o);
              dptr[j] = static_cast<OutT>(o);
            }
          }
        }
      }
    }
    // the staging buffers must have been read before the CTA exits; the global writes themselves complete with the kernel
    // (every consumer of D or of the split-K scratch is a later kernel in the stream)
    if (w0) {
      if (elect_one()) bulk_wait_read<0>();
      __syncwarp();
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // no CTA leaves while its peer may still signal its barriers / read its tiles
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Sums the `splits` partial tiles of every split tile in split order and applies the epilogue.  One thread per
// (tile row, 4 columns); partials are [tile][split][128][BN] fp32, L2-resident right after the GEMM.
template <int BN, bool D16>
__global__ void __launch_bounds__(256) splitk_reduce_k(const __grid_constant__ Params p, int ntail) {
  using OutT = typename std::conditional<D16, __half, float>::type;
  constexpr int QN = BN / 4;
  const long total = static_cast<long>(ntail) * BM * QN;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(i % QN);
    const int r = static_cast<int>((i / QN) % BM);
    const int tt = static_cast<int>(i / (QN * BM));
    const Tile t = decode_tile<BN>(p, p.full_tiles + tt);
    const int nc = t.n0 + cq * 4;
    long d_off, r_off;
    if (nc >= p.N || !tile_row(p, t, r, &d_off, &r_off)) continue;
    const float4* src = reinterpret_cast<const float4*>(p.ws + (static_cast<size_t>(tt) * p.splits * BM + r) * BN) + cq;
    float4 a = __ldcg(src);
    for (int s = 1; s < p.splits; ++s) {
      const float4 b = __ldcg(src + static_cast<size_t>(s) * BM * QN);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float o[4] = {a.x * p.alpha, a.y * p.alpha, a.z * p.alpha, a.w * p.alpha};
    OutT* dptr = static_cast<OutT*>(p.D) + d_off + nc;
    const OutT* rptr = p.R ? static_cast<const OutT*>(p.R) + r_off + nc : nullptr;
    const int nv = min(4, p.N - nc);
    for (int e = 0; e < nv; ++e) {
      if (p.bias) o[e] += p.bias[nc + e];
      if (rptr) o[e] += p.beta * static_cast<float>(rptr[e]);
      if (p.round_tf32 && !D16) o[e] = rna_tf32(o[e]);
    }
    if (nv == 4) {
      if constexpr (D16) {
        uint2 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
        oh[0] = __floats2half2_rn(o[0], o[1]); oh[1] = __floats2half2_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(dptr) = ov;
      } else {
        *reinterpret_cast<float4*>(dptr) = make_float4(o[0], o[1], o[2], o[3]);
      }
    } else {
      for (int e = 0; e < nv; ++e) dptr[e] = static_cast<OutT>(o[e]);
    }
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

// 4-D tiled tensor map over fp32 (f16 = 0) or fp16 (f16 = 1) elements; strides in bytes; swizzle 128, 64 or 32 bytes
int g_tmap_promo256 = 0;   // set around an encode call: 256-byte L2 promotion (streams whose next box continues the same rows)
const char* encode4x(CUtensorMap* m, const void* base, int f16, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4], int swizzle_bytes) {
  EncodeFn fn = get_encode();
  if (!fn) return "cuTensorMapEncodeTiled entry point not available";
  cuuint64_t gd[4]; cuuint64_t gs[3]; cuuint32_t bx[4]; cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 3; ++i) gs[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return "tensor base not 16-byte aligned";
  for (int i = 0; i < 3; ++i)
    if (gs[i] % 16 != 0 || gs[i] == 0) return "tensor stride not a positive multiple of 16 bytes";
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gd,
                  gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  g_tmap_promo256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
const char* encode4(CUtensorMap* m, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                    const uint32_t box[4]) {
  return encode4x(m, base, 0, dims, strides_bytes, box, 128);
}

// plain operand [rows][K] with (h, b) batch strides (in elements); stride 0 => broadcast (extent 1).
// box = box_cols x min(box_rows, rows); *bytes = bytes one box delivers
const char* encode_plainx(CUtensorMap* m, const void* base, int f16, int rows, int K, long ld, long sh, int nh, long sb,
                          int nb, int box_cols, int box_rows, int swizzle_bytes, int* hmul, int* bmul, uint32_t* bytes) {
  const uint64_t es = f16 ? 2 : 4;
  box_rows = std::min(box_rows, rows);
  *bytes = uint32_t(box_rows) * box_cols * es;
  *hmul = (sh != 0 && nh > 1) ? 1 : 0;
  *bmul = (sb != 0 && nb > 1) ? 1 : 0;
  uint64_t dims[4] = {uint64_t(K), uint64_t(rows), uint64_t(*hmul ? nh : 1), uint64_t(*bmul ? nb : 1)};
  uint64_t st[3] = {uint64_t(ld) * es, uint64_t(*hmul ? sh : ld) * es, uint64_t(*bmul ? sb : ld) * es};
  uint32_t box[4] = {uint32_t(box_cols), uint32_t(box_rows), 1, 1};
  return encode4x(m, base, f16, dims, st, box, swizzle_bytes);
}
const char* encode_plain(CUtensorMap* m, const float* base, int rows, int K, long ld, long sh, int nh, long sb,
                         int nb, int box_rows, int* hmul, int* bmul, uint32_t* bytes) {
  return encode_plainx(m, base, 0, rows, K, ld, sh, nh, sb, nb, BK, box_rows, 128, hmul, bmul, bytes);
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

static int sm_count() { return pbhost::sm_count(); }

template <int BN, int STAGES, bool AB16, bool D16, bool PAIR = false, bool GEGLU = false>
static const char* launch_t(const Params& p, int grid, cudaStream_t st) {
  using S = Smem<BN, STAGES, PAIR>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(2 * BN <= 512 && BN <= ACC_STRIDE, "two accumulator stages must fit TMEM");
  if (const char* err = pbhost::optin_smem(gemm_tc_kernel<BN, STAGES, AB16, D16, PAIR, GEGLU>, S::TOTAL)) return err;
  cudaError_t e;
  if constexpr (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, AB16, D16, PAIR, GEGLU>, p);
  } else {
    gemm_tc_kernel<BN, STAGES, AB16, D16><<<grid, NTHREADS, S::TOTAL, st>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cudaGetErrorString(e);
  if (p.splits > 1) {
    const int ntail = (p.items - p.full_tiles) / p.splits;
    const long threads = static_cast<long>(ntail) * BM * (BN / 4);
    const int blocks = (int)std::min<long>((threads + 255) / 256, 8L * sm_count());
    splitk_reduce_k<BN, D16><<<blocks, 256, 0, st>>>(p, ntail);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// CTA-pair kernels: 256 x 160 (6 stages of 26 KB) and 256 x 256 (5 stages of 32 KB)
template <bool D16>
static const char* launch_pair(int BN, const Params& p, int grid, cudaStream_t st) {
  if (BN == 160) return launch_t<160, 6, true, D16, true>(p, grid, st);
  if (BN == 256) return launch_t<256, 5, true, D16, true>(p, grid, st);
  return "gemm: unsupported pair tile width";
}

// co-resident CTA pairs of the pair kernel on the current device (74 on a B200), cached per device
static int pair_clusters() {
  static int n[pbhost::kMaxDevices] = {};
  const int dev = pbhost::current_device();
  if (dev < 0 || dev >= pbhost::kMaxDevices) return 0;
  if (!n[dev]) {
    using S = Smem<256, 5, true>;
    auto* fn = gemm_tc_kernel<256, 5, true, true, true>;
    n[dev] = -1;
    if (pbhost::optin_smem(fn, S::TOTAL) == nullptr) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sm_count()); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = S::TOTAL;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int c = 0;
      if (cudaOccupancyMaxActiveClusters(&c, fn, &cfg) == cudaSuccess && c > 0) n[dev] = c;
      else (void)cudaGetLastError();
    }
  }
  return n[dev] > 0 ? n[dev] : 0;
}

template <bool AB16, bool D16>
static const char* launch_bn(int BN, const Params& p, int grid, cudaStream_t st) {
  if (BN == 160) return launch_t<160, 4, AB16, D16>(p, grid, st);
  if (BN == 128) return launch_t<128, 5, AB16, D16>(p, grid, st);
  if (BN == 64) return launch_t<64, 6, AB16, D16>(p, grid, st);
  return "gemm: unsupported tile width";
}

}  // namespace pbgemm

// tuning hooks (scripts/bench_gemm.py): force a tile width (0 = heuristic), split policy (0 off, 1 heuristic,
// >1 forced split count), allow BN = 160
static int g_force_bn = 0, g_split = 1, g_use160 = 1, g_split_min_kb = 24;
static int g_pair = getenv("PB_GEMM_PAIR") ? atoi(getenv("PB_GEMM_PAIR")) : 1;   // CTA-pair kernels (A/B switch)
extern "C" __attribute__((visibility("default"))) void pb_gemm_tune(int force_bn, int split, int use160) {
  g_force_bn = force_bn; g_split = split; g_use160 = use160;
}
extern "C" __attribute__((visibility("default"))) void pb_gemm_tune_split_min_kb(int kb) { g_split_min_kb = kb; }
extern "C" __attribute__((visibility("default"))) void pb_gemm_tune_pair(int on) { g_pair = on; }

// Returns nullptr on success, else a static error string.  Stream-ordered, no host sync.
const char* pb_gemm_launch(const PbGemm& g, cudaStream_t st) {
  using namespace pbgemm;
  if (g.M <= 0 || g.N <= 0) return "gemm: empty problem";
  const int ab16 = g.ab_dtype == PB_GEMM_F16, d16 = g.d_dtype == PB_GEMM_F16;
  const int KB = ab16 ? BK16 : BK;            // k-block in elements (128 bytes)
  const long aes = ab16 ? 2 : 4, des = d16 ? 2 : 4;
  const int dq = d16 ? 8 : 4;                 // elements per 16 bytes of D / R
  Params p;
  memset(&p, 0, sizeof p);
  p.nseg = g.nseg; p.M = g.M; p.N = g.N; p.nb = g.nb; p.nh = g.nh;
  p.conv = g.conv; p.H = g.H; p.W = g.W;
  p.D = g.D; p.R = g.R; p.bias = g.bias;
  p.ldd = g.ldd; p.sDb = g.sDb; p.sDh = g.sDh; p.ldr = g.ldr; p.sRb = g.sRb; p.sRh = g.sRh;
  p.alpha = g.alpha; p.beta = g.R ? g.beta : 0.f; p.round_tf32 = g.round_tf32;
  const bool geglu = g.gg != nullptr;
  if (geglu) {
    if (!ab16 || !d16 || g.nseg != 1 || g.conv || g.nb != 1 || g.nh != 1 || g.R || g.bias || g.alpha != 1.f || g.N != 2 * g.gg_F ||
        g.gg_F % 128 || g.gg_rows_p <= 0 || (reinterpret_cast<uintptr_t>(g.gg) & 15) || (g.gg_p_stride % 4))
      return "gemm: the GEGLU epilogue takes an fp16 plain GEMM (nb = nh = 1, no residual / bias, alpha = 1) with N = 2 F, F % 128 == 0";
    p.gg = g.gg; p.gg_F = g.gg_F; p.gg_rows_p = unsigned(g.gg_rows_p);
    p.gg_k_slot = g.gg_k_slot > 0 ? unsigned(g.gg_k_slot) : 0x7fffffffu; p.gg_p_stride = g.gg_k_slot > 0 ? g.gg_p_stride : 0;
  }
  if ((g.ldd % dq) || (g.R && (g.ldr % dq)) || (reinterpret_cast<uintptr_t>(g.D) & 15) ||
      (g.R && (reinterpret_cast<uintptr_t>(g.R) & 15)) || (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)))
    return "gemm: D/R/bias must be 16-byte aligned with a leading dimension that is a multiple of 16 bytes";

  long mt;
  if (g.conv) {
    p.bw = std::min(pow2_floor(g.W), 128);
    p.bh = std::min(pow2_floor(g.H), 128 / p.bw);
    p.bb = std::min(128 / (p.bw * p.bh), g.nb);
    p.tiles_w = (g.W + p.bw - 1) / p.bw;
    p.tiles_h = (g.H + p.bh - 1) / p.bh;
    mt = (long)p.tiles_w * p.tiles_h * ((g.nb + p.bb - 1) / p.bb);
    p.taps = 9;
  } else {
    mt = (g.M + BM - 1) / BM;
    p.taps = 1;
  }
  const long nz = g.conv ? 1 : (long)g.nb * g.nh;
  const int nsm = sm_count();
  // tile width: 160 when it divides N (every SD channel count is a multiple of 160: no padded columns), 64 for narrow
  // outputs, else 128
  int BN = 128;
  if (g_use160 && g.N % 160 == 0) BN = 160;
  else if (g.N <= 64) BN = 64;
  else if (g.N <= 128 || g.N % 128 == 0) BN = 128;
  else if (g.N % 128 <= 64 && mt * nz * ((g.N + 63) / 64) <= 2L * nsm) BN = 64;
  if (g_force_bn) BN = g_force_bn;
  // CTA-pair kernel (256-row tiles over two SMs, see gemm_tc_kernel): fp16 weight GEMMs and convolutions with at least one
  // full wave of tiles; attention products (batched B, two segments) and the small-M layers keep the one-CTA kernel
  int pair = 0;
  int nsm_eff = nsm;
  const int kb_tile = (g.conv ? 9 : 1) * ((g.seg[0].K + KB - 1) / KB);   // k-blocks per tile
  if (geglu && (!g_pair || pair_clusters() <= 0)) return "gemm: the GEGLU epilogue needs the CTA-pair kernel";
  if (g_pair && ab16 && g.nseg == 1 && (g.conv || (g.nb == 1 && g.nh == 1)) && g.N >= EPI_W &&
      (g.N % 256 == 0 || g.N % 160 == 0) && (!g_force_bn || g_force_bn == 160 || g_force_bn == 256)) {
    const int ncl = pair_clusters();
    const int pbn = g_force_bn ? g_force_bn : (g.N % 256 == 0 ? 256 : 160);
    const long ptiles = ((mt + 1) / 2) * 2 * ((g.N + pbn - 1) / pbn);
    // at least one full wave of tiles; the deep-K small-M layers (8 x 8 convolutions: 70 tiles of 180 k-blocks, all cut along
    // K) gain from the pair MMA as well (62 -> 49 us), the short ones do not
    static const long pair_min_env = getenv("PB_GEMM_PAIR_MIN") ? atol(getenv("PB_GEMM_PAIR_MIN")) : 0;
    const long pair_min = pair_min_env ? pair_min_env : (kb_tile >= 64 ? 48 : 2L * ncl);
    if (geglu && (pbn != 256 || g_force_bn)) return "gemm: the GEGLU epilogue needs 256-column tiles";
    if (ncl > 0 && g.N % pbn == 0 && (ptiles >= pair_min || geglu)) { pair = 1; BN = pbn; nsm_eff = 2 * ncl; mt = ((mt + 1) / 2) * 2; }
  }
  p.pair = pair;
  const long nt = (g.N + BN - 1) / BN;

  int ktot = 0;
  for (int s = 0; s < g.nseg; ++s) {
    const PbGemmSeg& sg = g.seg[s];
    if (sg.K <= 0) return "gemm: empty K segment";
    p.kblocks[s] = (sg.K + KB - 1) / KB;
    ktot += p.kblocks[s];
    const char* err;
    if (g.conv) {
      if (g.nseg != 1) return "gemm: conv mode takes one segment";
      if (sg.K % 8) return "gemm: conv channels must be a multiple of 8";
      p.conv_ctot = sg.K;
      uint64_t dims[4] = {uint64_t(sg.K), uint64_t(g.W), uint64_t(g.H), uint64_t(g.nb)};
      uint64_t stb[3] = {uint64_t(sg.lda) * aes, uint64_t(sg.lda) * aes * g.W, uint64_t(sg.lda) * aes * g.W * g.H};
      uint32_t box[4] = {uint32_t(KB), uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bb)};
      err = encode4x(&p.mapA[s], sg.A, ab16, dims, stb, box, 128);
      if (err) return err;
      // packed filter [N][9][C] as a (C, tap, N) tensor
      uint64_t bd[4] = {uint64_t(sg.K), 9, uint64_t(g.N), 1};
      uint64_t bs[3] = {uint64_t(sg.K) * aes, uint64_t(sg.ldb) * aes, uint64_t(sg.ldb) * aes};
      const uint32_t brow = (uint32_t)std::min(pair ? BN / 2 : BN, g.N);
      uint32_t bbox[4] = {uint32_t(KB), 1, brow, 1};
      err = encode4x(&p.mapB[s], sg.B, ab16, bd, bs, bbox, 128);
      if (err) return err;
      p.tx_bytes[s] = uint32_t(p.bw * p.bh * p.bb + brow) * 128;
    } else {
      uint32_t abytes, bbytes;
      // problem slots: a primal (batch-broadcast) operand gets the problem index as its batch axis
      const int kslot = (g.k_slot > 0 && g.k_slot < g.nb) ? g.k_slot : 0;
      const int nslots = kslot ? g.nb / kslot : 1;
      if (kslot && (g.nb % kslot || g.p_stride <= 0 || g.p_stride % 16)) return "gemm: problem slots need nb % k_slot == 0 and a positive p_stride multiple of 16 bytes";
      long sAb = sg.sAb, sBb = sg.sBb;
      int nbA = g.nb, nbB = g.nb;
      if (nslots > 1 && sAb == 0) { sAb = g.p_stride / (long)aes; nbA = nslots; p.a_bdiv[s] = kslot; }
      if (nslots > 1 && sBb == 0) { sBb = g.p_stride / (long)aes; nbB = nslots; p.b_bdiv[s] = kslot; }
      err = encode_plainx(&p.mapA[s], sg.A, ab16, g.M, sg.K, sg.lda, sg.sAh, g.nh, sAb, nbA, KB, BM, 128, &p.a_hmul[s],
                          &p.a_bmul[s], &abytes);
      if (err) return err;
      err = encode_plainx(&p.mapB[s], sg.B, ab16, g.N, sg.K, sg.ldb, sg.sBh, g.nh, sBb, nbB, KB, pair ? BN / 2 : BN, 128,
                          &p.b_hmul[s], &p.b_bmul[s], &bbytes);
      if (err) return err;
      p.tx_bytes[s] = abytes + bbytes;
    }
  }
  ktot *= p.taps;

  // epilogue tensor maps (D store / R load through the staging buffers: 32-column chunks, 128-byte fp32 or 64-byte fp16
  // rows); narrow outputs use guarded direct stores
  p.epi_tma = g.N >= EPI_W ? 1 : 0;
  if (p.epi_tma) {
    const char* err;
    const int swz = d16 ? 64 : 128;
    if (g.conv) {
      uint32_t box[4] = {uint32_t(EPI_W), uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bb)};
      uint64_t dims[4] = {uint64_t(g.N), uint64_t(g.W), uint64_t(g.H), uint64_t(g.nb)};
      uint64_t sd[3] = {uint64_t(g.ldd) * des, uint64_t(g.ldd) * des * g.W, uint64_t(g.ldd) * des * g.W * g.H};
      err = encode4x(&p.mapD, g.D, d16, dims, sd, box, swz);
      if (err) return err;
      if (g.R) {
        uint64_t sr[3] = {uint64_t(g.ldr) * des, uint64_t(g.ldr) * des * g.W, uint64_t(g.ldr) * des * g.W * g.H};
        err = encode4x(&p.mapR, g.R, d16, dims, sr, box, swz);
        if (err) return err;
      }
      p.r_bytes = uint32_t(p.bw * p.bh * p.bb) * EPI_W * des;
    } else {
      int hm, bm;
      err = encode_plainx(&p.mapD, g.D, d16, g.M, geglu ? g.gg_F : g.N, g.ldd, g.sDh, g.nh, g.sDb, g.nb, EPI_W, BM, swz, &p.d_hmul,
                          &p.d_bmul, &p.r_bytes);
      if (err) return err;
      if ((g.nh > 1 && !p.d_hmul) || (g.nb > 1 && !p.d_bmul)) return "gemm: D needs a stride for every batch dimension";
      if (g.R) {
        err = encode_plainx(&p.mapR, g.R, d16, g.M, g.N, g.ldr, g.sRh, g.nh, g.sRb, g.nb, EPI_W, BM, swz, &hm, &bm, &p.r_bytes);
        if (err) return err;
        if (hm != p.d_hmul || bm != p.d_bmul) return "gemm: R must be batched like D";
      }
    }
  }

  if (!g.conv && g.nb > 1 && g.seg[0].sAb == 0 && (g.nseg == 1 || g.seg[1].sAb == 0)) p.raster_b = 1;
  p.mt = (int)mt; p.nt = (int)nt;
  p.splits = 1; p.kb_per_split = ktot;
  const long tiles = mt * nt * nz;
  // tiles of the last, partial wave are cut along K so that the wave fills the machine (see the header comment); only
  // when the K loop is deep enough for the saving to outweigh the reduction pass
  long full = tiles, tail = 0;
  static const int env_min_kb = getenv("PB_SPLIT_MIN_KB") ? atoi(getenv("PB_SPLIT_MIN_KB")) : 0;     // A/B switch
  if (g_split && g.ws && !geglu && ktot >= (env_min_kb ? env_min_kb : g_split_min_kb)) {
    tail = tiles % nsm_eff;
    // split count of the tail wave: minimise (rounds the tail items need) x (k-blocks per item); a tail of 80 tiles is
    // better cut in 5 (3 rounds of 1/5) than run whole on 80 of the 148 SMs
    int s = 1;
    if (tail) {
      long best = (long)ktot + 1;
      const int smax = (int)std::min<long>(std::min(32, ktot / 8), g.ws_floats / (tail * BM * BN));
      for (int c = 1; c <= smax; ++c) {
        const long kb = (ktot + c - 1) / c;
        // in k-block times: rounds x (k-blocks + pipeline fill and partial-tile store) + the reduce pass over c partials
        const long cost = c == 1 ? kb : ((tail * c + nsm_eff - 1) / nsm_eff) * (kb + 6) + tail * c / 16;
        if (cost < best) { best = cost; s = c; }
      }
    }
    if (g_split > 1) s = g_split;
    s = std::min(s, ktot / 8);
    if (tail) s = (int)std::min<long>(s, g.ws_floats / (tail * BM * BN));
    if (tail && s > 1) {
      p.kb_per_split = (ktot + s - 1) / s;
      p.splits = (ktot + p.kb_per_split - 1) / p.kb_per_split;
      full = tiles - tail;
      p.ws = g.ws;
      uint64_t dims[4] = {uint64_t(BN), uint64_t(tail) * p.splits * BM, 1, 1};
      uint64_t sw[3] = {uint64_t(BN) * 4, uint64_t(BN) * 4, uint64_t(BN) * 4};
      uint32_t box[4] = {uint32_t(EPI_W), uint32_t(BM), 1, 1};
      const char* err = encode4x(&p.mapW, g.ws, 0, dims, sw, box, 128);
      if (err) return err;
    } else {
      tail = 0;
    }
  }
  p.full_tiles = (int)full;
  p.items = (int)(full + tail * p.splits);
  static const int env_trace = getenv("PB_GEMM_TRACE") ? atoi(getenv("PB_GEMM_TRACE")) : 0;
  static long long* trace_buf = nullptr;
  if (env_trace) {
    if (!trace_buf) cudaMalloc(&trace_buf, 64 * 8 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 64 * 8 * sizeof(long long), st);
    p.trace = trace_buf;
  }
  const int grid = (int)std::min<long>(p.items, nsm_eff);
  const char* lerr;
  if (geglu) lerr = pair ? launch_t<256, 5, true, true, true, true>(p, grid, st) : "gemm: the GEGLU epilogue needs the CTA-pair kernel";
  else if (pair) lerr = d16 ? launch_pair<true>(BN, p, grid, st) : launch_pair<false>(BN, p, grid, st);
  else if (ab16) lerr = d16 ? launch_bn<true, true>(BN, p, grid, st) : launch_bn<true, false>(BN, p, grid, st);
  else lerr = d16 ? launch_bn<false, true>(BN, p, grid, st) : launch_bn<false, false>(BN, p, grid, st);
  if (env_trace && !lerr) {          // debugging aid: synchronous dump of CTA 0's per-item clocks
    static int dumps = 0;
    long long h[64 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_buf, sizeof h, cudaMemcpyDeviceToHost);
    if (dumps++ < env_trace) {
      fprintf(stderr, "GEMMTRACE M=%d N=%d kb=%d BN=%d pair=%d items=%d grid=%d\n", g.M, g.N, ktot, BN, pair, p.items, grid);
      const long long t0 = h[0];
      for (int i = 0; i < 64 && h[i * 8 + 2]; ++i)
        fprintf(stderr, "  item %2d  prod %7lld..%7lld  mma %7lld..%7lld  epi wait %7lld start %7lld\n", i, h[i * 8] - t0,
                h[i * 8 + 1] - t0, h[i * 8 + 2] - t0, h[i * 8 + 3] - t0, h[i * 8 + 6] - t0, h[i * 8 + 4] - t0);
    }
  }
  return lerr;
}
