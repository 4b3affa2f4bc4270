// Subspace re-orthonormalisation for the power iteration, entirely on the device and free of
// cuSOLVER/cuBLAS: the reference calls torch.linalg.svd on the k x n_in matrix W = U^T J every
// iteration (`src/utils/utils.py:799`) and keeps (s, V).  For k <= 64 << n_in the same (s, V) follow from
// the k x k Gram matrix: W W^T = X diag(lambda) X^T  =>  svdvals(W) = sqrt(lambda),  V = diag(lambda)^-1/2 X^T W.
//   pbk_gram2  : G = W W^T and M = W Vprev^T, fp32 products accumulated in fp64 (HBM-bound, W read once)
//   pbk_jacobi : cyclic Jacobi eigen-solver on G in fp64 (one warp), descending sort, row signs chosen so
//                that <V_i, Vprev_i> >= 0 (LAPACK's sign is arbitrary; a continuous sign makes the reference's
//                sign-sensitive convergence test meaningful), emits the k x k map Rm and s = lambda^(1/4)
//   pbk_rotate : V = Rm W plus the convergence metrics against Vprev, fused in one pass
#include <cuda_runtime.h>
#include <cstdint>
#include <algorithm>

#include "pb_host_util.h"
#include "pb_kernels.h"

namespace {

constexpr int kMaxK = 64;
constexpr int kChunkMax = 512;  // columns of W staged per block (shrinks for large k to fit shared memory)

inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void gram2_k(const float* __restrict__ Wm, const float* __restrict__ Vp, int k, long n, int kChunk,
                        double* __restrict__ G, double* __restrict__ M) {
  extern __shared__ float sh[];                  // W tile [k][kChunk], V tile [k][kChunk]
  float* sw = sh;
  float* sv = sh + (size_t)k * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long c0 = (long)blockIdx.x * kChunk; c0 < n; c0 += (long)gridDim.x * kChunk) {
    const int len = (int)min((long)kChunk, n - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < k * kChunk; i += blockDim.x) {
      const int r = i / kChunk, c = i % kChunk;
      sw[i] = c < len ? Wm[(long)r * n + c0 + c] : 0.f;
      sv[i] = (Vp && c < len) ? Vp[(long)r * n + c0 + c] : 0.f;
    }
    __syncthreads();
    const int npair = k * k;
    for (int pr = warp; pr < 2 * npair; pr += nwarp) {
      const bool second = pr >= npair;
      if (second && !Vp) break;
      const int q = second ? pr - npair : pr;
      const int i = q / k, j = q % k;
      if (!second && j < i) continue;            // G is symmetric: upper triangle only
      const float* a = sw + (size_t)i * kChunk;
      const float* b = (second ? sv : sw) + (size_t)j * kChunk;
      double acc = 0.0;
      for (int c = lane; c < kChunk; c += 32) acc += (double)a[c] * (double)b[c];
      acc = warp_sum_d(acc);
      if (lane == 0) {
        if (second) atomicAdd(&M[q], acc);
        else { atomicAdd(&G[i * k + j], acc); if (i != j) atomicAdd(&G[j * k + i], acc); }
      }
    }
  }
}

// one warp; A (k x k, fp64) and eigenvectors X live in shared memory
__global__ void jacobi_k(const double* __restrict__ G, const double* __restrict__ M, int k, float* __restrict__ Rm,
                         float* __restrict__ sv) {
  extern __shared__ double jsh[];                // A [k][k], X [k][k]
  double* A = jsh;
  double* X = jsh + k * k;
  __shared__ double lam[kMaxK];
  __shared__ int order[kMaxK];
  __shared__ double cs[2];
  const int lane = threadIdx.x;
  for (int i = lane; i < k * k; i += 32) { A[i] = G[i]; X[i] = (i / k == i % k) ? 1.0 : 0.0; }
  __syncwarp();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = lane; i < k * k; i += 32) {
      const double v = A[i] * A[i];
      if (i / k == i % k) diag += v; else off += v;
    }
    off = warp_sum_d(off); diag = warp_sum_d(diag);
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < k - 1; ++p) {
      for (int q = p + 1; q < k; ++q) {
        if (lane == 0) {
          const double apq = A[p * k + q];
          double c = 1.0, s = 0.0;
          if (fabs(apq) > 1e-300) {
            const double tau = (A[q * k + q] - A[p * k + p]) / (2.0 * apq);
            const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t); s = t * c;
          }
          cs[0] = c; cs[1] = s;
        }
        __syncwarp();
        const double c = cs[0], s = cs[1];
        // A <- J^T A J, X <- X J  (J rotates columns p,q)
        for (int i = lane; i < k; i += 32) {
          const double aip = A[i * k + p], aiq = A[i * k + q];
          A[i * k + p] = c * aip - s * aiq; A[i * k + q] = s * aip + c * aiq;
          const double xip = X[i * k + p], xiq = X[i * k + q];
          X[i * k + p] = c * xip - s * xiq; X[i * k + q] = s * xip + c * xiq;
        }
        __syncwarp();
        for (int i = lane; i < k; i += 32) {
          const double api = A[p * k + i], aqi = A[q * k + i];
          A[p * k + i] = c * api - s * aqi; A[q * k + i] = s * api + c * aqi;
        }
        __syncwarp();
      }
    }
  }
  for (int i = lane; i < k; i += 32) lam[i] = A[i * k + i];
  __syncwarp();
  if (lane == 0) {                               // descending selection sort of k <= 64 eigenvalues
    for (int i = 0; i < k; ++i) order[i] = i;
    for (int i = 0; i < k; ++i) {
      int best = i;
      for (int j = i + 1; j < k; ++j) if (lam[order[j]] > lam[order[best]]) best = j;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
  }
  __syncwarp();
  const double lmax = fmax(lam[order[0]], 0.0);
  for (int i = lane; i < k; i += 32) {
    const int e = order[i];
    // Numerically null direction of W: W W^T squares the conditioning, so an eigenvalue below ~1e-13 of the largest carries
    // no information (fp64 Gram of fp32 data).  torch.linalg.svd would return an arbitrary orthonormal completion there; this
    // returns a ZERO row with s = 0 instead of round-off divided by ~0 (a non-orthonormal basis would silently corrupt the
    // next J V).  A zero column stays zero through the iteration (J 0 = 0), so the caller sees s_i = 0, v_i = 0.
    if (!(lam[e] > 1e-13 * lmax)) {
      for (int j = 0; j < k; ++j) Rm[i * k + j] = 0.f;
      sv[i] = 0.f;
      continue;
    }
    const double l = fmax(lam[e], 1e-300);
    const double inv = 1.0 / sqrt(l);
    double dot = 0.0;                            // <V_i, Vprev_i> = inv * sum_j X[j][e] M[j][i]
    if (M) for (int j = 0; j < k; ++j) dot += X[j * k + e] * M[j * k + i];
    const double sg = (M && dot < 0.0) ? -1.0 : 1.0;
    for (int j = 0; j < k; ++j) Rm[i * k + j] = (float)(sg * inv * X[j * k + e]);
    sv[i] = (float)sqrt(sqrt(l));
  }
}

__global__ void rotate_k(const float* __restrict__ Wm, const float* __restrict__ Rm, const float* __restrict__ Vp, int k,
                         long n, float atol, float rtol, float* __restrict__ V, float* __restrict__ metrics) {
  extern __shared__ float rs[];                  // Rm [k][k]
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) rs[i] = Rm[i];
  __syncthreads();
  float d2 = 0.f, viol = 0.f;
  for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
    float w[kMaxK];
#pragma unroll 8
    for (int j = 0; j < kMaxK; ++j) w[j] = j < k ? Wm[(long)j * n + c] : 0.f;
    for (int i = 0; i < k; ++i) {
      float v = 0.f;
#pragma unroll 8
      for (int j = 0; j < kMaxK; ++j) if (j < k) v = fmaf(rs[i * k + j], w[j], v);
      V[(long)i * n + c] = v;
      if (Vp) {
        const float d = v - Vp[(long)i * n + c];
        d2 = fmaf(d, d, d2);
        if (fabsf(d) > atol + rtol * fabsf(v)) viol += 1.f;
      }
    }
  }
  if (metrics) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d2 += __shfl_xor_sync(0xffffffffu, d2, o); viol += __shfl_xor_sync(0xffffffffu, viol, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&metrics[0], d2); atomicAdd(&metrics[1], viol); }
  }
}

}  // namespace

PBK pbk_gram2(const float* Wm, const float* Vprev, int k, long n, double* G, double* M, pb_stream st) {
  if (k < 1 || k > kMaxK) return "ortho: pca_rank must be in [1, 64]";
  cudaStream_t s = static_cast<cudaStream_t>(st);
  cudaMemsetAsync(G, 0, sizeof(double) * k * k, s);
  if (M) cudaMemsetAsync(M, 0, sizeof(double) * k * k, s);
  int kChunk = std::min(kChunkMax, (96 * 1024 / (2 * k * 4)) / 32 * 32);
  const size_t shmem = (size_t)2 * k * kChunk * sizeof(float);
  if (const char* err = pbhost::optin_smem(gram2_k, 100 * 1024)) return err;
  const unsigned grid = (unsigned)std::max<long>(1, std::min<long>((n + kChunk - 1) / kChunk, 148 * 2));
  gram2_k<<<grid, 256, shmem, s>>>(Wm, Vprev, k, n, kChunk, G, M);
  return last_err();
}
PBK pbk_jacobi(const double* G, const double* M, int k, float* Rm, float* sv, pb_stream st) {
  if (k < 1 || k > kMaxK) return "ortho: pca_rank must be in [1, 64]";
  if (const char* err = pbhost::optin_smem(jacobi_k, 2 * kMaxK * kMaxK * 8)) return err;
  jacobi_k<<<1, 32, (size_t)2 * k * k * sizeof(double), static_cast<cudaStream_t>(st)>>>(G, M, k, Rm, sv);
  return last_err();
}
PBK pbk_rotate(const float* Wm, const float* Rm, const float* Vprev, int k, long n, float atol, float rtol, float* V,
               float* metrics, pb_stream st) {
  if (k < 1 || k > kMaxK) return "ortho: pca_rank must be in [1, 64]";
  cudaStream_t s = static_cast<cudaStream_t>(st);
  if (metrics) cudaMemsetAsync(metrics, 0, 2 * sizeof(float), s);
  const unsigned grid = (unsigned)std::max<long>(1, std::min<long>((n + 127) / 128, 148 * 8));
  rotate_k<<<grid, 128, (size_t)k * k * sizeof(float), s>>>(Wm, Rm, Vprev, k, n, atol, rtol, V, metrics);
  return last_err();
}
