// Launchers of the fp16-tangent elementwise kernels (pb_lin16.cu); called from the pbk_* entry points of pb_kernels.cu when
// the io flags of a call say "tangent in halves, result in halves" (include/pb_kernels.h: PB_IN_F16 | PB_OUT_F16).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstddef>

namespace pb16 {

size_t gn_tmp_floats(int HW, int C, int G, int nb);
const char* gn_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, const float* beta, int HW, int C,
                   int G, int silu, const __half* t, int nb, int mode, __half* out, float acc, float* tmp, int k_slot,
                   long p_stride, cudaStream_t st);
const char* ln_lin(const float* xp, const float* mean, const float* rstd, const float* gamma, long rows_p, int C, const __half* t,
                   int nb, int mode, __half* out, float acc, int k_slot, long p_stride, cudaStream_t st);
const char* copy2d(__half* dst, long ldd, const __half* src, long lds, long rows, int cols, float beta, cudaStream_t st);
const char* col2im_s2(const __half* col, int nb, int H, int W, int C, int pad_lo, int Ho, int Wo, __half* gx, float beta,
                      cudaStream_t st);
const char* upsample2x_vjp(const __half* gy, int nb, int H, int W, int C, __half* gx, float beta, cudaStream_t st);
const char* to_f32(float* dst, const __half* src, size_t n, cudaStream_t st);

}  // namespace pb16
