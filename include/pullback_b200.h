/* pullback_b200 -- C ABI of the B200-native pullback hot path.
 *
 * Drop-in boundary for the one hot path of enkeejunior1/Diffusion-Pullback:
 *   local_encoder_pullback_zt   (reference src/utils/utils.py:722-816, bound at utils.py:333)
 *   local_encoder_pullback_xt   (reference src/utils/utils.py:165-249, bound at utils.py:104)
 *   get_h / get_h_uncond        (reference src/utils/utils.py:438-527 / :114-163, bound at :326 / :103)
 * The reference is pure Python; these entry points are what a ctypes binding for that path binds
 * (see INTEGRATION.md for the stub).  Plain pointers and sizes only, no torch / C++ types.
 *
 * Memory: the library never allocates user-visible device memory.  The caller allocates three
 * device regions whose sizes the library reports after pb_plan(): the packed-weight region, the
 * primal cache and the tangent workspace.  All calls are ordered on the cudaStream_t passed in
 * (as void*); only pb_pullback() synchronises (once per iteration after min_iter, to evaluate the
 * reference's early-exit test).  One handle per (device, U-Net); not thread-safe per handle.
 *
 * Every function returns 0 on success or a negative pb_status; pb_last_error() gives the text.
 */
#ifndef PULLBACK_B200_H
#define PULLBACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_handle pb_handle;

enum pb_status { PB_OK = 0, PB_EINVAL = -1, PB_ESTATE = -2, PB_ECUDA = -3, PB_EMISSING = -4 };

enum pb_unet_kind {
  PB_UNET_COND = 0,   /* diffusers UNet2DConditionModel (Stable Diffusion): get_h, utils.py:438-527   */
  PB_UNET_UNCOND = 1  /* diffusers UNet2DModel (DDPM CelebA-HQ ...): get_h_uncond, utils.py:114-163   */
};
enum pb_op { PB_OP_MID = 0, PB_OP_UP = 1,   /* `op='down'` raises in the reference  */
             PB_OP_FULL = 2,                 /* the whole U-Net: x_t -> eps (up path, conv_norm_out, SiLU, conv_out); block_idx 0 */
             PB_OP_DEC = 3 };                /* decoder side (get_h_to_e, utils.py:529-635): the map h -> eps with h the mid-block output and
                                                the skip connections of x_t held fixed; block_idx 0.  n_in = numel(h), n_out = numel(x_t):
                                                pb_jvp / pb_vjp / pb_pullback then act on J_dec = d eps / d h (local_decoder_pullback_zt,
                                                utils.py:818-898); pb_set_point takes x_t as before (h_out receives eps) */

#define PB_MAX_LEVELS 8

/* Architecture of the U-Net whose x_t -> h map is linearised (diffusers 0.11.0 config subset). */
typedef struct pb_unet_cfg {
  int32_t kind;                              /* pb_unet_kind */
  int32_t in_channels;
  int32_t n_levels;                          /* len(block_out_channels) */
  int32_t block_out_channels[PB_MAX_LEVELS];
  int32_t down_has_attn[PB_MAX_LEVELS];      /* CrossAttnDownBlock2D / AttnDownBlock2D */
  int32_t up_has_attn[PB_MAX_LEVELS];        /* CrossAttnUpBlock2D (COND only) */
  int32_t heads[PB_MAX_LEVELS];              /* attention heads per level (SD 1.x: 8; SD 2.x: 5,10,20,20; DDPM: 1) */
  int32_t layers_per_block;
  int32_t cross_attention_dim;               /* COND only */
  int32_t norm_num_groups;
  float norm_eps;                            /* resnet GroupNorm eps (transformer GroupNorm uses 1e-6, LayerNorm 1e-5) */
  int32_t flip_sin_to_cos;
  float freq_shift;
  int32_t downsample_padding;                /* 1: symmetric pad (SD); 0: (0,1,0,1) pad (DDPM) */
} pb_unet_cfg;

/* One named parameter in PyTorch layout, resident on the device (a state_dict entry). */
typedef struct pb_tensor_desc {
  const char* name;                          /* diffusers state_dict key, e.g. "mid_block.resnets.0.conv1.weight" */
  const float* data;                         /* device pointer, contiguous fp32 */
  int32_t ndim;
  int64_t shape[4];
} pb_tensor_desc;

typedef struct pb_sizes {
  size_t packed_weight_bytes;                /* both conv/linear layouts, TF32-rounded */
  size_t primal_cache_bytes;                 /* linearisation cache for one (x_t, t, prompt) */
  size_t workspace_bytes;                    /* k_max tangents/cotangents + scratch */
  int64_t n_in, n_out;                       /* numel(x_t), numel(h); decoder side (PB_OP_DEC): numel(h), numel(eps) */
  int32_t out_channels, out_h, out_w;
  int32_t in_channels, in_h, in_w;           /* shape of the tangent input: x_t, or h on the decoder side */
  int64_t n_x;                               /* numel(x_t) -- what pb_set_point reads, whatever the op */
} pb_sizes;

typedef struct pb_iter_info {
  int32_t iters_done;
  int32_t converged;                         /* reference early-exit fired (allclose && i > min_iter) */
  float last_dist;                           /* || V_i - V_{i-1} ||_2  (the reference's printed metric) */
} pb_iter_info;

int pb_create(const pb_unet_cfg* cfg, pb_handle** out);
void pb_destroy(pb_handle* h);
const char* pb_last_error(const pb_handle* h);
const char* pb_backend(void);                /* "cuda-sm100a" for the product library */

/* Fix the problem geometry: latent H x W, truncation point (op, block_idx), largest rank, text length. */
int pb_plan(pb_handle* h, int32_t height, int32_t width, int32_t op, int32_t block_idx, int32_t k_max,
            int32_t ctx_len, pb_sizes* sizes);

/* Repack the state_dict once into `packed` (size pb_sizes.packed_weight_bytes). */
int pb_bind_weights(pb_handle* h, const pb_tensor_desc* table, int32_t n, void* packed, void* stream);

/* Primal pass at (x_t [C,H,W] NCHW, t, ctx [ctx_len, cross_attention_dim] or NULL): fills the linearisation
 * cache.  `primal_cache` / `workspace` are the caller's regions.  Optionally returns h (NCHW) in h_out. */
int pb_set_point(pb_handle* h, const float* x, float t, const float* ctx, void* primal_cache, void* workspace,
                 float* h_out, void* stream);
/* Decoder side (op = PB_OP_DEC) only, the nonlinear map get_h_to_e (utils.py:529-635): replaces the mid-block output cached by the
 * last pb_set_point by h_in ([C][H][W] as get_h returns it) and re-runs the decoder half on the same skip connections; eps_out
 * (optional) receives the noise prediction [in_channels][H][W].  The linearisation point of pb_jvp / pb_vjp moves to h_in. */
int pb_decode_from(pb_handle* h, const float* h_in, float* eps_out, void* stream);

/* Problem slots (throughput mode; the reference solves its problems one process after another, scripts of SURVEY.md s.3.1):
 * `slots` independent problems (x_t, t, ctx) share the handle's weights and run their tangent columns as ONE batch, so the
 * weight GEMMs see M = HW * k * slots.  k_max of pb_plan must be a multiple of `slots`; the region handed to pb_set_point as
 * `primal_cache` is then `slots` caches of cache_stride_bytes each.  pb_select_slot picks the slot the next pb_set_point
 * fills; pb_pullback takes the rank PER PROBLEM with V0 / vT [slots][k][n_in], u [slots][k][n_out], s [slots][k], and every
 * problem runs the same number of iterations (early exit only when all have converged).  slots = 1 (default): unchanged. */
int pb_set_slots(pb_handle* h, int32_t slots, size_t cache_stride_bytes);
int pb_select_slot(pb_handle* h, int32_t slot);

/* U[k][n_out] = J V,  V: [k][n_in]   (both NCHW-flattened rows; replaces the jacfwd call, utils.py:766-775) */
int pb_jvp(pb_handle* h, const float* V, int32_t k, float* U, void* stream);
/* W[k][n_in] = U^T J, U: [k][n_out]  (replaces autograd.functional.jacobian, utils.py:790-797) */
int pb_vjp(pb_handle* h, const float* U, int32_t k, float* W, void* stream);
/* (s, V) from W as torch.linalg.svd would give (utils.py:799): V rows orthonormal, s = sqrt(svdvals(W)),
 * descending.  Vprev (may be NULL) fixes the row signs and feeds metrics[0]=||V-Vprev||^2, metrics[1]=#violations
 * of allclose(atol,rtol=1e-5).  All device pointers; metrics may be NULL.  A numerically null direction of W (Gram
 * eigenvalue below 1e-13 of the largest: W has rank < k) comes back as s_i = 0 with a ZERO row V_i, where LAPACK would
 * return an arbitrary orthonormal completion. */
int pb_orthonormalize(pb_handle* h, const float* W, const float* Vprev, int32_t k, float atol, float* V, float* s,
                      float* metrics, void* stream);

/* The whole loop of utils.py:756-808: V0 [k][n_in] -> u [k][n_out] (the reference returns its transpose view),
 * s [k], vT [k][n_in].  Device pointers; info is host memory.  With problem slots (pb_set_slots) k is the rank per problem
 * and every buffer is slot-major: V0 / vT [slots][k][n_in], u [slots][k][n_out], s [slots][k]. */
int pb_pullback(pb_handle* h, const float* V0, int32_t k, int32_t min_iter, int32_t max_iter, float tol, float* u,
                float* s, float* vT, pb_iter_info* info, void* stream);

/* The same call with HOST buffers -- what a host-language binding of the reference's method hands over
 * (x_t, t, encoder_hidden_states, V0 in; u, s, vT out): copies in, runs the primal pass (pb_set_point) and
 * the iteration (pb_pullback) on the cache / workspace attached by an earlier pb_set_point, copies out.
 * Blocking. */
int pb_pullback_host(pb_handle* h, const float* x_host, float t, const float* ctx_host, const float* V0_host,
                     int32_t k, int32_t min_iter, int32_t max_iter, float tol, float* u_host, float* s_host,
                     float* vT_host, pb_iter_info* info, void* stream);

/* The host-buffer call for a handle with problem slots: x_host [slots][n_in], t_host [slots], ctx_host [slots][ctx_len][dim],
 * V0_host / vT_host [slots][k][n_in], u_host [slots][k][n_out], s_host [slots][k]; k is the rank per problem. */
int pb_pullback_host_slots(pb_handle* h, const float* x_host, const float* t_host, const float* ctx_host,
                           const float* V0_host, int32_t k, int32_t min_iter, int32_t max_iter, float tol, float* u_host,
                           float* s_host, float* vT_host, pb_iter_info* info, void* stream);

/* After pb_plan(): the state_dict entries (diffusers key + PyTorch shape, ndim <= 4) the planned path consumes. */
int pb_weight_count(const pb_handle* h);
int pb_weight_info(const pb_handle* h, int32_t index, const char** name, int32_t* ndim, int64_t* shape);

/* Diagnostics: kernels launched through this handle so far (graph replays count their kernel nodes);
 * options: "use_graph" (default 1), "round_tf32" (default 1; re-bind weights after changing it). */
int64_t pb_kernel_launches(const pb_handle* h);
int pb_set_option(pb_handle* h, const char* name, int value);

/* Diagnostics of the current plan (host-only, valid after pb_plan): how many ops of each kind the truncated U-Net became and
 * which of them the fp16-operand policy / the fused attention kernel cover. */
typedef struct pb_plan_info {
  int32_t n_ops, n_gemm, n_conv3x3;
  int32_t n_gemm_f16_jvp;             /* GEMMs whose JVP reads fp16 operands (A stored as halves by its producer) */
  int32_t n_gemm_f16_vjp_stored;      /* VJP: cotangent stored as halves by its single elementwise contributor */
  int32_t n_gemm_f16_vjp_converted;   /* VJP: 3x3 convs whose fp32 cotangent is converted once */
  int32_t n_gemm_d16_jvp;             /* JVP GEMMs writing fp16 output (sole consumer reads halves) */
  int32_t n_attn, n_attn_fused_self, n_attn_fused_cross, n_attn_p16;
} pb_plan_info;
int pb_plan_summary(const pb_handle* h, pb_plan_info* info);

/* One deterministic DDIM update (the reference's custom scheduler `step`, src/utils/utils.py:288-315, eta = 0):
 *   pred_x0 = (x - sqrt(1 - a_t) * eps) / sqrt(a_t);   x_next = sqrt(a_next) * pred_x0 + sqrt(1 - a_next) * eps
 * over n contiguous floats; x_next may alias x, pred_x0 may be NULL.  a_t / a_next are alphas_cumprod gathered by the
 * caller at t.long() / t_next.long() like `extract` (utils.py:1302-1317). */
int pb_ddim_step(const float* x, const float* eps, float a_t, float a_next, float* x_next, float* pred_x0, int64_t n,
                 void* stream);

/* out = a * x + b * y + c * z over n contiguous floats (y / z may be NULL with b / c = 0; out may alias any input): the two
 * elementwise updates of the reference's x-space guidance (src/modules/edit.py:484-502):
 *   zt_edit = zt + single_edit_step * vk            and            zt_edit = zt + scale * (et_edit - et_null). */
int pb_lincomb3(float* out, float a, const float* x, float b, const float* y, float c, const float* z, int64_t n, void* stream);

/* Timing probes for the roofline report (bench.py): after pb_profile_begin() every contraction-kernel launch made
 * through this handle is bracketed by an event pair on its stream (launches go out eagerly, no graph replay);
 * pb_profile_read() stops probing, waits, and returns the summed device time, the algorithmic flops (2 M N K per
 * product) and the launch count of one kernel class. */
#define PB_PROBE_GEMM 0                      /* gemm_tc_kernel: conv / linear / attention products */
#define PB_PROBE_ATTN 1                      /* attn_lin_kernel: fused attention linearisation */
#define PB_PROBE_GEMM_TF32 2                 /* the gemm_tc_kernel launches with fp32 operands (tcgen05 kind::tf32) */
#define PB_PROBE_GEMM_F16 3                  /* the gemm_tc_kernel launches with fp16 operands (tcgen05 kind::f16) */
int pb_profile_begin(pb_handle* h);
int pb_profile_read(pb_handle* h, int32_t kind, double* ms, double* flops, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* PULLBACK_B200_H */
