// pb_engine -- host side of the pullback hot path behind the C ABI of include/pullback_b200.h.
//
// The truncated U-Net x_t -> h of the reference's get_h / get_h_uncond (src/utils/utils.py:438-527,
// :114-163; module semantics: diffusers 0.11.0) is planned ONCE into a flat op list over NHWC
// activations.  Three interpreters walk that list:
//   primal : runs the network at (x_t, t, ctx) and caches the linearisation (activations, norm statistics,
//            attention probabilities and the transposed operand copies the tcgen05 GEMM needs)
//   jvp    : pushes k tangents (packed on the batch axis) forward   U = J V     (replaces utils.py:766-775)
//   vjp    : pulls  k cotangents backward through the transposed ops W = U^T J  (replaces utils.py:790-797)
// and pb_pullback() runs the reference's subspace iteration (utils.py:756-808) on top of them, one CUDA
// graph replay per iteration, re-orthonormalising on the device (pb_ortho.cu).
//
// This file is device-agnostic: it only sequences the leaf kernels declared in pb_kernels.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "pb_kernels.h"
#include "pullback_b200.h"

#define PB_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr size_t kAlign = 256;
constexpr size_t kSplitFloats = size_t(8) << 20;   // split-K partial tiles (32 MB)

inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }
inline int round4(int v) { return (v + 3) / 4 * 4; }

// ---------------------------------------------------------------------------------------------
// plan data structures
// ---------------------------------------------------------------------------------------------
struct Val {                 // one activation tensor, [rows][C] per image (NHWC / token layout)
  long rows; int C;
  size_t p_off;              // primal copy in the primal cache (bytes)
  size_t t_off;              // k_max tangents / cotangents in the workspace (bytes)
  bool ginit;                // vjp bookkeeping: cotangent buffer already holds a contribution
  // all-fp16 tangent plan (pb_handle::t16): the tangent / cotangent of this tensor is stored as halves -- every tensor whose
  // channel count keeps 16-byte rows (C % 8 == 0); the x_t-shaped input of conv_in (and the 4-channel eps of the full plan)
  // stays fp32
  bool h16 = false;
};

enum WKind { WK_VEC, WK_RAW, WK_LIN, WK_CONV3, WK_CONV3_S2 };
struct WSpec {
  WKind kind;
  std::vector<std::string> names;   // sources, concatenated along the output dimension
  int out, in;                      // total rows, columns (conv: Co, Ci)
  size_t fwd_off, bwd_off;          // bytes in the packed region
  size_t fwd16_off = 0, bwd16_off = 0;   // fp16 copies of the two GEMM layouts (0: none)
  size_t fwd16g_off = 0;                 // ff1 of a GEGLU block: fp16 forward copy with the a / gate rows interleaved (PbGemm::gg; 0: none)
};

enum OpKind { OP_IN, OP_CONV_DIRECT, OP_GN, OP_LN, OP_GEMM, OP_CONCAT, OP_IM2COL, OP_UPSAMPLE, OP_GEGLU, OP_ATTN, OP_OUT };
struct Op {
  OpKind kind;
  int x = -1, x2 = -1, y = -1, res = -1;
  int H = 0, W = 0;
  int w = -1, bias = -1, gamma = -1, beta = -1, temb_w = -1, temb_b = -1;
  size_t bias_eff_off = 0, mean_off = 0, rstd_off = 0;
  float eps = 0.f; int silu = 0, groups = 0;
  int conv = 0;                     // OP_GEMM: 1 = 3x3 / s1 / p1 implicit GEMM
  int pad_lo = 0, Ho = 0, Wo = 0;   // OP_IM2COL
  // OP_ATTN
  int heads = 0, d = 0, cross = 0, Nq = 0, Nk = 0, ldk = 0, ldq = 0, kv = -1;
  float scale = 0.f;
  size_t P_off = 0, Pt_off = 0, Qt_off = 0, Kt_off = 0, Vt_off = 0;
  // fp16 copies for the fused linearisation kernel (p16 = 1): scaled P and P^T, the transposed primal operands, and (all-fp16
  // tangent plan) X16 = the primal score operands as halves: the [N][3C] q/k/v projection (self) or the [Nk][2C] text k/v (cross)
  int p16 = 0;
  size_t P16_off = 0, Pt16_off = 0, Qt16_off = 0, Kt16_off = 0, Vt16_off = 0, X16_off = 0;
};

}  // namespace

struct pb_handle {
  pb_unet_cfg cfg{};
  bool planned = false, bound = false, point = false;
  int H = 0, W = 0, op = 0, block_idx = 0, kmax = 0, ctx_len = 0;
  std::vector<Val> vals;
  std::vector<WSpec> wspecs;
  std::vector<Op> ops;
  std::map<std::string, int> windex;
  pb_sizes sizes{};
  size_t cache_top = 0, work_top = 0, packed_top = 0;
  // fixed primal-cache / workspace slots (byte offsets)
  size_t c_temb = 0, c_sin = 0, c_e1 = 0, c_ctx = 0;
  size_t w_s1 = 0, w_s2 = 0, w_s3 = 0, w_delta = 0, w_gn = 0;
  size_t w_V = 0, w_Vprev = 0, w_W = 0, w_U = 0, w_G = 0, w_M = 0, w_R = 0, w_sv = 0, w_met = 0, w_x = 0;
  size_t n_s1 = 0, n_s2 = 0, n_s3 = 0, n_delta = 0, n_gn = 0;   // floats per tangent
  size_t w_splitk = 0, n_splitk = 0;
  char* packed = nullptr; char* cache = nullptr; char* work = nullptr;
  int in_val = -1, out_val = -1;
  // decoder side (PB_OP_DEC): tangents enter at value dec_val (the mid-block output), produced by op dec_op; the values of the
  // encoder half that later ops read (skip connections) carry zero tangents
  int dec_val = -1, dec_op = -1, dec_H = 0, dec_W = 0;
  std::vector<int> dec_zero;
  long n_in = 0, n_out = 0;
  long n_x = 0;                             // numel(x_t) (= n_in except on the decoder side)
  // numerics policy (DESIGN.md "precision"): producers RNA-round GEMM operands to TF32 (rnd_*); GEMMs run one TF32
  // pass (0), error-compensated 3xTF32 (1) or split-weight 2xTF32 (2)
  int rnd_p = 1, rnd_t = 1, rnd_w = 1;      // primal activations / tangents / packed weights
  int prec_p = 0, prec_t = 0, prec_a = 0;   // primal GEMMs / tangent weight GEMMs / tangent attention GEMMs
  int rnd = 1;                              // rounding flag of the pass being interpreted
  int fused_min_tokens = 512;               // self-attention layers with >= this many tokens use the fused kernel
  int fuse_geglu = 1;                       // JVP of ff1 with the GEGLU linearisation in the GEMM epilogue (all-fp16 plan, device backend)
  int fuse_geglu_min_k = 0;                 // ... for layers with at least this many input channels
  int f16 = -1;                             // fp16 tangents + fp16-operand GEMMs on the tangent passes: -1 = if the backend has them
  bool t16 = false;                         // the plan stores every tangent / cotangent as halves (decided by pb_plan: f16 and an eligible geometry)
  size_t w_cvt = 0, n_cvt = 0;              // fp32 staging of an fp16 tangent for the materialised attention path (floats per tangent)
  std::string err;
  long launches = 0;
  // CUDA graph of one iteration (jvp + vjp + orthonormalise)
  void* graph = nullptr; int graph_k = 0; float graph_tol = 0.f; int use_graph = 1; bool warm = false;
  const float* graph_u = nullptr; const float* graph_s = nullptr; long graph_nodes = 0;
  // CUDA graphs of a stand-alone pb_jvp / pb_vjp at k columns (the tangent-sharded loop calls them once per iteration):
  // they run between the handle's own staging buffers, so one capture serves every caller pointer
  struct DirGraph { void* exec = nullptr; long nodes = 0; };
  std::map<int, DirGraph> jvp_graphs, vjp_graphs;
  bool warm_jvp = false, warm_vjp = false;
  // timing probes (pb_profile_begin / pb_profile_read): one event pair per contraction-kernel launch
  struct Probe { void* e0; void* e1; double flops; int kind; std::string label; int f16 = 0; };
  int probe_f16 = 0;                          // operand type of the GEMM being probed (PB_PROBE_GEMM_TF32 / _F16)
  std::string probe_label;                   // shape of the launch being probed (PB_PROFILE_DUMP)
  bool profiling = false;
  std::vector<Probe> probes;

  // Problem slots (pb_set_slots): `slots` independent problems (x_t, t, ctx) share the weights and run their k_slot tangent
  // columns as ONE batch of slots * k_slot images.  Each slot has its own primal cache (cache + slot * cache_stride); a tangent
  // buffer holds the slots back to back ([slot][column][rows][C]), so the weight GEMMs and the data-movement kernels take the
  // whole batch in one launch while the ops that read primal quantities run once per slot on their k_slot images.
  int slots = 1, slot = 0, k_slot = 0;
  size_t cache_stride = 0;
  bool pass_vjp = false;                      // element size of a tangent buffer: Val::t16 (JVP) / Val::g16 (VJP)
  std::vector<char> slot_point;               // pb_set_point done for slot i

  float* P(int v) const { return reinterpret_cast<float*>(cache + slot * cache_stride + vals[v].p_off); }
  // VJP only: cotangent buffer of val v may be another val's buffer handed over without a copy (residual fan-in, run_gemm_bwd)
  std::vector<int> alias;
  float* T(int v) const {
    const Val& a = vals[alias.empty() ? v : alias[v]];
    size_t off = a.t_off;
    if (slot) off += (size_t)slot * k_slot * a.rows * a.C * (vals[v].h16 ? 2 : 4);
    return reinterpret_cast<float*>(work + off);
  }
  bool is16(int v) const { return vals[v].h16; }
  // problem slots handled INSIDE a kernel (all-fp16 plan): images per problem and the primal stride between problems
  // (the fp32 plan runs those ops once per slot on one problem's images: k_slot = 0 there)
  int ks(int nb) const { return slots > 1 && t16 ? nb / slots : 0; }
  long pstride_f() const { return slots > 1 && t16 ? (long)(cache_stride / 4) : 0; }
  // io flags of a tangent-path kernel reading the tangent of val `vin` and writing the tangent of val `vout` (pb_kernels.h)
  int io(int vin, int vout) const { return (vin >= 0 && is16(vin) ? PB_IN_F16 : 0) | (vout >= 0 && is16(vout) ? PB_OUT_F16 : rnd); }
  float* CP(size_t off) const { return reinterpret_cast<float*>(cache + slot * cache_stride + off); }
  float* WP(size_t off) const { return reinterpret_cast<float*>(work + off); }
  float* Wf(int w) const { return reinterpret_cast<float*>(packed + wspecs[w].fwd_off); }
  float* Wb(int w) const { return reinterpret_cast<float*>(packed + wspecs[w].bwd_off); }
  void* Wf16(int w) const { return packed + wspecs[w].fwd16_off; }
  void* Wb16(int w) const { return packed + wspecs[w].bwd16_off; }
  bool use_f16() const { return f16 > 0; }
};

namespace {

float attn_pscale(int n);

int fail(pb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

// ---------------------------------------------------------------------------------------------
// planner
// ---------------------------------------------------------------------------------------------
struct Planner {
  pb_handle* h;
  std::string error;

  size_t cache_alloc(size_t floats) { size_t o = h->cache_top; h->cache_top = align_up(o + floats * 4); return o; }
  size_t work_alloc(size_t floats) { size_t o = h->work_top; h->work_top = align_up(o + floats * 4); return o; }

  int val(long rows, int C) {
    Val v{rows, C, 0, 0, false};
    v.p_off = cache_alloc((size_t)rows * C);
    v.t_off = work_alloc((size_t)rows * C * h->kmax);
    h->vals.push_back(v);
    return (int)h->vals.size() - 1;
  }
  int weight(WKind kind, std::vector<std::string> names, int out, int in) {
    std::string key = names[0] + "#" + std::to_string((int)kind) + "#" + std::to_string(names.size());
    auto it = h->windex.find(key);
    if (it != h->windex.end()) return it->second;
    WSpec s{kind, std::move(names), out, in, 0, 0};
    size_t n = 0;
    switch (kind) {
      case WK_VEC: n = (size_t)out; break;
      case WK_RAW: case WK_LIN: n = (size_t)out * in; break;
      case WK_CONV3: case WK_CONV3_S2: n = (size_t)out * in * 9; break;
    }
    s.fwd_off = h->packed_top; h->packed_top = align_up(h->packed_top + n * 4);
    if (kind == WK_LIN || kind == WK_CONV3 || kind == WK_CONV3_S2) {
      s.bwd_off = h->packed_top; h->packed_top = align_up(h->packed_top + n * 4);
      if (h->use_f16() && n % 4 == 0 && in % 8 == 0 && out % 8 == 0) {
        s.fwd16_off = h->packed_top; h->packed_top = align_up(h->packed_top + n * 2);
        s.bwd16_off = h->packed_top; h->packed_top = align_up(h->packed_top + n * 2);
      }
    }
    h->wspecs.push_back(s);
    h->windex[key] = (int)h->wspecs.size() - 1;
    return (int)h->wspecs.size() - 1;
  }
  int vec(const std::string& name, int n) { return weight(WK_VEC, {name}, n, 1); }

  Op& push(OpKind k) { h->ops.emplace_back(); h->ops.back().kind = k; return h->ops.back(); }

  int gn(int x, const std::string& p, float eps, int silu) {
    const Val vx = h->vals[x];                       // by value: val() below may reallocate h->vals
    const int G = h->cfg.norm_num_groups;
    if (vx.C % G || vx.C % 4) { error = "GroupNorm channels must divide into groups and be a multiple of 4"; return -1; }
    int y = val(vx.rows, vx.C);
    Op& o = push(OP_GN);
    o.x = x; o.y = y; o.eps = eps; o.silu = silu; o.groups = G;
    o.gamma = vec(p + ".weight", h->vals[x].C); o.beta = vec(p + ".bias", h->vals[x].C);
    o.mean_off = cache_alloc(G); o.rstd_off = cache_alloc(G);
    // the scratch of the chunked path is not monotone in the image count (chunk count x images), and the op may run on any
    // number of images up to k_max (primal pass: 1; problem slots: k_max / slots)
    for (int nb = 1; nb <= h->kmax; ++nb) h->n_gn = std::max(h->n_gn, pbk_gn_tmp_floats((int)vx.rows, vx.C, G, nb));
    return y;
  }
  int ln(int x, const std::string& p) {
    int y = val(h->vals[x].rows, h->vals[x].C);
    Op& o = push(OP_LN);
    o.x = x; o.y = y; o.eps = 1e-5f;
    o.gamma = vec(p + ".weight", h->vals[x].C); o.beta = vec(p + ".bias", h->vals[x].C);
    o.mean_off = cache_alloc(h->vals[x].rows); o.rstd_off = cache_alloc(h->vals[x].rows);
    return y;
  }
  // y = x W^T (+ bias) (+ res);  names: weight sources concatenated along the output dimension
  int linear(int x, std::vector<std::string> wnames, std::vector<std::string> bnames, int out, int res = -1) {
    const int in = h->vals[x].C;
    int y = val(h->vals[x].rows, out);
    Op& o = push(OP_GEMM);
    o.x = x; o.y = y; o.res = res;
    o.w = weight(WK_LIN, std::move(wnames), out, in);
    if (!bnames.empty()) o.bias = weight(WK_VEC, std::move(bnames), out, 1);
    return y;
  }
  int conv3(int x, const std::string& p, int out, int H, int W, int res = -1, const std::string& temb = "") {
    const int in = h->vals[x].C;
    if (in % 32) { error = "3x3 conv input channels must be a multiple of 32 (" + p + ")"; return -1; }
    int y = val(h->vals[x].rows, out);
    Op& o = push(OP_GEMM);
    o.x = x; o.y = y; o.res = res; o.conv = 1; o.H = H; o.W = W;
    o.w = weight(WK_CONV3, {p + ".weight"}, out, in);
    o.bias = vec(p + ".bias", out);
    if (!temb.empty()) {
      const int ted = h->cfg.block_out_channels[0] * 4;
      o.temb_w = weight(WK_RAW, {temb + ".weight"}, out, ted);
      o.temb_b = vec(temb + ".bias", out);
      o.bias_eff_off = cache_alloc(out);
    }
    return y;
  }
  int resnet(int x, int Cout, int H, int W, const std::string& p) {
    const int Cin = h->vals[x].C;
    int a = gn(x, p + ".norm1", h->cfg.norm_eps, 1); if (a < 0) return -1;
    int h1 = conv3(a, p + ".conv1", Cout, H, W, -1, p + ".time_emb_proj"); if (h1 < 0) return -1;
    int b = gn(h1, p + ".norm2", h->cfg.norm_eps, 1); if (b < 0) return -1;
    int sc = x;
    if (Cin != Cout) sc = linear(x, {p + ".conv_shortcut.weight"}, {p + ".conv_shortcut.bias"}, Cout);
    return conv3(b, p + ".conv2", Cout, H, W, sc);
  }
  // attention core; self: qkv is the fused [N][3C] projection; cross: q is [N][C], kv the cached [Nk][2C] text projection
  int attn(int q, int heads, int C, int cross, int Nk, int kv) {
    const long N = h->vals[q].rows;
    int y = val(N, C);
    Op& o = push(OP_ATTN);
    o.x = q; o.y = y; o.heads = heads; o.d = C / heads; o.cross = cross; o.Nq = (int)N; o.Nk = Nk; o.kv = kv;
    // leading dimensions of the probability matrices / transposed operands: 16-byte rows for fp32 AND for their fp16 copies
    o.ldk = h->use_f16() ? (Nk + 7) / 8 * 8 : round4(Nk); o.ldq = h->use_f16() ? ((int)N + 7) / 8 * 8 : round4((int)N);
    o.scale = 1.0f / std::sqrt((float)o.d);
    if (C % heads || o.d % 4) { error = "attention head dim must be a multiple of 4"; return -1; }
    o.P_off = cache_alloc((size_t)heads * N * o.ldk);
    o.Vt_off = cache_alloc((size_t)C * o.ldk);
    o.Kt_off = cache_alloc((size_t)C * o.ldk);
    if (!cross) {
      o.Pt_off = cache_alloc((size_t)heads * Nk * o.ldq);
      o.Qt_off = cache_alloc((size_t)C * o.ldq);
    }
    if (h->use_f16() && o.d % 8 == 0 && pbk_attn_lin_supported(o.d, (int)N, Nk) == nullptr) {
      // fp16 copies for the fused kernel (same geometry as the fp32 tensors): scaled P (and P^T), V^T, K^T (and Q^T), and the
      // primal score operands X16 ([N][3C] q/k/v projection, or the [Nk][2C] text k/v of a cross-attention layer)
      o.p16 = 1;
      o.P16_off = cache_alloc((size_t)heads * N * o.ldk / 2);
      o.Vt16_off = cache_alloc((size_t)C * o.ldk / 2);
      o.Kt16_off = cache_alloc((size_t)C * o.ldk / 2);
      o.X16_off = cache_alloc(cross ? (size_t)Nk * 2 * C / 2 : (size_t)N * 3 * C / 2);
      if (!cross) {
        o.Pt16_off = cache_alloc((size_t)heads * Nk * o.ldq / 2);
        o.Qt16_off = cache_alloc((size_t)C * o.ldq / 2);
      }
    }
    h->n_cvt = std::max(h->n_cvt, (size_t)N * (cross ? C : 3 * C));
    h->n_s1 = std::max(h->n_s1, (size_t)heads * N * o.ldk);
    if (!cross) h->n_s2 = std::max(h->n_s2, (size_t)heads * Nk * o.ldq);
    h->n_s3 = std::max(h->n_s3, (size_t)C * std::max(o.ldk, o.ldq));
    h->n_delta = std::max(h->n_delta, (size_t)heads * N);
    return y;
  }
  int transformer(int x, int heads, int H, int W, const std::string& p) {
    const int C = h->vals[x].C;
    const std::string tb = p + ".transformer_blocks.0";
    int n0 = gn(x, p + ".norm", 1e-6f, 0); if (n0 < 0) return -1;
    int t0 = linear(n0, {p + ".proj_in.weight"}, {p + ".proj_in.bias"}, C);
    int l1 = ln(t0, tb + ".norm1");
    int qkv = linear(l1, {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, {}, 3 * C);
    int o1 = attn(qkv, heads, C, 0, H * W, -1); if (o1 < 0) return -1;
    int t1 = linear(o1, {tb + ".attn1.to_out.0.weight"}, {tb + ".attn1.to_out.0.bias"}, C, t0);
    int l2 = ln(t1, tb + ".norm2");
    int q2 = linear(l2, {tb + ".attn2.to_q.weight"}, {}, C);
    // text keys / values: a primal-only GEMM on ctx, recorded as a weight pair and a cache slot
    int kvw = weight(WK_LIN, {tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, 2 * C, h->cfg.cross_attention_dim);
    size_t kv_off = cache_alloc((size_t)h->ctx_len * 2 * C);
    int o2 = attn(q2, heads, C, 1, h->ctx_len, kvw); if (o2 < 0) return -1;
    h->ops.back().bias_eff_off = kv_off;     // OP_ATTN (cross): location of the [Nk][2C] text projection
    int t2 = linear(o2, {tb + ".attn2.to_out.0.weight"}, {tb + ".attn2.to_out.0.bias"}, C, t1);
    int l3 = ln(t2, tb + ".norm3");
    int f1 = linear(l3, {tb + ".ff.net.0.proj.weight"}, {tb + ".ff.net.0.proj.bias"}, 8 * C);
    {
      // the JVP of ff1 can run with the GEGLU linearisation in its epilogue (the 8C-wide tangent never reaches HBM): it takes a
      // copy of the fp16 forward weight with the a / gate rows interleaved
      WSpec& ws = h->wspecs[h->ops.back().w];
      if (ws.fwd16_off && !ws.fwd16g_off && (4 * C) % 128 == 0 && pbk_gemm_geglu_supported()) {
        ws.fwd16g_off = h->packed_top; h->packed_top = align_up(h->packed_top + (size_t)8 * C * C * 2);
      }
    }
    int gg = val(h->vals[f1].rows, 4 * C);
    { Op& o = push(OP_GEGLU); o.x = f1; o.y = gg; }
    int t3 = linear(gg, {tb + ".ff.net.2.weight"}, {tb + ".ff.net.2.bias"}, C, t2);
    return linear(t3, {p + ".proj_out.weight"}, {p + ".proj_out.bias"}, C, x);
  }
  // diffusers 0.11.0 AttentionBlock (UNet2DModel)
  int attn_block(int x, int heads, int H, int W, const std::string& p) {
    const int C = h->vals[x].C;
    int n0 = gn(x, p + ".group_norm", h->cfg.norm_eps, 0); if (n0 < 0) return -1;
    int qkv = linear(n0, {p + ".query.weight", p + ".key.weight", p + ".value.weight"},
                     {p + ".query.bias", p + ".key.bias", p + ".value.bias"}, 3 * C);
    int o1 = attn(qkv, heads, C, 0, H * W, -1); if (o1 < 0) return -1;
    return linear(o1, {p + ".proj_attn.weight"}, {p + ".proj_attn.bias"}, C, x);
  }
  int downsample(int x, int H, int W, const std::string& p) {
    const int C = h->vals[x].C;
    const int pad = h->cfg.downsample_padding ? 1 : 0;
    const int Ho = pad ? (H + 2 - 3) / 2 + 1 : (H + 1 - 3) / 2 + 1;
    const int Wo = pad ? (W + 2 - 3) / 2 + 1 : (W + 1 - 3) / 2 + 1;
    int col = val((long)Ho * Wo, 9 * C);
    { Op& o = push(OP_IM2COL); o.x = x; o.y = col; o.H = H; o.W = W; o.Ho = Ho; o.Wo = Wo; o.pad_lo = pad; }
    int y = val((long)Ho * Wo, C);
    Op& o = push(OP_GEMM);
    o.x = col; o.y = y;
    o.w = weight(WK_CONV3_S2, {p + ".conv.weight"}, C, C);
    o.bias = vec(p + ".conv.bias", C);
    return y;
  }
  int upsample(int x, int H, int W, const std::string& p) {
    const int C = h->vals[x].C;
    int up = val(4L * H * W, C);
    { Op& o = push(OP_UPSAMPLE); o.x = x; o.y = up; o.H = H; o.W = W; }
    return conv3(up, p + ".conv", C, 2 * H, 2 * W);
  }
  int concat(int a, int b) {
    int y = val(h->vals[a].rows, h->vals[a].C + h->vals[b].C);
    Op& o = push(OP_CONCAT); o.x = a; o.x2 = b; o.y = y;
    return y;
  }

  bool build() {
    const pb_unet_cfg& c = h->cfg;
    const int L = c.n_levels;
    int H = h->H, W = h->W;
    const bool cond = c.kind == PB_UNET_COND;
    int x0 = val((long)H * W, c.in_channels);
    { Op& o = push(OP_IN); o.y = x0; o.H = H; o.W = W; }
    h->in_val = x0;
    int x = val((long)H * W, c.block_out_channels[0]);
    {
      Op& o = push(OP_CONV_DIRECT);
      o.x = x0; o.y = x; o.H = H; o.W = W;
      o.w = weight(WK_CONV3, {"conv_in.weight"}, c.block_out_channels[0], c.in_channels);
      o.bias = vec("conv_in.bias", c.block_out_channels[0]);
    }
    const int ted = c.block_out_channels[0] * 4;
    weight(WK_RAW, {"time_embedding.linear_1.weight"}, ted, c.block_out_channels[0]);
    vec("time_embedding.linear_1.bias", ted);
    weight(WK_RAW, {"time_embedding.linear_2.weight"}, ted, ted);
    vec("time_embedding.linear_2.bias", ted);
    struct Skip { int v, H, W; };
    std::vector<Skip> skips{{x, H, W}};
    for (int i = 0; i < L; ++i) {
      const std::string bp = "down_blocks." + std::to_string(i);
      for (int j = 0; j < c.layers_per_block; ++j) {
        x = resnet(x, c.block_out_channels[i], H, W, bp + ".resnets." + std::to_string(j));
        if (x < 0) return false;
        if (c.down_has_attn[i]) {
          const std::string ap = bp + ".attentions." + std::to_string(j);
          x = cond ? transformer(x, c.heads[i], H, W, ap) : attn_block(x, c.heads[i], H, W, ap);
          if (x < 0) return false;
        }
        skips.push_back({x, H, W});
      }
      if (i != L - 1) {
        x = downsample(x, H, W, bp + ".downsamplers.0");
        H = h->ops[h->ops.size() - 2].Ho;        // the OP_IM2COL emitted just before the GEMM
        W = h->ops[h->ops.size() - 2].Wo;
        skips.push_back({x, H, W});
      }
    }
    const int Cm = c.block_out_channels[L - 1];
    x = resnet(x, Cm, H, W, "mid_block.resnets.0"); if (x < 0) return false;
    x = cond ? transformer(x, c.heads[L - 1], H, W, "mid_block.attentions.0")
             : attn_block(x, c.heads[L - 1], H, W, "mid_block.attentions.0");
    if (x < 0) return false;
    x = resnet(x, Cm, H, W, "mid_block.resnets.1"); if (x < 0) return false;
    if (h->op == PB_OP_DEC) { h->dec_val = x; h->dec_op = (int)h->ops.size() - 1; h->dec_H = H; h->dec_W = W; }
    const bool full = h->op == PB_OP_FULL || h->op == PB_OP_DEC;
    if (h->op == PB_OP_UP || full) {
      // get_h_uncond stops at the mid block (utils.py:158-163); the unconditional up path exists for the full forward only
      if (!cond && h->op == PB_OP_UP) { error = "(op, block_idx) is not valid: get_h_uncond supports ('mid', 0) only"; return false; }
      const int last_up = full ? L - 1 : h->block_idx;
      for (int i = 0; i <= last_up; ++i) {
        const std::string bp = "up_blocks." + std::to_string(i);
        const int out_ch = c.block_out_channels[L - 1 - i];
        for (int j = 0; j <= c.layers_per_block; ++j) {
          Skip s = skips.back(); skips.pop_back();
          if (s.H != H || s.W != W) { error = "internal: skip geometry mismatch"; return false; }
          int cat = concat(x, s.v);
          x = resnet(cat, out_ch, H, W, bp + ".resnets." + std::to_string(j)); if (x < 0) return false;
          if (c.up_has_attn[i]) {
            const std::string ap = bp + ".attentions." + std::to_string(j);
            x = cond ? transformer(x, c.heads[L - 1 - i], H, W, ap) : attn_block(x, c.heads[L - 1 - i], H, W, ap);
            if (x < 0) return false;
          }
        }
        if (i != L - 1) { x = upsample(x, H, W, bp + ".upsamplers.0"); if (x < 0) return false; H *= 2; W *= 2; }
      }
    }
    if (full) {
      // eps head: GroupNorm -> SiLU -> Conv3x3(C0 -> in_channels), the thin direct conv (few output channels)
      int a = gn(x, "conv_norm_out", c.norm_eps, 1); if (a < 0) return false;
      int y = val((long)H * W, c.in_channels);
      Op& o = push(OP_CONV_DIRECT);
      o.x = a; o.y = y; o.H = H; o.W = W;
      o.w = weight(WK_CONV3, {"conv_out.weight"}, c.in_channels, h->vals[a].C);
      o.bias = vec("conv_out.bias", c.in_channels);
      x = y;
    }
    { Op& o = push(OP_OUT); o.x = x; o.H = H; o.W = W; }
    h->out_val = x;
    h->sizes.out_channels = h->vals[x].C; h->sizes.out_h = H; h->sizes.out_w = W;
    if (h->dec_val >= 0) {
      // values of the encoder half (and x_t itself) read by ops of the decoder half: their tangents are zero
      std::vector<int> producer(h->vals.size(), -1);
      for (size_t i = 0; i < h->ops.size(); ++i) if (h->ops[i].y >= 0) producer[h->ops[i].y] = (int)i;
      std::vector<char> seen(h->vals.size(), 0);
      for (size_t i = h->dec_op + 1; i < h->ops.size(); ++i)
        for (int v : {h->ops[i].x, h->ops[i].x2, h->ops[i].res})
          if (v >= 0 && v != h->dec_val && producer[v] <= h->dec_op && !seen[v]) { seen[v] = 1; h->dec_zero.push_back(v); }
    }
    return true;
  }
};

// All-fp16 tangent plan.  fp16 carries the same 10-bit mantissa as TF32 at half the bytes, and every tangent / cotangent
// tensor of an iteration is read and written through HBM / L2 by a bandwidth- or operand-delivery-bound kernel
// (profiles/r1final_launches_sd15_mid_k5.txt), so the plan stores ALL of them as halves: every weight GEMM runs kind::f16 with
// fp16 output (and fp16 residual), the elementwise linearisation kernels read and write halves (pb_lin16.cu), the fused
// attention kernel takes fp16 score operands.  Only the x_t-shaped ends stay fp32 (conv_in's 3- / 4-channel side, V, W, U).
// Eligible when every GEMM has fp16 weight copies and every tensor a GEMM / elementwise kernel touches keeps 16-byte rows.
void decide_t16(pb_handle* h) {
  h->t16 = false;
  if (!h->use_f16()) return;
  std::vector<char> thin(h->vals.size(), 0);                 // touched only by the direct conv / the NCHW <-> NHWC ends
  for (const Op& o : h->ops) {
    if (o.kind == OP_IN) thin[o.y] = 1;
    if (o.kind == OP_CONV_DIRECT && h->vals[o.y].C % 8) thin[o.y] = 1;    // eps head of the full plan (4 channels)
  }
  for (const Op& o : h->ops) {
    if (o.kind == OP_GEMM && h->wspecs[o.w].fwd16_off == 0) return;
    if (o.kind == OP_IN || o.kind == OP_OUT || o.kind == OP_CONV_DIRECT) continue;
    for (int v : {o.x, o.x2, o.y, o.res})
      if (v >= 0 && (thin[v] || h->vals[v].C % 8)) return;
    if (o.kind == OP_ATTN && ((o.d % 8) || (o.ldk % 8) || (!o.cross && o.ldq % 8))) return;
  }
  h->t16 = true;
  for (size_t v = 0; v < h->vals.size(); ++v) h->vals[v].h16 = !thin[v] && h->vals[v].C % 8 == 0;
}

// ---------------------------------------------------------------------------------------------
// interpreters
// ---------------------------------------------------------------------------------------------
#define CK(call)                                                      \
  do {                                                                \
    const char* e__ = (call);                                         \
    ++h->launches;                                                    \
    if (e__) return fail(h, PB_ECUDA, std::string(#call ": ") + e__); \
  } while (0)

// timing probe around one leaf launch (eager launches only: event records are not captured into the iteration graph)
template <class F>
const char* probed(pb_handle* h, int kind, double flops, pb_stream st, F&& launch) {
  if (!h->profiling) return launch();
  pb_handle::Probe pr{nullptr, nullptr, flops, kind, h->probe_label, h->probe_f16};
  if (const char* e = pbk_event_record(&pr.e0, st)) return e;
  const char* err = launch();
  if (const char* e = pbk_event_record(&pr.e1, st)) return e;
  h->probes.push_back(pr);
  return err;
}

// every GEMM gets the handle's split-K scratch (partial tiles of the K-split tail wave)
const char* gemm_call(pb_handle* h, PbGemm& g, pb_stream st) {
  g.ws = h->WP(h->w_splitk); g.ws_floats = (long)h->n_splitk;
  double k = 0;
  for (int s = 0; s < g.nseg; ++s) k += g.seg[s].K;
  const double flops = 2.0 * g.M * g.N * k * (g.conv ? 9.0 : (double)g.nb * g.nh);
  h->probe_f16 = g.ab_dtype == PB_GEMM_F16;
  if (h->profiling) {
    char b[160];
    snprintf(b, sizeof b, "gemm M=%d N=%d K=%d nseg=%d nb=%d nh=%d conv=%d HW=%dx%d res=%d ab16=%d d16=%d", g.M, g.N, (int)k, g.nseg, g.nb, g.nh,
             g.conv, g.H, g.W, g.R ? 1 : 0, g.ab_dtype == PB_GEMM_F16, g.d_dtype == PB_GEMM_F16);
    h->probe_label = b;
  }
  return probed(h, PB_PROBE_GEMM, flops, st, [&] { return pbk_gemm(&g, st); });
}
// fused attention linearisation: S (nseg products over the head dim) + T . C1 per (tangent, head)
const char* attn_lin_call(pb_handle* h, PbAttnLin& a, pb_stream st) {
  if (h->slots > 1 && h->t16) { a.k_slot = a.nb / h->slots; a.p_stride = (long)h->cache_stride; }   // every problem in one launch
  // algorithmic products only: the score product of the VJP-B launch (column deltas) recomputes, transposed, the O-bar V^T of
  // the VJP-A launch -- work of this implementation, not of the algorithm -- and is not counted
  const int nprod = (a.delta && a.delta_mode == 2 ? 0 : a.nseg) + 1 + (a.C2 ? 1 : 0);
  const double flops = 2.0 * a.Mr * a.Nc * (double)a.d * nprod * a.nb * a.nh;
  if (h->profiling) {
    char b[160];
    snprintf(b, sizeof b, "attn Mr=%d Nc=%d d=%d nseg=%d c2=%d nb=%d nh=%d", a.Mr, a.Nc, a.d, a.nseg, a.C2 ? 1 : 0, a.nb, a.nh);
    h->probe_label = b;
  }
  return probed(h, PB_PROBE_ATTN, flops, st, [&] { return pbk_attn_lin(&a, st); });
}

PbGemm plain_gemm(const float* A, long lda, long M, const float* B, long ldb, int N, int K, float* D, long ldd) {
  PbGemm g = pb_gemm_init();
  g.M = (int)M; g.N = N;
  g.seg[0].A = A; g.seg[0].lda = lda; g.seg[0].B = B; g.seg[0].ldb = ldb; g.seg[0].K = K;
  g.D = D; g.ldd = ldd;
  return g;
}

// y = x (*) W (+ bias) (+ res), nb images; mode 0 primal, 1 jvp
int run_gemm_fwd(pb_handle* h, const Op& o, int nb, bool primal, pb_stream st) {
  const Val& vx = h->vals[o.x]; const Val& vy = h->vals[o.y];
  const float* A = primal ? h->P(o.x) : h->T(o.x);
  float* D = primal ? h->P(o.y) : h->T(o.y);
  PbGemm g = plain_gemm(A, vx.C, vx.rows * nb, h->Wf(o.w), o.conv ? 9L * vx.C : vx.C, vy.C, vx.C, D, vy.C);
  if (o.conv) { g.conv = 1; g.H = o.H; g.W = o.W; g.nb = nb; }
  if (primal && o.bias >= 0) g.bias = o.temb_w >= 0 ? h->CP(o.bias_eff_off) : h->Wf(o.bias);
  if (o.res >= 0) { g.R = primal ? h->P(o.res) : h->T(o.res); g.ldr = vy.C; g.beta = 1.f; }
  g.round_tf32 = h->rnd;
  g.precise = primal ? h->prec_p : h->prec_t;
  if (!primal && h->t16) {                                   // halves in (tangent, weights), halves out (and residual)
    g.seg[0].B = h->Wf16(o.w); g.ab_dtype = PB_GEMM_F16; g.d_dtype = PB_GEMM_F16; g.round_tf32 = 0;
  }
  CK(gemm_call(h, g, st));
  return PB_OK;
}
// JVP of ff1 + GEGLU in one launch: d(a gelu(g)) = da gelu(g) + dg a gelu'(g) combined in the GEMM epilogue from the factor cache
bool can_fuse_geglu(const pb_handle* h, const Op& o, const Op* next) {
  static const int env_mink = getenv("PB_FUSE_GEGLU_MINK") ? atoi(getenv("PB_FUSE_GEGLU_MINK")) : -1;   // A/B switch
  const int mink = env_mink >= 0 ? env_mink : h->fuse_geglu_min_k;
  return h->fuse_geglu && h->t16 && next && o.kind == OP_GEMM && next->kind == OP_GEGLU && next->x == o.y && !o.conv && o.res < 0 &&
         h->wspecs[o.w].fwd16g_off != 0 && h->vals[o.x].C >= mink;
}
int run_gemm_geglu_jvp(pb_handle* h, const Op& o, const Op& ge, int nb, pb_stream st) {
  const Val& vx = h->vals[o.x]; const Val& vy = h->vals[o.y]; const Val& vo = h->vals[ge.y];
  PbGemm g = plain_gemm(h->T(o.x), vx.C, vx.rows * nb, reinterpret_cast<const float*>(h->packed + h->wspecs[o.w].fwd16g_off), vx.C, vy.C, vx.C,
                        h->T(ge.y), vo.C);
  g.ab_dtype = PB_GEMM_F16; g.d_dtype = PB_GEMM_F16; g.round_tf32 = 0;
  g.gg = h->P(o.y); g.gg_F = vo.C; g.gg_rows_p = vy.rows; g.gg_k_slot = h->ks(nb); g.gg_p_stride = h->pstride_f();
  const double flops = 2.0 * g.M * g.N * vx.C;
  h->probe_f16 = 1;
  if (h->profiling) {
    char b[160];
    snprintf(b, sizeof b, "gemm+geglu M=%d N=%d K=%d nseg=1 nb=1 nh=1 conv=0 HW=0x0 res=0 ab16=1 d16=1", g.M, g.N, vx.C);
    h->probe_label = b;
  }
  CK(probed(h, PB_PROBE_GEMM, flops, st, [&] { return pbk_gemm(&g, st); }));   // no split-K scratch: the epilogue is not a plain sum
  return PB_OK;
}
// gx (+)= gy (*) W^T ; gres (+)= gy
int run_gemm_bwd(pb_handle* h, const Op& o, int nb, pb_stream st) {
  Val& vx = h->vals[o.x]; const Val& vy = h->vals[o.y];
  PbGemm g = plain_gemm(h->T(o.y), vy.C, vy.rows * nb, h->Wb(o.w), o.conv ? 9L * vy.C : vy.C, vx.C, vy.C, h->T(o.x), vx.C);
  if (o.conv) { g.conv = 1; g.H = o.H; g.W = o.W; g.nb = nb; }
  if (vx.ginit) { g.R = h->T(o.x); g.ldr = vx.C; g.beta = 1.f; }
  g.round_tf32 = h->rnd;
  g.precise = h->prec_t;
  if (h->t16) { g.seg[0].B = h->Wb16(o.w); g.ab_dtype = PB_GEMM_F16; g.d_dtype = PB_GEMM_F16; g.round_tf32 = 0; }
  CK(gemm_call(h, g, st));
  vx.ginit = true;
  if (o.res >= 0) {
    Val& vr = h->vals[o.res];
    if (!vr.ginit) {
      // first contribution to the residual input: y's cotangent buffer is dead after this op, so it BECOMES the residual's
      // cotangent buffer (later contributions accumulate into it in place) instead of being copied
      h->alias[o.res] = h->alias[o.y];
    } else {
      CK(pbk_copy2d(h->T(o.res), vr.C, h->T(o.y), vy.C, vy.rows * nb, vy.C, vr.ginit ? 1.f : 0.f, h->io(o.y, o.res), st));
    }
    vr.ginit = true;
  }
  return PB_OK;
}

struct AttnPtrs { const float *Q, *K, *V; long ld; };   // primal operands, row stride

int run_attn_primal(pb_handle* h, const Op& o, const float* ctx, pb_stream st) {
  const int hd = o.heads, d = o.d, C = hd * d, N = o.Nq, Nk = o.Nk, ldk = o.ldk;
  const float *Q, *K, *V; long ldq_, ldkv;
  if (o.cross) {
    // text keys / values: kv = ctx [Nk][ctx_dim] x Wkv^T -> [Nk][2C]
    float* kv = h->CP(o.bias_eff_off);
    PbGemm g = plain_gemm(ctx, h->cfg.cross_attention_dim, Nk, h->Wf(o.kv), h->cfg.cross_attention_dim, 2 * C,
                          h->cfg.cross_attention_dim, kv, 2 * C);
    g.round_tf32 = h->rnd;
    g.precise = h->prec_p; CK(gemm_call(h, g, st));
    Q = h->P(o.x); ldq_ = C; K = kv; V = kv + C; ldkv = 2 * C;
  } else {
    Q = h->P(o.x); K = Q + C; V = Q + 2 * C; ldq_ = ldkv = 3 * C;
  }
  float* P = h->CP(o.P_off);
  {
    PbGemm g = plain_gemm(Q, ldq_, N, K, ldkv, Nk, d, P, ldk);
    g.seg[0].sAh = d; g.seg[0].sBh = d; g.sDh = (long)N * ldk; g.nh = hd; g.alpha = o.scale;
    g.precise = h->prec_p; CK(gemm_call(h, g, st));
  }
  CK(pbk_softmax_fwd(P, (long)hd * N, Nk, ldk, h->rnd, st));
  float* Vt = h->CP(o.Vt_off); float* Kt = h->CP(o.Kt_off);
  CK(pbk_transpose(Vt, ldk, 0, (long)d * ldk, V, ldkv, 0, d, 1, hd, Nk, d, 0.f, h->rnd, st));
  CK(pbk_transpose(Kt, ldk, 0, (long)d * ldk, K, ldkv, 0, d, 1, hd, Nk, d, 0.f, h->rnd, st));
  if (!o.cross) {
    CK(pbk_transpose(h->CP(o.Qt_off), o.ldq, 0, (long)d * o.ldq, Q, ldq_, 0, d, 1, hd, N, d, 0.f, h->rnd, st));
    CK(pbk_transpose(h->CP(o.Pt_off), o.ldq, 0, (long)Nk * o.ldq, P, ldk, 0, (long)N * ldk, 1, hd, N, Nk, 0.f, h->rnd, st));
  }
  if (o.p16 && h->t16) {
    // fp16 copies for the fused kernel: scaled P (and P^T, attn_pscale), the transposed primal operands, the score operands
    CK(pbk_to_f16_scaled(h->CP(o.P16_off), P, (size_t)hd * N * ldk, attn_pscale(Nk), st));
    CK(pbk_to_f16(h->CP(o.Vt16_off), Vt, (size_t)C * ldk, st));
    CK(pbk_to_f16(h->CP(o.Kt16_off), Kt, (size_t)C * ldk, st));
    if (o.cross) {
      CK(pbk_to_f16(h->CP(o.X16_off), h->CP(o.bias_eff_off), (size_t)Nk * 2 * C, st));
    } else {
      CK(pbk_to_f16(h->CP(o.X16_off), h->P(o.x), (size_t)N * 3 * C, st));
      CK(pbk_to_f16_scaled(h->CP(o.Pt16_off), h->CP(o.Pt_off), (size_t)hd * Nk * o.ldq, attn_pscale(N), st));
      CK(pbk_to_f16(h->CP(o.Qt16_off), h->CP(o.Qt_off), (size_t)C * o.ldq, st));
    }
  }
  {
    PbGemm g = plain_gemm(P, ldk, N, Vt, ldk, d, Nk, h->P(o.y), C);
    g.seg[0].sAh = (long)N * ldk; g.seg[0].sBh = (long)d * ldk; g.sDh = d; g.nh = hd; g.round_tf32 = h->rnd;
    g.precise = h->prec_p; CK(gemm_call(h, g, st));
  }
  return PB_OK;
}

// Scale of the fp16 probability copies: sqrt(row length), rounded to a power of two.  Unscaled, a near-uniform row of 4096
// entries (2.4e-4 each) puts T = P o (S - delta) into fp16's subnormal range; scaled by the full row length, a peaked row
// (P -> 1) would overflow for |S - delta| > 16.  sqrt(N) keeps uniform rows normal down to |S| ~ 4e-3 and peaked rows finite up
// to |S| ~ 1e3; the kernel divides the products by the same factor and saturates T at the fp16 maximum.
float attn_pscale(int n) { return std::exp2(std::floor(0.5f * std::log2((float)std::max(n, 1)) + 0.5f)); }

bool use_fused(const pb_handle* h, const Op& o) {
  return !o.cross && o.Nq >= h->fused_min_tokens && pbk_attn_lin_supported(o.d, o.Nq, o.Nk) == nullptr && (!h->t16 || o.p16);
}
// cross-attention (77 text keys: three 32-column steps per tile) through the same kernel: one launch instead of
// GEMM + softmax-linearisation + GEMM and no score tangent in HBM
bool use_fused_cross(const pb_handle* h, const Op& o) {
  static const bool off = getenv("PB_NO_FUSED_CROSS") != nullptr;      // A/B switch
  return !off && o.cross && o.Nq >= h->fused_min_tokens && pbk_attn_lin_supported(o.d, o.Nq, o.Nk) == nullptr && (!h->t16 || o.p16);
}

// element offset into a tangent buffer that holds floats (es = 4) or halves (es = 2)
inline float* el(const void* p, long n, int es) { return reinterpret_cast<float*>(const_cast<char*>(static_cast<const char*>(p)) + n * es); }

// problem slots inside the materialised attention path (all-fp16 plan): the primal operands of its GEMMs (batch stride 0: K, V,
// P, their transposes) are indexed by the tangent's problem, b / k_slot, so one launch covers every slot
inline void slot_gemm(const pb_handle* h, PbGemm& g) {
  if (h->slots > 1 && h->t16 && g.nb > 1) { g.k_slot = g.nb / h->slots; g.p_stride = (long)h->cache_stride; }
}

int run_attn_jvp(pb_handle* h, const Op& o, int nb, pb_stream st) {
  const int hd = o.heads, d = o.d, C = hd * d, N = o.Nq, Nk = o.Nk, ldk = o.ldk;
  float* P = h->CP(o.P_off); float* Vt = h->CP(o.Vt_off);
  float* dS = h->WP(h->w_s1);
  const long sS = (long)hd * N * ldk;
  const bool t16 = h->t16;
  const int es = t16 ? 2 : 4;                    // element size of the tangent buffers
  if (use_fused(h, o)) {
    // dO = [P o dS] V - rowsum(P o dS) o O + P dV   with dS = (dQ K^T + Q dK^T)/sqrt(d) never stored and P streamed once
    const float* qkv = t16 ? h->CP(o.X16_off) : h->P(o.x);       // primal [N][3C] (fp16 copy in the all-fp16 plan)
    const float* dqkv = h->T(o.x);
    float* dVt = h->WP(h->w_s3);
    CK(pbk_transpose(dVt, ldk, (long)C * ldk, (long)d * ldk, el(dqkv, 2 * C, es), 3 * C, (long)N * 3 * C, d, nb, hd, Nk, d, 0.f,
                     t16 ? (PB_IN_F16 | PB_OUT_F16) : h->rnd, st));
    PbAttnLin a{};
    a.Mr = N; a.Nc = Nk; a.d = d; a.nb = nb; a.nh = hd; a.nseg = 2;
    a.seg[0].A = dqkv; a.seg[0].lda = 3 * C; a.seg[0].sAb = (long)N * 3 * C; a.seg[0].sAh = d;
    a.seg[0].B = el(qkv, C, es); a.seg[0].ldb = 3 * C; a.seg[0].sBb = 0; a.seg[0].sBh = d;
    a.seg[1].A = qkv; a.seg[1].lda = 3 * C; a.seg[1].sAb = 0; a.seg[1].sAh = d;
    a.seg[1].B = el(dqkv, C, es); a.seg[1].ldb = 3 * C; a.seg[1].sBb = (long)N * 3 * C; a.seg[1].sBh = d;
    a.alpha1 = o.scale; a.alpha2 = 1.f; a.beta = 0.f;
    a.Pm = P; a.ldp = ldk; a.sPh = (long)N * ldk;
    a.want_rsum = 1; a.O = h->P(o.y); a.ldo = C;
    a.C1 = Vt; a.ldc = ldk; a.sCh = (long)d * ldk;
    a.C2 = dVt; a.ldc2 = ldk; a.sC2h = (long)d * ldk; a.sC2b = (long)C * ldk;
    if (t16) { a.p16 = a.s16 = 1; a.p_scale = attn_pscale(Nk); a.Pm = h->CP(o.P16_off); a.C1 = h->CP(o.Vt16_off); }
    a.D = h->T(o.y); a.ldd = C; a.sDb = (long)N * C;
    a.round_tf32 = h->rnd;
    CK(attn_lin_call(h, a, st));
    return PB_OK;
  }
  if (o.cross && use_fused_cross(h, o)) {
    // dO = [P o dS] V - rowsum(P o dS) o O with dS = dQ K^T / sqrt(d); the text keys / values are constants
    const float* kv = t16 ? h->CP(o.X16_off) : h->CP(o.bias_eff_off);
    PbAttnLin a{};
    a.Mr = N; a.Nc = Nk; a.d = d; a.nb = nb; a.nh = hd; a.nseg = 1;
    a.seg[0].A = h->T(o.x); a.seg[0].lda = C; a.seg[0].sAb = (long)N * C; a.seg[0].sAh = d;
    a.seg[0].B = kv; a.seg[0].ldb = 2 * C; a.seg[0].sBb = 0; a.seg[0].sBh = d;
    a.alpha1 = o.scale; a.alpha2 = 1.f;
    a.Pm = P; a.ldp = ldk; a.sPh = (long)N * ldk;
    a.want_rsum = 1; a.O = h->P(o.y); a.ldo = C;
    a.C1 = Vt; a.ldc = ldk; a.sCh = (long)d * ldk;
    if (t16) { a.p16 = a.s16 = 1; a.p_scale = attn_pscale(Nk); a.Pm = h->CP(o.P16_off); a.C1 = h->CP(o.Vt16_off); }
    a.D = h->T(o.y); a.ldd = C; a.sDb = (long)N * C;
    a.round_tf32 = h->rnd;
    CK(attn_lin_call(h, a, st));
    return PB_OK;
  }
  // materialised path (few tokens or a head dim the fused kernel does not take): fp32 / TF32 inside; in the all-fp16 plan the
  // tangent is converted once on the way in and the last product writes halves
  const long ldx = o.cross ? C : 3 * C;
  const float* dx = h->T(o.x);
  if (t16) {
    CK(pbk_to_f32(h->WP(h->w_cvt), dx, (size_t)nb * N * ldx, st));
    dx = h->WP(h->w_cvt);
  }
  if (o.cross) {
    const float* kv = h->CP(o.bias_eff_off);
    PbGemm g = plain_gemm(dx, C, N, kv, 2 * C, Nk, d, dS, ldk);
    g.seg[0].sAb = (long)N * C; g.seg[0].sAh = d; g.seg[0].sBh = d;
    g.sDb = sS; g.sDh = (long)N * ldk; g.nb = nb; g.nh = hd; g.alpha = o.scale;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  } else {
    const float* qkv = h->P(o.x); const float* dqkv = dx;
    PbGemm g = plain_gemm(dqkv, 3 * C, N, qkv + C, 3 * C, Nk, d, dS, ldk);       // dQ K^T
    g.seg[0].sAb = (long)N * 3 * C; g.seg[0].sAh = d; g.seg[0].sBh = d;
    g.nseg = 2;                                                                    // + Q dK^T
    g.seg[1].A = qkv; g.seg[1].lda = 3 * C; g.seg[1].sAb = 0; g.seg[1].sAh = d;
    g.seg[1].B = dqkv + C; g.seg[1].ldb = 3 * C; g.seg[1].sBb = (long)N * 3 * C; g.seg[1].sBh = d; g.seg[1].K = d;
    g.sDb = sS; g.sDh = (long)N * ldk; g.nb = nb; g.nh = hd; g.alpha = o.scale;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  CK(pbk_softmax_lin(P, (long)hd * N, dS, nb, Nk, ldk, h->rnd, h->ks(nb), h->pstride_f(), st));
  PbGemm g = plain_gemm(dS, ldk, N, Vt, ldk, d, Nk, h->T(o.y), C);                // dP V
  g.seg[0].sAb = sS; g.seg[0].sAh = (long)N * ldk; g.seg[0].sBh = (long)d * ldk;
  g.sDb = (long)N * C; g.sDh = d; g.nb = nb; g.nh = hd; g.round_tf32 = h->rnd;
  if (t16) { g.d_dtype = PB_GEMM_F16; g.round_tf32 = 0; }
  if (!o.cross) {                                                                  // + P dV
    float* dVt = h->WP(h->w_s3);
    CK(pbk_transpose(dVt, ldk, (long)C * ldk, (long)d * ldk, dx + 2 * C, 3 * C, (long)N * 3 * C, d, nb, hd, Nk, d, 0.f,
                     h->rnd, st));
    g.nseg = 2;
    g.seg[1].A = P; g.seg[1].lda = ldk; g.seg[1].sAb = 0; g.seg[1].sAh = (long)N * ldk;
    g.seg[1].B = dVt; g.seg[1].ldb = ldk; g.seg[1].sBb = (long)C * ldk; g.seg[1].sBh = (long)d * ldk; g.seg[1].K = Nk;
  }
  g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  return PB_OK;
}

int run_attn_vjp(pb_handle* h, const Op& o, int nb, pb_stream st) {
  const int hd = o.heads, d = o.d, C = hd * d, N = o.Nq, Nk = o.Nk, ldk = o.ldk, ldq = o.ldq;
  float* P = h->CP(o.P_off); float* Kt = h->CP(o.Kt_off);
  const float* gO = h->T(o.y);
  float* gS = h->WP(h->w_s1);
  const long sS = (long)hd * N * ldk;
  const bool t16 = h->t16;
  const int es = t16 ? 2 : 4;
  const float* V; long ldkv;
  if (o.cross) { V = h->CP(o.bias_eff_off) + C; ldkv = 2 * C; } else { V = h->P(o.x) + 2 * C; ldkv = 3 * C; }
  const bool fused_self = use_fused(h, o), fused_cross = o.cross && use_fused_cross(h, o);
  if (fused_self || fused_cross) {
    // the value operand of the score product as the fused kernel reads it: halves in the all-fp16 plan (X16: [N][3C] / [Nk][2C])
    const float* Vs = t16 ? el(h->CP(o.X16_off), o.cross ? C : 2 * C, 2) : V;
    float* gx = h->T(o.x); float* delta = h->WP(h->w_delta);
    CK(pbk_attn_delta(gO, C, h->P(o.y), C, nb, N, hd, d, delta, t16 ? PB_IN_F16 : 0, h->ks(nb), h->pstride_f(), st));   // delta = rowsum(Obar o O)
    const long ldx = o.cross ? C : 3 * C;
    PbAttnLin a{};
    // Qbar = scale * [P o (Obar V^T - delta_row)] K   (cross-attention: the only cotangent -- text keys / values are constants)
    a.Mr = N; a.Nc = Nk; a.d = d; a.nb = nb; a.nh = hd; a.nseg = 1;
    a.seg[0].A = gO; a.seg[0].lda = C; a.seg[0].sAb = (long)N * C; a.seg[0].sAh = d;
    a.seg[0].B = Vs; a.seg[0].ldb = ldkv; a.seg[0].sBb = 0; a.seg[0].sBh = d;
    a.alpha1 = 1.f; a.alpha2 = o.scale;
    a.Pm = P; a.ldp = ldk; a.sPh = (long)N * ldk;
    a.delta = delta; a.delta_mode = 1;
    a.C1 = Kt; a.ldc = ldk; a.sCh = (long)d * ldk;
    if (t16) { a.p16 = a.s16 = 1; a.p_scale = attn_pscale(Nk); a.Pm = h->CP(o.P16_off); a.C1 = h->CP(o.Kt16_off); }
    a.D = gx; a.ldd = ldx; a.sDb = (long)N * ldx;
    a.round_tf32 = h->rnd;
    CK(attn_lin_call(h, a, st));
    h->vals[o.x].ginit = true;
    if (o.cross) return PB_OK;
    // Kbar = scale * [P^T o (V Obar^T - delta_col)] Q  and  Vbar = P^T Obar  (rows = keys, columns = queries; P^T streamed once)
    float* gOt = h->WP(h->w_s3);
    float* Pt = h->CP(o.Pt_off); float* Qt = h->CP(o.Qt_off);
    CK(pbk_transpose(gOt, ldq, (long)C * ldq, (long)d * ldq, gO, C, (long)N * C, d, nb, hd, N, d, 0.f,
                     t16 ? (PB_IN_F16 | PB_OUT_F16) : h->rnd, st));
    PbAttnLin b{};
    b.Mr = Nk; b.Nc = N; b.d = d; b.nb = nb; b.nh = hd; b.nseg = 1;
    b.seg[0].A = Vs; b.seg[0].lda = ldkv; b.seg[0].sAb = 0; b.seg[0].sAh = d;
    b.seg[0].B = gO; b.seg[0].ldb = C; b.seg[0].sBb = (long)N * C; b.seg[0].sBh = d;
    b.alpha1 = 1.f; b.alpha2 = o.scale;
    b.Pm = Pt; b.ldp = ldq; b.sPh = (long)Nk * ldq;
    b.delta = delta; b.delta_mode = 2;
    b.C1 = Qt; b.ldc = ldq; b.sCh = (long)d * ldq;
    b.D = el(gx, C, es); b.ldd = 3 * C; b.sDb = (long)N * 3 * C;
    b.C2 = gOt; b.ldc2 = ldq; b.sC2h = (long)d * ldq; b.sC2b = (long)C * ldq;
    b.D2 = el(gx, 2 * C, es); b.ldd2 = 3 * C; b.sD2b = (long)N * 3 * C;
    if (t16) { b.p16 = b.s16 = 1; b.p_scale = attn_pscale(N); b.Pm = h->CP(o.Pt16_off); b.C1 = h->CP(o.Qt16_off); }
    b.round_tf32 = h->rnd;
    CK(attn_lin_call(h, b, st));
    return PB_OK;
  }
  // materialised path: fp32 / TF32 inside; all-fp16 plan: convert the cotangent once, the products into gx write halves
  if (t16) {
    CK(pbk_to_f32(h->WP(h->w_cvt), gO, (size_t)nb * N * C, st));
    gO = h->WP(h->w_cvt);
  }
  const int dd = t16 ? PB_GEMM_F16 : PB_GEMM_F32;
  {                                                                                // dP = gO V^T
    PbGemm g = plain_gemm(gO, C, N, V, ldkv, Nk, d, gS, ldk);
    g.seg[0].sAb = (long)N * C; g.seg[0].sAh = d; g.seg[0].sBh = d;
    g.sDb = sS; g.sDh = (long)N * ldk; g.nb = nb; g.nh = hd;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  CK(pbk_softmax_lin(P, (long)hd * N, gS, nb, Nk, ldk, h->rnd, h->ks(nb), h->pstride_f(), st));              // gS = P o (dP - rowsum(P o dP))
  const long ldx = o.cross ? C : 3 * C;
  float* gx = h->T(o.x);
  {                                                                                // gQ = scale gS K
    PbGemm g = plain_gemm(gS, ldk, N, Kt, ldk, d, Nk, gx, ldx);
    g.seg[0].sAb = sS; g.seg[0].sAh = (long)N * ldk; g.seg[0].sBh = (long)d * ldk;
    g.sDb = (long)N * ldx; g.sDh = d; g.nb = nb; g.nh = hd; g.alpha = o.scale; g.round_tf32 = t16 ? 0 : h->rnd; g.d_dtype = dd;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  h->vals[o.x].ginit = true;
  if (o.cross) return PB_OK;
  // keys / values need the query index contracted: work on the transposed score matrix
  float* Pt = h->CP(o.Pt_off); float* Qt = h->CP(o.Qt_off);
  float* gSt = h->WP(h->w_s2); float* delta = h->WP(h->w_delta); float* gOt = h->WP(h->w_s3);
  const long sSt = (long)hd * Nk * ldq;
  CK(pbk_attn_delta(gO, C, h->P(o.y), C, nb, N, hd, d, delta, 0, h->ks(nb), h->pstride_f(), st));
  {                                                                                // dP^T = V gO^T
    PbGemm g = plain_gemm(V, ldkv, Nk, gO, C, N, d, gSt, ldq);
    g.seg[0].sAh = d; g.seg[0].sBb = (long)N * C; g.seg[0].sBh = d;
    g.sDb = sSt; g.sDh = (long)Nk * ldq; g.nb = nb; g.nh = hd;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  CK(pbk_attn_ds(Pt, gSt, delta, 1.f, nb, hd, Nk, N, ldq, 1, h->rnd, h->ks(nb), h->pstride_f(), st));        // gS^T = P^T o (dP^T - delta_i)
  {                                                                                // gK = scale gS^T Q
    PbGemm g = plain_gemm(gSt, ldq, Nk, Qt, ldq, d, N, el(gx, C, es), ldx);
    g.seg[0].sAb = sSt; g.seg[0].sAh = (long)Nk * ldq; g.seg[0].sBh = (long)d * ldq;
    g.sDb = (long)N * ldx; g.sDh = d; g.nb = nb; g.nh = hd; g.alpha = o.scale; g.round_tf32 = t16 ? 0 : h->rnd; g.d_dtype = dd;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  CK(pbk_transpose(gOt, ldq, (long)C * ldq, (long)d * ldq, gO, C, (long)N * C, d, nb, hd, N, d, 0.f, h->rnd, st));
  {                                                                                // gV = P^T gO
    PbGemm g = plain_gemm(Pt, ldq, Nk, gOt, ldq, d, N, el(gx, 2 * C, es), ldx);
    g.seg[0].sAh = (long)Nk * ldq; g.seg[0].sBb = (long)C * ldq; g.seg[0].sBh = (long)d * ldq;
    g.sDb = (long)N * ldx; g.sDh = d; g.nb = nb; g.nh = hd; g.round_tf32 = t16 ? 0 : h->rnd; g.d_dtype = dd;
    g.precise = h->prec_a; slot_gemm(h, g); CK(gemm_call(h, g, st));
  }
  return PB_OK;
}

// first > 0 (pb_decode_from): the ops [first, end) only, on the time embedding / text projection / skip connections the last
// pb_set_point cached
int run_primal(pb_handle* h, const float* x, float t, const float* ctx, float* h_out, pb_stream st, size_t first = 0) {
  h->rnd = h->rnd_p;
  const pb_unet_cfg& c = h->cfg;
  const int c0 = c.block_out_channels[0], ted = 4 * c0;
  auto W = [&](const char* name, WKind kind, int n) { return h->Wf(h->windex.at(std::string(name) + "#" + std::to_string((int)kind) + "#" + std::to_string(n))); };
  const float* ctx_r = nullptr;
  if (first == 0) {
    CK(pbk_timestep_embedding(t, c0, c.flip_sin_to_cos, c.freq_shift, h->CP(h->c_sin), st));
    CK(pbk_gemv(W("time_embedding.linear_1.weight", WK_RAW, 1), h->CP(h->c_sin), W("time_embedding.linear_1.bias", WK_VEC, 1), ted, c0,
                0, 1, h->CP(h->c_e1), st));
    CK(pbk_gemv(W("time_embedding.linear_2.weight", WK_RAW, 1), h->CP(h->c_e1), W("time_embedding.linear_2.bias", WK_VEC, 1), ted, ted,
                0, 0, h->CP(h->c_temb), st));
    if (c.kind == PB_UNET_COND) {
      if (!ctx) return fail(h, PB_EINVAL, "encoder_hidden_states is required for a conditional U-Net");
      const size_t nctx = (size_t)h->ctx_len * c.cross_attention_dim;
      if (h->rnd) CK(pbk_round_tf32(h->CP(h->c_ctx), ctx, nctx, st));
      else CK(pbk_copy(h->CP(h->c_ctx), ctx, nctx * 4, st));
    }
  }
  if (c.kind == PB_UNET_COND) ctx_r = h->CP(h->c_ctx);
  for (size_t oi = first; oi < h->ops.size(); ++oi) {
    const Op& o = h->ops[oi];
    switch (o.kind) {
      case OP_IN: {
        const Val& v = h->vals[o.y];
        CK(pbk_transpose(h->P(o.y), v.C, 0, 0, x, v.rows, 0, 0, 1, 1, v.C, (int)v.rows, 0.f, 0, st));
        break;
      }
      case OP_CONV_DIRECT:
        CK(pbk_conv3x3_direct(h->P(o.x), 1, o.H, o.W, h->vals[o.x].C, h->Wf(o.w), h->Wf(o.bias), h->vals[o.y].C, h->P(o.y), 0.f, 0, st));
        break;
      case OP_GN: {
        const Val& v = h->vals[o.x];
        CK(pbk_gn_stats(h->P(o.x), 1, (int)v.rows, v.C, o.groups, o.eps, h->CP(o.mean_off), h->CP(o.rstd_off), h->WP(h->w_gn), st));
        CK(pbk_gn_apply(h->P(o.x), h->CP(o.mean_off), h->CP(o.rstd_off), h->Wf(o.gamma), h->Wf(o.beta), 1, (int)v.rows, v.C,
                        o.groups, o.silu, h->rnd, h->P(o.y), st));
        break;
      }
      case OP_LN: {
        const Val& v = h->vals[o.x];
        CK(pbk_ln_fwd(h->P(o.x), v.rows, v.C, h->Wf(o.gamma), h->Wf(o.beta), o.eps, h->P(o.y), h->CP(o.mean_off), h->CP(o.rstd_off),
                      h->rnd, st));
        break;
      }
      case OP_GEMM: {
        if (o.temb_w >= 0) {   // bias_eff = conv bias + time_emb_proj(SiLU(temb))
          const int Co = h->vals[o.y].C;
          CK(pbk_gemv(h->Wf(o.temb_w), h->CP(h->c_temb), h->Wf(o.temb_b), Co, ted, 1, 0, h->CP(o.bias_eff_off), st));
          CK(pbk_copy2d(h->CP(o.bias_eff_off), Co, h->Wf(o.bias), Co, 1, Co, 1.f, 0, st));
        }
        if (int e = run_gemm_fwd(h, o, 1, true, st)) return e;
        break;
      }
      case OP_CONCAT: {
        const Val& a = h->vals[o.x]; const Val& b = h->vals[o.x2]; const Val& y = h->vals[o.y];
        CK(pbk_copy2d(h->P(o.y), y.C, h->P(o.x), a.C, a.rows, a.C, 0.f, 0, st));
        CK(pbk_copy2d(h->P(o.y) + a.C, y.C, h->P(o.x2), b.C, b.rows, b.C, 0.f, 0, st));
        break;
      }
      case OP_IM2COL:
        CK(pbk_im2col_s2(h->P(o.x), 1, o.H, o.W, h->vals[o.x].C, o.pad_lo, o.Ho, o.Wo, h->P(o.y), h->rnd, st));
        break;
      case OP_UPSAMPLE:
        CK(pbk_upsample2x(h->P(o.x), 1, o.H, o.W, h->vals[o.x].C, h->P(o.y), h->rnd, st));
        break;
      case OP_GEGLU:
        // the cached ff1 output [a | g] becomes the linearisation factors [gelu(g) | a gelu'(g)] in the same pass
        CK(pbk_geglu_fwd(h->P(o.x), h->vals[o.x].rows, h->vals[o.y].C, h->P(o.y), h->rnd, 1, st));
        break;
      case OP_ATTN:
        if (int e = run_attn_primal(h, o, ctx_r, st)) return e;
        break;
      case OP_OUT:
        if (h_out) {
          const Val& v = h->vals[o.x];
          CK(pbk_transpose(h_out, v.rows, 0, 0, h->P(o.x), v.C, 0, 0, 1, 1, (int)v.rows, v.C, 0.f, 0, st));
        }
        break;
    }
  }
  return PB_OK;
}

// ops that read primal quantities (normalisation statistics, activation arguments, attention probabilities): with several
// problem slots they run once per slot; everything else (weight GEMMs, data movement) takes all slots in one launch
// ... except in the all-fp16 plan, whose kernels index the primal tensors per image (b / k_slot), the GEMMs of the
// materialised attention path included (PbGemm::k_slot)
bool per_slot_op(const pb_handle* h, const Op& o) {
  if (h->t16) return false;                  // every kernel of the all-fp16 plan indexes the primal tensors per image (b / k_slot)
  return o.kind == OP_GN || o.kind == OP_LN || o.kind == OP_GEGLU || o.kind == OP_ATTN;
}

int jvp_op(pb_handle* h, const Op& o, const float* V, int nb, float* U, pb_stream st) {
  {
    switch (o.kind) {
      case OP_IN: {
        const Val& v = h->vals[o.y];
        CK(pbk_transpose(h->T(o.y), v.C, v.rows * v.C, 0, V, v.rows, v.rows * v.C, 0, nb, 1, v.C, (int)v.rows, 0.f, 0, st));
        break;
      }
      case OP_CONV_DIRECT:
        CK(pbk_conv3x3_direct(h->T(o.x), nb, o.H, o.W, h->vals[o.x].C, h->Wf(o.w), nullptr, h->vals[o.y].C, h->T(o.y), 0.f,
                              h->io(o.x, o.y) & ~1, st));
        break;
      case OP_GN: {
        const Val& v = h->vals[o.x];
        CK(pbk_gn_lin(h->P(o.x), h->CP(o.mean_off), h->CP(o.rstd_off), h->Wf(o.gamma), h->Wf(o.beta), (int)v.rows, v.C, o.groups,
                      o.silu, h->T(o.x), nb, 0, h->T(o.y), 0.f, h->io(o.x, o.y), h->WP(h->w_gn), h->ks(nb), h->pstride_f(), st));
        break;
      }
      case OP_LN: {
        const Val& v = h->vals[o.x];
        CK(pbk_ln_lin(h->P(o.x), h->CP(o.mean_off), h->CP(o.rstd_off), h->Wf(o.gamma), v.rows, v.C, h->T(o.x), nb, 0, h->T(o.y), 0.f,
                      h->io(o.x, o.y), h->ks(nb), h->pstride_f(), st));
        break;
      }
      case OP_GEMM:
        if (int e = run_gemm_fwd(h, o, nb, false, st)) return e;
        break;
      case OP_CONCAT: {
        const Val& a = h->vals[o.x]; const Val& b = h->vals[o.x2]; const Val& y = h->vals[o.y];
        const int f = h->t16 ? (PB_IN_F16 | PB_OUT_F16) : 0, es = h->t16 ? 2 : 4;
        CK(pbk_copy2d(h->T(o.y), y.C, h->T(o.x), a.C, a.rows * nb, a.C, 0.f, f, st));
        CK(pbk_copy2d(el(h->T(o.y), a.C, es), y.C, h->T(o.x2), b.C, b.rows * nb, b.C, 0.f, f, st));
        break;
      }
      case OP_IM2COL:
        CK(pbk_im2col_s2(h->T(o.x), nb, o.H, o.W, h->vals[o.x].C, o.pad_lo, o.Ho, o.Wo, h->T(o.y), h->io(o.x, o.y), st));
        break;
      case OP_UPSAMPLE:
        CK(pbk_upsample2x(h->T(o.x), nb, o.H, o.W, h->vals[o.x].C, h->T(o.y), h->io(o.x, o.y), st));
        break;
      case OP_GEGLU:
        CK(pbk_geglu_jvp(h->P(o.x), h->vals[o.x].rows, h->T(o.x), nb, h->vals[o.y].C, h->T(o.y), h->io(o.x, o.y), h->ks(nb),
                         h->pstride_f(), st));
        break;
      case OP_ATTN:
        if (int e = run_attn_jvp(h, o, nb, st)) return e;
        break;
      case OP_OUT: {
        const Val& v = h->vals[o.x];
        CK(pbk_transpose(U, v.rows, v.rows * v.C, 0, h->T(o.x), v.C, v.rows * v.C, 0, nb, 1, (int)v.rows, v.C, 0.f,
                         h->is16(o.x) ? PB_IN_F16 : 0, st));
        break;
      }
    }
  }
  return PB_OK;
}

struct SlotGuard { pb_handle* h; ~SlotGuard() { h->slot = 0; } };

int run_jvp(pb_handle* h, const float* V, int nb, float* U, pb_stream st) {
  h->rnd = h->rnd_t; h->pass_vjp = false; h->k_slot = nb / h->slots;
  SlotGuard guard{h};
  size_t first = 0;
  if (h->dec_val >= 0) {
    // decoder side: V is h-shaped and enters at the mid-block output; the skip connections carry zero tangents (re-zeroed every
    // time: the transpose pass accumulates cotangents into the same buffers)
    for (int v : h->dec_zero) {
      const Val& a = h->vals[v];
      CK(pbk_memset0(h->T(v), (size_t)nb * a.rows * a.C * (h->is16(v) ? 2 : 4), st));
    }
    const Val& v = h->vals[h->dec_val];
    CK(pbk_transpose(h->T(h->dec_val), v.C, v.rows * v.C, 0, V, v.rows, v.rows * v.C, 0, nb, 1, v.C, (int)v.rows, 0.f,
                     h->is16(h->dec_val) ? PB_OUT_F16 : h->rnd, st));
    first = (size_t)h->dec_op + 1;
  }
  for (size_t oi = first; oi < h->ops.size(); ++oi) {
    const Op& o = h->ops[oi];
    if (can_fuse_geglu(h, o, oi + 1 < h->ops.size() ? &h->ops[oi + 1] : nullptr)) {
      if (int e = run_gemm_geglu_jvp(h, o, h->ops[oi + 1], nb, st)) return e;
      ++oi;                                     // the GEGLU op ran in the GEMM's epilogue
      continue;
    }
    if (h->slots > 1 && per_slot_op(h, o)) {
      for (int p = 0; p < h->slots; ++p) {
        h->slot = p;
        if (int e = jvp_op(h, o, V, h->k_slot, U, st)) return e;
      }
      h->slot = 0;
    } else if (int e = jvp_op(h, o, V, nb, U, st)) return e;
  }
  return PB_OK;
}

int vjp_op(pb_handle* h, const Op& o, const float* U, int nb, float* Wout, pb_stream st) {
  {
    switch (o.kind) {
      case OP_OUT: {
        Val& v = h->vals[o.x];
        CK(pbk_transpose(h->T(o.x), v.C, v.rows * v.C, 0, U, v.rows, v.rows * v.C, 0, nb, 1, v.C, (int)v.rows, 0.f, h->io(-1, o.x), st));
        v.ginit = true;
        break;
      }
      case OP_ATTN:
        if (int e = run_attn_vjp(h, o, nb, st)) return e;
        break;
      case OP_GEGLU:
        if (h->vals[o.x].ginit) return fail(h, PB_ESTATE, "internal: GEGLU input has several consumers");
        CK(pbk_geglu_vjp(h->P(o.x), h->vals[o.x].rows, h->T(o.y), nb, h->vals[o.y].C, h->T(o.x), h->io(o.y, o.x), h->ks(nb),
                         h->pstride_f(), st));
        h->vals[o.x].ginit = true;
        break;
      case OP_UPSAMPLE: {
        Val& v = h->vals[o.x];
        CK(pbk_upsample2x_vjp(h->T(o.y), nb, o.H, o.W, v.C, h->T(o.x), v.ginit ? 1.f : 0.f, h->io(o.y, o.x), st));
        v.ginit = true;
        break;
      }
      case OP_IM2COL: {
        Val& v = h->vals[o.x];
        CK(pbk_col2im_s2(h->T(o.y), nb, o.H, o.W, v.C, o.pad_lo, o.Ho, o.Wo, h->T(o.x), v.ginit ? 1.f : 0.f, h->io(o.y, o.x), st));
        v.ginit = true;
        break;
      }
      case OP_CONCAT: {
        Val& a = h->vals[o.x]; Val& b = h->vals[o.x2]; const Val& y = h->vals[o.y];
        const int f = h->t16 ? (PB_IN_F16 | PB_OUT_F16) : h->rnd, es = h->t16 ? 2 : 4;
        CK(pbk_copy2d(h->T(o.x), a.C, h->T(o.y), y.C, a.rows * nb, a.C, a.ginit ? 1.f : 0.f, f, st));
        CK(pbk_copy2d(h->T(o.x2), b.C, el(h->T(o.y), a.C, es), y.C, b.rows * nb, b.C, b.ginit ? 1.f : 0.f, f, st));
        a.ginit = b.ginit = true;
        break;
      }
      case OP_GEMM:
        if (int e = run_gemm_bwd(h, o, nb, st)) return e;
        break;
      case OP_LN: {
        Val& v = h->vals[o.x];
        CK(pbk_ln_lin(h->P(o.x), h->CP(o.mean_off), h->CP(o.rstd_off), h->Wf(o.gamma), v.rows, v.C, h->T(o.y), nb, 1, h->T(o.x),
                      v.ginit ? 1.f : 0.f, h->io(o.y, o.x), h->ks(nb), h->pstride_f(), st));
        v.ginit = true;
        break;
      }
      case OP_GN: {
        Val& v = h->vals[o.x];
        CK(pbk_gn_lin(h->P(o.x), h->CP(o.mean_off), h->CP(o.rstd_off), h->Wf(o.gamma), h->Wf(o.beta), (int)v.rows, v.C, o.groups,
                      o.silu, h->T(o.y), nb, 1, h->T(o.x), v.ginit ? 1.f : 0.f, h->io(o.y, o.x), h->WP(h->w_gn), h->ks(nb), h->pstride_f(), st));
        v.ginit = true;
        break;
      }
      case OP_CONV_DIRECT: {
        Val& v = h->vals[o.x];
        CK(pbk_conv3x3_direct(h->T(o.y), nb, o.H, o.W, h->vals[o.y].C, h->Wb(o.w), nullptr, v.C, h->T(o.x), 0.f, h->io(o.y, o.x) & ~1, st));
        v.ginit = true;
        break;
      }
      case OP_IN: {
        const Val& v = h->vals[o.y];
        CK(pbk_transpose(Wout, v.rows, v.rows * v.C, 0, h->T(o.y), v.C, v.rows * v.C, 0, nb, 1, (int)v.rows, v.C, 0.f, 0, st));
        break;
      }
    }
  }
  return PB_OK;
}

int run_vjp(pb_handle* h, const float* U, int nb, float* Wout, pb_stream st) {
  h->rnd = h->rnd_t; h->pass_vjp = true; h->k_slot = nb / h->slots;
  SlotGuard guard{h};
  for (Val& v : h->vals) v.ginit = false;
  h->alias.resize(h->vals.size());
  for (size_t i = 0; i < h->alias.size(); ++i) h->alias[i] = (int)i;
  struct Unalias { pb_handle* h; ~Unalias() { h->alias.clear(); } } unalias{h};       // the JVP / primal see every val's own buffer
  const size_t stop = h->dec_val >= 0 ? (size_t)h->dec_op + 1 : 0;        // decoder side: the walk ends at the mid-block output
  for (size_t i = h->ops.size(); i-- > stop;) {
    const Op& o = h->ops[i];
    if (o.y >= 0 && !h->vals[o.y].ginit) return fail(h, PB_ESTATE, "internal: cotangent consumed before it was produced");
    if (h->slots > 1 && per_slot_op(h, o)) {
      const bool had = h->vals[o.x].ginit;               // every slot sees the bookkeeping state the op started from
      for (int p = 0; p < h->slots; ++p) {
        h->slot = p;
        h->vals[o.x].ginit = had;
        if (int e = vjp_op(h, o, U, h->k_slot, Wout, st)) return e;
      }
      h->slot = 0;
    } else if (int e = vjp_op(h, o, U, nb, Wout, st)) return e;
  }
  if (h->dec_val >= 0) {
    const Val& v = h->vals[h->dec_val];
    if (!v.ginit) return fail(h, PB_ESTATE, "internal: the decoder half never reads h");
    CK(pbk_transpose(Wout, v.rows, v.rows * v.C, 0, h->T(h->dec_val), v.C, v.rows * v.C, 0, nb, 1, (int)v.rows, v.C, 0.f,
                     h->is16(h->dec_val) ? PB_IN_F16 : 0, st));
  }
  return PB_OK;
}

int run_ortho(pb_handle* h, const float* Wm, const float* Vprev, int k, float atol, float* V, float* s, float* metrics, pb_stream st) {
  double* G = reinterpret_cast<double*>(h->work + h->w_G);
  double* M = reinterpret_cast<double*>(h->work + h->w_M);
  CK(pbk_gram2(Wm, Vprev, k, h->n_in, G, Vprev ? M : nullptr, st));
  CK(pbk_jacobi(G, Vprev ? M : nullptr, k, h->WP(h->w_R), s, st));
  CK(pbk_rotate(Wm, h->WP(h->w_R), Vprev, k, h->n_in, atol, 1e-5f, V, metrics, st));
  return PB_OK;
}

int check_ready(pb_handle* h, int k, bool need_point) {
  if (!h) return PB_EINVAL;
  if (!h->planned) return fail(h, PB_ESTATE, "pb_plan has not been called");
  if (!h->bound) return fail(h, PB_ESTATE, "pb_bind_weights has not been called");
  if (need_point && !h->point) return fail(h, PB_ESTATE, h->slots > 1 ? "pb_set_point has not been called for every problem slot"
                                                                       : "pb_set_point has not been called");
  if (k < 1 || k > h->kmax) return fail(h, PB_EINVAL, "pca_rank must be in [1, k_max]");
  if (need_point && h->slots > 1 && k % h->slots) return fail(h, PB_EINVAL, "the column count must be a multiple of the problem slots");
  return PB_OK;
}

void drop_graph(pb_handle* h) {
  if (h->graph) { pbk_graph_destroy(h->graph); h->graph = nullptr; }
  for (auto* m : {&h->jvp_graphs, &h->vjp_graphs}) {
    for (auto& kv : *m) if (kv.second.exec) pbk_graph_destroy(kv.second.exec);
    m->clear();
  }
}

// Stand-alone J V / U^T J through a cached CUDA graph: user buffers are copied to / from the handle's staging buffers
// (k x n_in and k x n_out floats: a few MB) so that the captured pointers never change.
template <class Run>
int run_direction(pb_handle* h, std::map<int, pb_handle::DirGraph>& graphs, bool& warm, int k, const float* in, size_t in_bytes,
                  float* stage_in, float* stage_out, float* out, size_t out_bytes, pb_stream stream, Run&& run) {
  if (!h->use_graph || h->profiling) return run(in, out);
  CK(pbk_copy(stage_in, in, in_bytes, stream));
  pb_handle::DirGraph& g = graphs[k];
  if (!g.exec && warm) {
    if (pbk_graph_begin(stream) == nullptr) {
      const long before = h->launches;
      const int rc = run(stage_in, stage_out);
      const char* e2 = pbk_graph_end(stream, &g.exec, &g.nodes);
      h->launches = before;
      if (rc) { drop_graph(h); return rc; }
      if (e2) { drop_graph(h); return fail(h, PB_ECUDA, std::string("graph capture: ") + e2); }
    } else {
      h->use_graph = 0;                    // backend without graph support: launch directly
      return run(in, out);
    }
  }
  if (g.exec) {
    if (const char* e = pbk_graph_launch(g.exec, stream)) return fail(h, PB_ECUDA, std::string("graph launch: ") + e);
    h->launches += g.nodes;
  } else {
    if (int e = run(stage_in, stage_out)) return e;          // the first call ever runs eagerly (one-time kernel attribute setup)
    warm = true;
  }
  CK(pbk_copy(out, stage_out, out_bytes, stream));
  return PB_OK;
}

}  // namespace

// ==================================================================================================
// C ABI
// ==================================================================================================
PB_API const char* pb_backend(void) { return pbk_backend_name(); }

PB_API const char* pb_last_error(const pb_handle* h) { return h ? h->err.c_str() : "null handle"; }

PB_API int pb_create(const pb_unet_cfg* cfg, pb_handle** out) {
  if (!cfg || !out) return PB_EINVAL;
  *out = nullptr;
  if (cfg->n_levels < 1 || cfg->n_levels > PB_MAX_LEVELS || cfg->in_channels < 1 || cfg->layers_per_block < 1 ||
      cfg->norm_num_groups < 1 || (cfg->kind != PB_UNET_COND && cfg->kind != PB_UNET_UNCOND))
    return PB_EINVAL;
  for (int i = 0; i < cfg->n_levels; ++i)
    if (cfg->block_out_channels[i] < 4 || cfg->heads[i] < 1) return PB_EINVAL;
  pb_handle* h = new pb_handle();
  h->cfg = *cfg;
  *out = h;
  return PB_OK;
}

PB_API void pb_destroy(pb_handle* h) {
  if (!h) return;
  drop_graph(h);
  for (auto& p : h->probes) { pbk_event_destroy(p.e0); pbk_event_destroy(p.e1); }
  delete h;
}

PB_API int64_t pb_kernel_launches(const pb_handle* h) { return h ? h->launches : 0; }

PB_API int pb_plan_summary(const pb_handle* h, pb_plan_info* info) {
  if (!h || !info) return PB_EINVAL;
  if (!h->planned) return PB_ESTATE;
  *info = pb_plan_info{};
  info->n_ops = (int32_t)h->ops.size();
  for (const Op& o : h->ops) {
    if (o.kind == OP_GEMM) {
      ++info->n_gemm;
      info->n_conv3x3 += o.conv ? 1 : 0;
      info->n_gemm_f16_jvp += h->t16 ? 1 : 0;                // all-fp16 tangent plan: every GEMM of both passes, halves in and out
      info->n_gemm_f16_vjp_stored += h->t16 ? 1 : 0;
      info->n_gemm_d16_jvp += h->t16 ? 1 : 0;
    } else if (o.kind == OP_ATTN) {
      ++info->n_attn;
      info->n_attn_fused_self += use_fused(h, o) ? 1 : 0;
      info->n_attn_fused_cross += use_fused_cross(h, o) ? 1 : 0;
      info->n_attn_p16 += ((use_fused(h, o) || use_fused_cross(h, o)) && o.p16 && h->t16) ? 1 : 0;
    }
  }
  return PB_OK;
}

PB_API int pb_ddim_step(const float* x, const float* eps, float a_t, float a_next, float* x_next, float* pred_x0, int64_t n,
                        void* stream) {
  if (!x || !eps || !x_next || n < 0) return PB_EINVAL;
  return pbk_ddim_step(x, eps, a_t, a_next, x_next, pred_x0, (long)n, stream) ? PB_EINVAL : PB_OK;
}

PB_API int pb_lincomb3(float* out, float a, const float* x, float b, const float* y, float c, const float* z, int64_t n, void* stream) {
  if (!out || !x || n < 0) return PB_EINVAL;
  return pbk_lincomb3(out, a, x, b, y, c, z, (long)n, stream) ? PB_ECUDA : PB_OK;
}

PB_API int pb_profile_begin(pb_handle* h) {
  if (!h) return PB_EINVAL;
  for (auto& p : h->probes) { pbk_event_destroy(p.e0); pbk_event_destroy(p.e1); }
  h->probes.clear();
  h->profiling = true;
  return PB_OK;
}
PB_API int pb_profile_read(pb_handle* h, int32_t kind, double* ms, double* flops, int64_t* launches) {
  if (!h || !ms || !flops || !launches) return PB_EINVAL;
  h->profiling = false;
  *ms = 0; *flops = 0; *launches = 0;
  // PB_PROBE_GEMM_TF32 / PB_PROBE_GEMM_F16: the GEMM launches of one operand type (each has its own tensor-pipe peak)
  const int base = (kind == PB_PROBE_GEMM_TF32 || kind == PB_PROBE_GEMM_F16) ? PB_PROBE_GEMM : kind;
  for (const auto& p : h->probes) {
    if (p.kind != base) continue;
    if (kind == PB_PROBE_GEMM_TF32 && p.f16) continue;
    if (kind == PB_PROBE_GEMM_F16 && !p.f16) continue;
    const float t = pbk_event_elapsed_ms(p.e0, p.e1);
    if (t < 0.f) return fail(h, PB_ECUDA, "event timing failed");
    *ms += t; *flops += p.flops; ++*launches;
  }
  // PB_PROFILE_DUMP=<path>: one line per probed launch (us, algorithmic GF, shape) for per-shape tables (profiles/)
  if (const char* path = (kind == base ? getenv("PB_PROFILE_DUMP") : nullptr)) {
    if (FILE* f = fopen(path, kind == PB_PROBE_GEMM ? "w" : "a")) {
      for (const auto& p : h->probes)
        if (p.kind == kind) fprintf(f, "%.2f us %.3f GF %s\n", 1e3 * pbk_event_elapsed_ms(p.e0, p.e1), p.flops * 1e-9, p.label.c_str());
      fclose(f);
    }
  }
  return PB_OK;
}

// Enumerates the state_dict entries the planned path consumes (name + PyTorch shape), in plan order.
PB_API int pb_weight_count(const pb_handle* h) {
  if (!h || !h->planned) return 0;
  int n = 0;
  for (const WSpec& s : h->wspecs) n += (int)s.names.size();
  return n;
}
PB_API int pb_weight_info(const pb_handle* h, int32_t index, const char** name, int32_t* ndim, int64_t* shape) {
  if (!h || !h->planned || !name || !ndim || !shape) return PB_EINVAL;
  for (const WSpec& s : h->wspecs) {
    if (index >= (int)s.names.size()) { index -= (int)s.names.size(); continue; }
    *name = s.names[index].c_str();
    const int64_t rows = s.out / (int64_t)s.names.size();
    shape[0] = rows; shape[1] = s.in; shape[2] = shape[3] = 3;
    switch (s.kind) {
      case WK_VEC: *ndim = 1; break;
      case WK_RAW: case WK_LIN: *ndim = 2; break;
      case WK_CONV3: case WK_CONV3_S2: *ndim = 4; break;
    }
    return PB_OK;
  }
  return PB_EINVAL;
}

PB_API int pb_set_option(pb_handle* h, const char* name, int value) {
  if (!h || !name) return PB_EINVAL;
  struct { const char* n; int* p; bool rebind; } opts[] = {
      {"round_primal", &h->rnd_p, false}, {"round_tangent", &h->rnd_t, false}, {"round_weights", &h->rnd_w, true},
      {"precise_primal", &h->prec_p, false}, {"precise_tangent", &h->prec_t, false}, {"precise_attn", &h->prec_a, false},
      {"fused_min_tokens", &h->fused_min_tokens, false}, {"fuse_geglu", &h->fuse_geglu, false}};
  if (!strcmp(name, "f16_operands")) {          // changes the plan (fp16 weight copies, operand dtypes): plan again
    if (value && !pbk_has_f16_operands()) return fail(h, PB_EINVAL, "this backend has no fp16-operand GEMMs");
    h->f16 = value ? 1 : 0; h->planned = h->bound = h->point = false; drop_graph(h);
    return PB_OK;
  }
  for (auto& o : opts)
    if (!strcmp(name, o.n)) {
      *o.p = value; drop_graph(h); h->point = false;
      if (o.rebind) h->bound = false;
      return PB_OK;
    }
  if (!strcmp(name, "use_graph")) { h->use_graph = value ? 1 : 0; drop_graph(h); return PB_OK; }
  return fail(h, PB_EINVAL, std::string("unknown option ") + name);
}

PB_API int pb_plan(pb_handle* h, int32_t height, int32_t width, int32_t op, int32_t block_idx, int32_t k_max, int32_t ctx_len,
                   pb_sizes* sizes) {
  if (!h) return PB_EINVAL;
  h->planned = h->bound = h->point = false;
  drop_graph(h);
  if (height < 1 || width < 1 || k_max < 1 || k_max > 64) return fail(h, PB_EINVAL, "bad geometry or k_max (1..64)");
  // the reference raises ValueError for every other (op, block_idx) (utils.py:527, :158-163); op='down' is broken there
  if (op == PB_OP_MID) { if (block_idx != 0) return fail(h, PB_EINVAL, "(op, block_idx) is not valid"); }
  else if (op == PB_OP_UP) {
    if (h->cfg.kind != PB_UNET_COND || block_idx < 0 || block_idx >= h->cfg.n_levels) return fail(h, PB_EINVAL, "(op, block_idx) is not valid");
  } else if (op == PB_OP_FULL) {
    if (block_idx != 0) return fail(h, PB_EINVAL, "(op, block_idx) is not valid");
  } else if (op == PB_OP_DEC) {
    // get_h_to_e asserts op in ['mid', 'down'] and substitutes h after the mid block only (utils.py:544, :603-606)
    if (block_idx != 0 || h->cfg.kind != PB_UNET_COND) return fail(h, PB_EINVAL, "(op, block_idx) is not valid");
  } else return fail(h, PB_EINVAL, "(op, block_idx) is not valid");
  if (h->cfg.kind == PB_UNET_COND && ctx_len < 1) return fail(h, PB_EINVAL, "ctx_len must be >= 1 for a conditional U-Net");
  h->H = height; h->W = width; h->op = op; h->block_idx = block_idx; h->kmax = k_max; h->ctx_len = ctx_len;
  h->vals.clear(); h->ops.clear(); h->wspecs.clear(); h->windex.clear();
  h->dec_val = h->dec_op = -1; h->dec_zero.clear();
  h->cache_top = h->work_top = h->packed_top = 0;
  h->n_s1 = h->n_s2 = h->n_s3 = h->n_delta = h->n_gn = 0; h->n_cvt = 0;
  h->sizes = pb_sizes{};
  if (h->f16 < 0) h->f16 = pbk_has_f16_operands() ? 1 : 0;
  Planner p{h, ""};
  const int c0 = h->cfg.block_out_channels[0];
  h->c_sin = p.cache_alloc(c0); h->c_e1 = p.cache_alloc(4 * c0); h->c_temb = p.cache_alloc(4 * c0);
  h->c_ctx = p.cache_alloc((size_t)std::max(1, ctx_len) * std::max(1, h->cfg.cross_attention_dim));
  if (!p.build()) return fail(h, PB_EINVAL, p.error);
  h->n_x = (long)h->cfg.in_channels * height * width;
  h->n_in = h->dec_val >= 0 ? h->vals[h->dec_val].rows * h->vals[h->dec_val].C : h->n_x;
  h->n_out = h->vals[h->out_val].rows * h->vals[h->out_val].C;
  decide_t16(h);
  const size_t K = k_max;
  h->w_s1 = p.work_alloc(h->n_s1 * K); h->w_s2 = p.work_alloc(h->n_s2 * K); h->w_s3 = p.work_alloc(h->n_s3 * K);
  h->w_delta = p.work_alloc(h->n_delta * K); h->w_gn = p.work_alloc(h->n_gn + 64);
  h->w_V = p.work_alloc(K * h->n_in); h->w_Vprev = p.work_alloc(K * h->n_in); h->w_W = p.work_alloc(K * h->n_in);
  h->w_U = p.work_alloc(K * h->n_out);
  h->w_G = p.work_alloc(2 * K * K); h->w_M = p.work_alloc(2 * K * K); h->w_R = p.work_alloc(K * K);
  h->w_sv = p.work_alloc(K); h->w_met = p.work_alloc(4 + 2 * K);      // (dist^2, not-close count) per problem slot
  h->w_x = p.work_alloc((size_t)h->n_x);
  h->n_splitk = kSplitFloats; h->w_splitk = p.work_alloc(kSplitFloats);
  h->w_cvt = p.work_alloc(h->t16 ? h->n_cvt * K : 0);
  h->sizes.packed_weight_bytes = h->packed_top; h->sizes.primal_cache_bytes = h->cache_top; h->sizes.workspace_bytes = h->work_top;
  h->sizes.n_in = h->n_in; h->sizes.n_out = h->n_out; h->sizes.n_x = h->n_x;
  if (h->dec_val >= 0) {
    h->sizes.in_channels = h->vals[h->dec_val].C; h->sizes.in_h = h->dec_H; h->sizes.in_w = h->dec_W;
  } else { h->sizes.in_channels = h->cfg.in_channels; h->sizes.in_h = height; h->sizes.in_w = width; }
  if (sizes) *sizes = h->sizes;
  h->planned = true;
  return PB_OK;
}

PB_API int pb_bind_weights(pb_handle* h, const pb_tensor_desc* table, int32_t n, void* packed, void* stream) {
  if (!h || !table || !packed) return PB_EINVAL;
  if (!h->planned) return fail(h, PB_ESTATE, "pb_plan has not been called");
  h->bound = false; h->point = false;
  drop_graph(h);
  std::map<std::string, const pb_tensor_desc*> byname;
  for (int i = 0; i < n; ++i) if (table[i].name) byname[table[i].name] = &table[i];
  h->packed = static_cast<char*>(packed);
  for (const WSpec& s : h->wspecs) {
    int row = 0;
    for (const std::string& name : s.names) {
      auto it = byname.find(name);
      if (it == byname.end()) return fail(h, PB_EMISSING, "state_dict entry missing: " + name);
      const pb_tensor_desc& d = *it->second;
      long numel = 1;
      for (int a = 0; a < d.ndim; ++a) numel *= d.shape[a];
      const long per_row = s.kind == WK_VEC ? 1 : (s.kind == WK_CONV3 || s.kind == WK_CONV3_S2) ? (long)s.in * 9 : s.in;
      if (numel % per_row || (s.names.size() == 1 && numel != per_row * s.out))
        return fail(h, PB_EINVAL, "unexpected shape for " + name);
      const int rows = (int)(numel / per_row);
      if (row + rows > s.out) return fail(h, PB_EINVAL, "unexpected shape for " + name);
      {
        // the PyTorch shape pb_weight_info reports: [rows] | [rows][in] | [rows][in][3][3] (a transposed or re-shaped tensor
        // with the right element count must not be accepted silently)
        const int want_nd = s.kind == WK_VEC ? 1 : (s.kind == WK_CONV3 || s.kind == WK_CONV3_S2) ? 4 : 2;
        // (a 1x1 convolution [rows][in][1][1] is the same matrix as a Linear weight: SD-1.x proj_in / proj_out / conv_shortcut)
        const bool conv1x1 = want_nd == 2 && d.ndim == 4 && d.shape[2] == 1 && d.shape[3] == 1;
        bool ok = (d.ndim == want_nd || conv1x1) && d.shape[0] == rows && rows == s.out / (int)s.names.size();
        if (ok && want_nd >= 2) ok = d.shape[1] == s.in;
        if (ok && want_nd == 4) ok = d.shape[2] == 3 && d.shape[3] == 3;
        if (!ok) return fail(h, PB_EINVAL, "unexpected shape for " + name);
      }
      float* fwd = reinterpret_cast<float*>(h->packed + s.fwd_off);
      float* bwd = reinterpret_cast<float*>(h->packed + s.bwd_off);
      switch (s.kind) {
        case WK_VEC: case WK_RAW:
          CK(pbk_copy(fwd + (size_t)row * per_row, d.data, (size_t)numel * 4, stream));
          break;
        case WK_LIN:
          if (h->rnd_w) CK(pbk_round_tf32(fwd + (size_t)row * s.in, d.data, (size_t)numel, stream));
          else CK(pbk_copy(fwd + (size_t)row * s.in, d.data, (size_t)numel * 4, stream));
          CK(pbk_transpose(bwd + row, s.out, 0, 0, d.data, s.in, 0, 0, 1, 1, rows, s.in, 0.f, h->rnd_w, stream));
          break;
        case WK_CONV3:
          // conv_in (thin direct kernel, fp32 FMA) keeps full precision
          CK(pbk_pack_conv3x3(d.data, s.out, s.in, fwd, bwd, (s.in % 32 == 0 && s.out % 32 == 0) ? h->rnd_w : 0, stream));
          break;
        case WK_CONV3_S2:
          CK(pbk_pack_conv3x3(d.data, s.out, s.in, fwd, nullptr, h->rnd_w, stream));
          CK(pbk_transpose(bwd, s.out, 0, 0, fwd, 9L * s.in, 0, 0, 1, 1, s.out, 9 * s.in, 0.f, 0, stream));
          break;
      }
      row += rows;
    }
    if (row != s.out) return fail(h, PB_EINVAL, "unexpected shape for " + s.names[0]);
    if (s.fwd16_off) {
      const size_t n = (size_t)s.out * s.in * ((s.kind == WK_CONV3 || s.kind == WK_CONV3_S2) ? 9 : 1);
      CK(pbk_to_f16(h->packed + s.fwd16_off, reinterpret_cast<const float*>(h->packed + s.fwd_off), n, stream));
      CK(pbk_to_f16(h->packed + s.bwd16_off, reinterpret_cast<const float*>(h->packed + s.bwd_off), n, stream));
      if (s.fwd16g_off) CK(pbk_interleave_rows16(h->packed + s.fwd16g_off, h->packed + s.fwd16_off, s.out / 2, s.in, stream));
    }
  }
  h->bound = true;
  return PB_OK;
}

// Problem slots: `slots` independent problems share this handle's weights and run their tangent columns as one batch (see
// pb_handle::slots).  k_max of pb_plan must be a multiple of `slots` (k_max / slots columns per problem); the primal cache
// handed to pb_set_point is `slots` caches of cache_stride_bytes (>= pb_sizes::primal_cache_bytes, 256-byte multiple) each.
PB_API int pb_set_slots(pb_handle* h, int32_t slots, size_t cache_stride_bytes) {
  if (!h) return PB_EINVAL;
  if (!h->planned) return fail(h, PB_ESTATE, "pb_plan has not been called");
  if (slots < 1 || h->kmax % slots) return fail(h, PB_EINVAL, "slots must divide k_max");
  if (slots > 1 && (cache_stride_bytes < h->sizes.primal_cache_bytes || cache_stride_bytes % 256))
    return fail(h, PB_EINVAL, "cache_stride_bytes must be a 256-byte multiple >= primal_cache_bytes");
  drop_graph(h);
  h->slots = slots; h->slot = 0; h->cache_stride = slots > 1 ? cache_stride_bytes : 0;
  h->slot_point.assign(slots, 0); h->point = false;
  return PB_OK;
}
// The slot the next pb_set_point fills (0 .. slots - 1).
PB_API int pb_select_slot(pb_handle* h, int32_t slot) {
  if (!h) return PB_EINVAL;
  if (slot < 0 || slot >= h->slots) return fail(h, PB_EINVAL, "slot out of range");
  h->slot = slot;
  return PB_OK;
}

PB_API int pb_set_point(pb_handle* h, const float* x, float t, const float* ctx, void* primal_cache, void* workspace, float* h_out,
                        void* stream) {
  if (int e = check_ready(h, 1, false)) return e;
  if (!x || !primal_cache || !workspace) return fail(h, PB_EINVAL, "null pointer");
  const int slot = h->slot;                       // chosen by pb_select_slot; back to 0 afterwards
  SlotGuard guard{h};
  h->point = false;
  if (h->cache != primal_cache || h->work != workspace) {
    drop_graph(h);
    h->slot_point.assign(h->slots, 0);
  }
  h->cache = static_cast<char*>(primal_cache); h->work = static_cast<char*>(workspace);
  if ((int)h->slot_point.size() != h->slots) h->slot_point.assign(h->slots, 0);
  h->slot_point[slot] = 0;
  if (int e = run_primal(h, x, t, ctx, h_out, stream)) return e;
  h->slot_point[slot] = 1;
  h->point = std::all_of(h->slot_point.begin(), h->slot_point.end(), [](char c) { return c != 0; });
  return PB_OK;
}

// Decoder side (PB_OP_DEC), the NONLINEAR map: replaces the cached mid-block output by `h_in` ([C][H][W] like get_h returns it) and
// re-runs the decoder half on the skip connections of the last pb_set_point: eps_out = get_h_to_e(x_t, t, ctx, input_h = h_in)
// (utils.py:529-635).  The linearisation point moves to h_in with it.
PB_API int pb_decode_from(pb_handle* h, const float* h_in, float* eps_out, void* stream) {
  if (int e = check_ready(h, 1, true)) return e;
  if (h->dec_val < 0) return fail(h, PB_ESTATE, "pb_decode_from needs a plan with op = PB_OP_DEC");
  if (h->slots > 1) return fail(h, PB_ESTATE, "pb_decode_from does not take problem slots");
  if (!h_in) return fail(h, PB_EINVAL, "null pointer");
  drop_graph(h);
  h->rnd = h->rnd_p;
  const Val& v = h->vals[h->dec_val];
  // [C][rows] -> [rows][C], rounded like the kernel that normally writes this value (its consumers are GEMM operands)
  if (const char* e = pbk_transpose(h->P(h->dec_val), v.C, 0, 0, h_in, v.rows, 0, 0, 1, 1, v.C, (int)v.rows, 0.f, h->rnd, stream))
    return fail(h, PB_ECUDA, e);
  return run_primal(h, nullptr, 0.f, nullptr, eps_out, stream, (size_t)h->dec_op + 1);
}

PB_API int pb_jvp(pb_handle* h, const float* V, int32_t k, float* U, void* stream) {
  if (int e = check_ready(h, k, true)) return e;
  if (!V || !U) return fail(h, PB_EINVAL, "null pointer");
  return run_direction(h, h->jvp_graphs, h->warm_jvp, k, V, (size_t)k * h->n_in * 4, h->WP(h->w_V), h->WP(h->w_U), U,
                       (size_t)k * h->n_out * 4, stream, [&](const float* a, float* b) { return run_jvp(h, a, k, b, stream); });
}

PB_API int pb_vjp(pb_handle* h, const float* U, int32_t k, float* W, void* stream) {
  if (int e = check_ready(h, k, true)) return e;
  if (!U || !W) return fail(h, PB_EINVAL, "null pointer");
  return run_direction(h, h->vjp_graphs, h->warm_vjp, k, U, (size_t)k * h->n_out * 4, h->WP(h->w_U), h->WP(h->w_W), W,
                       (size_t)k * h->n_in * 4, stream, [&](const float* a, float* b) { return run_vjp(h, a, k, b, stream); });
}

PB_API int pb_orthonormalize(pb_handle* h, const float* W, const float* Vprev, int32_t k, float atol, float* V, float* s, float* metrics,
                             void* stream) {
  if (int e = check_ready(h, k, true)) return e;
  if (!W || !V || !s) return fail(h, PB_EINVAL, "null pointer");
  return run_ortho(h, W, Vprev, k, atol, V, s, metrics, stream);
}

PB_API int pb_pullback(pb_handle* h, const float* V0, int32_t k, int32_t min_iter, int32_t max_iter, float tol, float* u, float* s,
                       float* vT, pb_iter_info* info, void* stream) {
  // with problem slots, k is the rank PER PROBLEM: V0 / vT are [slots][k][n_in], u is [slots][k][n_out], s is [slots][k]; every
  // problem runs the same number of iterations (the early exit needs all of them converged at the same check)
  if (!h) return PB_EINVAL;
  const int P = h->slots, kt = k * P;
  if (int e = check_ready(h, kt, true)) return e;
  if (!V0 || !u || !s || !vT) return fail(h, PB_EINVAL, "null pointer");
  if (max_iter < 1) return fail(h, PB_EINVAL, "max_iter must be >= 1");
  if (P > 64) return fail(h, PB_EINVAL, "at most 64 problem slots");
  float* Va = h->WP(h->w_V); float* Vb = h->WP(h->w_Vprev);
  float* Wm = h->WP(h->w_W); float* met = h->WP(h->w_met);
  const size_t vbytes = (size_t)kt * h->n_in * 4;
  CK(pbk_copy(Vb, V0, vbytes, stream));
  // One iteration (utils.py:758-799): Vprev = Vb -> U = J Vprev -> W = U^T J -> (s, Va) = svd(W); then Vb <- Va.
  // the iteration (and the CUDA graph captured from it) works on handle-owned buffers; the caller's u / s receive one copy at the
  // end, so a caller that hands over fresh output tensors for every problem re-uses the graph instead of re-capturing it
  float* ui = h->WP(h->w_U); float* si = h->WP(h->w_sv);
  auto iteration = [&]() -> int {
    if (int e = run_jvp(h, Vb, kt, ui, stream)) return e;
    if (int e = run_vjp(h, ui, kt, Wm, stream)) return e;
    for (int p = 0; p < P; ++p) {
      const size_t o = (size_t)p * k * h->n_in;
      if (int e = run_ortho(h, Wm + o, Vb + o, k, tol, Va + o, si + (size_t)p * k, met + 2 * p, stream)) return e;
    }
    return PB_OK;
  };
  int done = 0, converged = 0;
  float host_met[128] = {0.f, 0.f};
  for (int i = 0; i < max_iter; ++i) {
    bool replayed = false;
    if (h->use_graph && !h->profiling) {
      if (h->graph && (h->graph_k != k || h->graph_tol != tol)) drop_graph(h);
      if (!h->graph && h->warm) {     // the first iteration ever runs eagerly (one-time kernel attribute setup)
        const char* e = pbk_graph_begin(stream);
        if (!e) {
          const long before = h->launches;
          int rc = iteration();
          long nodes = 0;
          const char* e2 = pbk_graph_end(stream, &h->graph, &nodes);
          h->launches = before;
          if (rc) { drop_graph(h); return rc; }
          if (e2) { drop_graph(h); return fail(h, PB_ECUDA, std::string("graph capture: ") + e2); }
          h->graph_k = k; h->graph_tol = tol;
          h->graph_nodes = nodes;
        } else {
          h->use_graph = 0;            // backend without graph support: launch directly
        }
      }
      if (h->graph) {
        const char* e = pbk_graph_launch(h->graph, stream);
        if (e) return fail(h, PB_ECUDA, std::string("graph launch: ") + e);
        h->launches += h->graph_nodes;
        replayed = true;
      }
    }
    if (!replayed) { if (int e = iteration()) return e; h->warm = true; }
    ++done;
    const bool last = (i + 1 == max_iter);
    // the reference tests allclose(v_prev, v, atol) and i > min_iter after every iteration (utils.py:806-808)
    if (i > min_iter || last) {
      CK(pbk_download(host_met, met, sizeof(float) * 2 * P, stream));
      bool all_close = true;
      for (int p = 0; p < P; ++p) all_close = all_close && host_met[2 * p + 1] == 0.f;
      if (i > min_iter && all_close) { converged = 1; }
    }
    if (converged || last) break;
    CK(pbk_copy(Vb, Va, vbytes, stream));
  }
  CK(pbk_copy(vT, Va, vbytes, stream));
  if (u != ui) CK(pbk_copy(u, ui, (size_t)kt * h->n_out * 4, stream));
  if (s != si) CK(pbk_copy(s, si, (size_t)kt * 4, stream));
  if (info) { info->iters_done = done; info->converged = converged; info->last_dist = std::sqrt(host_met[0]); }
  return PB_OK;
}

// The same call with HOST buffers (what a host-language binding of the reference's method would hand over):
// copies x_t / ctx / V0 in, runs the primal pass and the iteration, copies (u, s, vT) out.  Blocking.
// With problem slots: x [slots][n_in], t [slots], ctx [slots][ctx_len][dim], V0 / vT [slots][k][n_in], u [slots][k][n_out], s [slots][k].
static int pullback_host_impl(pb_handle* h, const float* x_host, const float* t_host, const float* ctx_host, const float* V0_host, int32_t k,
                              int32_t min_iter, int32_t max_iter, float tol, float* u_host, float* s_host, float* vT_host,
                              pb_iter_info* info, void* stream) {
  if (!h) return PB_EINVAL;
  const int P = h->slots;
  const size_t kt = (size_t)k * P;
  if (int e = check_ready(h, (int)kt, false)) return e;
  if (!h->cache || !h->work) return fail(h, PB_ESTATE, "pb_set_point must have been called once to attach the cache / workspace");
  if (!x_host || !t_host || !V0_host || !u_host || !s_host || !vT_host) return fail(h, PB_EINVAL, "null pointer");
  if (h->cfg.kind == PB_UNET_COND && !ctx_host) return fail(h, PB_EINVAL, "encoder_hidden_states is required for a conditional U-Net");
  const size_t ctx_floats = (size_t)h->ctx_len * h->cfg.cross_attention_dim;
  if (h->cfg.kind == PB_UNET_COND && ctx_floats > h->n_s3 * (size_t)h->kmax)
    return fail(h, PB_ESTATE, "internal: scratch too small for the text context");
  h->point = false;
  h->slot_point.assign(P, 0);
  {
    SlotGuard guard{h};
    for (int p = 0; p < P; ++p) {
      // staged in the x / S3 scratch (S3 is free until the first attention op runs; run_primal copies the context into the cache
      // first); the next slot's upload is stream-ordered behind this slot's primal pass
      float* dx = h->WP(h->w_x);
      float* dctx = nullptr;
      CK(pbk_upload(dx, x_host + (size_t)p * h->n_in, (size_t)h->n_in * 4, stream));
      if (h->cfg.kind == PB_UNET_COND) {
        dctx = h->WP(h->w_s3);
        CK(pbk_upload(dctx, ctx_host + (size_t)p * ctx_floats, ctx_floats * 4, stream));
      }
      h->slot = p;
      if (int e = run_primal(h, dx, t_host[p], dctx, nullptr, stream)) return e;
      h->slot_point[p] = 1;
    }
  }
  h->point = true;
  float* dV0 = h->WP(h->w_W);          // W is overwritten only after V0 has been copied to Vprev
  CK(pbk_upload(dV0, V0_host, kt * h->n_in * 4, stream));
  float* du = h->WP(h->w_U); float* ds = h->WP(h->w_sv);
  float* dvT_slot = h->WP(h->w_Vprev); // Vprev is dead once the last iteration has produced V
  if (int e = pb_pullback(h, dV0, k, min_iter, max_iter, tol, du, ds, dvT_slot, info, stream)) return e;
  CK(pbk_download(u_host, du, kt * h->n_out * 4, stream));
  CK(pbk_download(s_host, ds, kt * 4, stream));
  CK(pbk_download(vT_host, dvT_slot, kt * h->n_in * 4, stream));
  return PB_OK;
}

PB_API int pb_pullback_host(pb_handle* h, const float* x_host, float t, const float* ctx_host, const float* V0_host, int32_t k,
                            int32_t min_iter, int32_t max_iter, float tol, float* u_host, float* s_host, float* vT_host,
                            pb_iter_info* info, void* stream) {
  if (!h) return PB_EINVAL;
  if (h->slots > 1) return fail(h, PB_ESTATE, "pb_pullback_host runs one problem per call: use pb_pullback_host_slots");
  return pullback_host_impl(h, x_host, &t, ctx_host, V0_host, k, min_iter, max_iter, tol, u_host, s_host, vT_host, info, stream);
}

// Host-buffer entry for a handle with problem slots: one call solves `slots` problems (see pullback_host_impl for the layouts).
PB_API int pb_pullback_host_slots(pb_handle* h, const float* x_host, const float* t_host, const float* ctx_host, const float* V0_host,
                                  int32_t k, int32_t min_iter, int32_t max_iter, float tol, float* u_host, float* s_host,
                                  float* vT_host, pb_iter_info* info, void* stream) {
  if (!h) return PB_EINVAL;
  return pullback_host_impl(h, x_host, t_host, ctx_host, V0_host, k, min_iter, max_iter, tol, u_host, s_host, vT_host, info, stream);
}
