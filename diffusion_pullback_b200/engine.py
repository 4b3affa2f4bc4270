"""Thin Python owner of one `pb_handle` (include/pullback_b200.h): allocates the three device regions the
library asks for (packed weights, primal cache, workspace) as torch byte tensors, hands raw pointers and
the CUDA stream across the C ABI, and turns status codes into the exceptions the reference raises
(`ValueError` for a bad (op, block_idx), `src/utils/utils.py:527`).  PyTorch is plumbing here: memory,
streams, nothing else."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N

PB_OP = {"mid": 0, "up": 1, "full": 2, "dec": 3}   # "full": the whole U-Net, x_t -> eps; "dec": the decoder side h -> eps (block_idx 0)
_ERRORS = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: KeyError}


def unet_config(unet) -> dict:
    """Architecture subset of a diffusers 0.11.0 `UNet2DConditionModel` / `UNet2DModel` (`unet.config`),
    or of any module tree exposing the same fields as `unet.cfg`."""
    cfg = getattr(unet, "config", None)
    if cfg is None:
        cfg = getattr(unet, "cfg", None)
    if cfg is None:
        raise ValueError("unet has neither .config (diffusers) nor .cfg")
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    boc = list(get("block_out_channels"))
    down = list(get("down_block_types"))
    cond = hasattr(unet, "up_blocks") and any("CrossAttn" in t for t in down + list(get("up_block_types", [])))
    L = len(boc)
    ahd = get("attention_head_dim")
    if cond:
        heads = list(ahd) if isinstance(ahd, (list, tuple)) else [int(ahd)] * L      # SD: "head dim" = #heads
        up = list(get("up_block_types"))
    else:
        heads = [1 if ahd is None else c // int(ahd) for c in boc]
        # UNet2DModel: `up_block_types` when the config has it, else the mirror of the down path (an AttnUpBlock2D where the
        # down path has an AttnDownBlock2D); only the full forward (op='full') reads it
        up = list(get("up_block_types", None) or ["AttnUpBlock2D" if "Attn" in t else "UpBlock2D" for t in reversed(down)])
    return dict(kind=0 if cond else 1, in_channels=int(get("in_channels")), block_out_channels=boc,
                down_has_attn=[int("Attn" in t) for t in down], up_has_attn=[int("Attn" in t) for t in up] + [0] * (L - len(up)),
                heads=heads, layers_per_block=int(get("layers_per_block")),
                cross_attention_dim=int(get("cross_attention_dim", 0) or 0) if cond else 0,
                norm_num_groups=int(get("norm_num_groups")), norm_eps=float(get("norm_eps")),
                flip_sin_to_cos=int(bool(get("flip_sin_to_cos"))), freq_shift=float(get("freq_shift")),
                downsample_padding=int(get("downsample_padding")))


class PullbackEngine:
    """One planned problem geometry: (U-Net, latent H x W, (op, block_idx), k_max, ctx_len) on one device."""

    def __init__(self, cfg: dict, height: int, width: int, op: str, block_idx: int, k_max: int, ctx_len: int,
                 device, _lib=None):
        self.L = _lib if _lib is not None else N.lib()
        if _lib is not None:
            N._declare(self.L)
        self.device = torch.device(device)
        self.cfg = cfg
        c = N.PbUnetCfg()
        c.kind, c.in_channels, c.n_levels = cfg["kind"], cfg["in_channels"], len(cfg["block_out_channels"])
        for i, v in enumerate(cfg["block_out_channels"]):
            c.block_out_channels[i] = v
            c.down_has_attn[i] = cfg["down_has_attn"][i]
            c.up_has_attn[i] = cfg["up_has_attn"][i]
            c.heads[i] = cfg["heads"][i]
        c.layers_per_block, c.cross_attention_dim = cfg["layers_per_block"], cfg["cross_attention_dim"]
        c.norm_num_groups, c.norm_eps = cfg["norm_num_groups"], cfg["norm_eps"]
        c.flip_sin_to_cos, c.freq_shift, c.downsample_padding = cfg["flip_sin_to_cos"], cfg["freq_shift"], cfg["downsample_padding"]
        self.h = C.c_void_p()
        if self.L.pb_create(C.byref(c), C.byref(self.h)) != 0:
            raise ValueError("pb_create: invalid U-Net configuration")
        if op not in PB_OP:
            # the reference raises for everything but 'mid' / 'up' ('down' is broken there: SURVEY.md s.2)
            raise ValueError(f"(op, block_idx) = ({op, block_idx}) is not valid")
        self.sizes = N.PbSizes()
        self._ck(self.L.pb_plan(self.h, height, width, PB_OP[op], int(block_idx), int(k_max), int(ctx_len), C.byref(self.sizes)))
        self.k_max, self.height, self.width, self.ctx_len = int(k_max), height, width, int(ctx_len)
        self.n_in, self.n_out, self.n_x = int(self.sizes.n_in), int(self.sizes.n_out), int(self.sizes.n_x)
        self.h_shape = (int(self.sizes.out_channels), int(self.sizes.out_h), int(self.sizes.out_w))
        self.in_shape = (int(self.sizes.in_channels), int(self.sizes.in_h), int(self.sizes.in_w))
        z = lambda n: torch.zeros(max(int(n), 16), dtype=torch.uint8, device=self.device)
        self.packed, self.cache, self.work = z(self.sizes.packed_weight_bytes), z(self.sizes.primal_cache_bytes), z(self.sizes.workspace_bytes)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.bound = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.pb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers ----
    def _ck(self, rc):
        if rc != 0:
            msg = self.L.pb_last_error(self.h).decode()
            g, self._guard = getattr(self, "_guard", None), None
            if g is not None:
                g.__exit__(None, None, None)
            raise _ERRORS.get(rc, RuntimeError)(msg)

    def _st(self):
        return C.c_void_p(self.stream.cuda_stream) if self.stream is not None else C.c_void_p(0)

    def _enter(self):
        # the C side launches on the CURRENT device (kernel attributes, tensor maps, streams): make it this engine's
        # device for the duration of the call, whatever the caller's current device is
        if self.stream is not None:
            self._guard = torch.cuda.device(self.device)
            self._guard.__enter__()
            self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def _exit(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            g, self._guard = self._guard, None
            if g is not None:
                g.__exit__(None, None, None)

    def _f32(self, t, shape=None):
        t = t.detach().to(self.device, torch.float32).contiguous()
        return t if shape is None else t.reshape(shape)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def set_option(self, name: str, value: int):
        self._ck(self.L.pb_set_option(self.h, name.encode(), int(value)))

    def profile_begin(self):
        """Start bracketing every contraction-kernel launch with an event pair (eager launches, no graph replay)."""
        self._ck(self.L.pb_profile_begin(self.h))

    def profile_read(self):
        """Stop probing; {kernel class: (device ms, algorithmic flops, launches)} since profile_begin()."""
        out = {}
        for name, kind in (("gemm_tc_kernel", 0), ("attn_lin_kernel", 1), ("gemm_tc_kernel[kind::tf32]", 2), ("gemm_tc_kernel[kind::f16]", 3)):
            ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
            self._ck(self.L.pb_profile_read(self.h, kind, C.byref(ms), C.byref(fl), C.byref(n)))
            out[name] = (ms.value, fl.value, n.value)
        return out

    @property
    def launches(self) -> int:
        return int(self.L.pb_kernel_launches(self.h))

    def weight_specs(self):
        """[(state_dict key, shape)] the planned path consumes (pb_weight_info)."""
        out = []
        for i in range(self.L.pb_weight_count(self.h)):
            name, nd, shp = C.c_char_p(), C.c_int32(), (C.c_int64 * 4)()
            self._ck(self.L.pb_weight_info(self.h, i, C.byref(name), C.byref(nd), shp))
            out.append((name.value.decode(), tuple(int(shp[j]) for j in range(nd.value))))
        return out

    # ---- C ABI calls ----
    def bind(self, state_dict):
        keep, descs = [], []
        for name, w in state_dict.items():
            if not torch.is_tensor(w) or not w.dtype.is_floating_point or w.dim() > 4:
                continue
            w = self._f32(w)
            keep.append(w)
            d = N.PbTensorDesc()
            d.name, d.data, d.ndim = name.encode(), w.data_ptr(), w.dim()
            for i, s in enumerate(w.shape):
                d.shape[i] = s
            descs.append(d)
        arr = (N.PbTensorDesc * len(descs))(*descs)
        self._enter()
        self._ck(self.L.pb_bind_weights(self.h, arr, len(descs), self._p(self.packed), self._st()))
        self._exit()
        if self.stream is not None:
            self.stream.synchronize()          # `keep` may be freed after this
        self.bound = True

    def set_slots(self, slots: int):
        """Throughput mode (pb_set_slots): `slots` independent problems share the weights and run their k_max / slots tangent
        columns each as one batch.  Re-allocates the primal cache (one per slot)."""
        slots = int(slots)
        stride = (int(self.sizes.primal_cache_bytes) + 255) // 256 * 256
        self._ck(self.L.pb_set_slots(self.h, slots, stride if slots > 1 else 0))
        self.slots = slots
        self.cache = torch.zeros(max(stride * slots, 16), dtype=torch.uint8, device=self.device)

    def set_point(self, x, t, ctx=None, want_h=False, slot=0):
        if slot:
            self._ck(self.L.pb_select_slot(self.h, int(slot)))
        x = self._f32(x, (-1,))
        assert x.numel() == self.n_x, "x_t shape does not match the planned geometry"
        if ctx is not None:
            ctx = self._f32(ctx, (-1,))
            assert ctx.numel() == self.ctx_len * self.cfg["cross_attention_dim"], "encoder_hidden_states shape mismatch"
        h_out = torch.empty((1,) + self.h_shape, device=self.device) if want_h else None
        self._enter()
        self._ck(self.L.pb_set_point(self.h, self._p(x), float(t), self._p(ctx), self._p(self.cache), self._p(self.work),
                                     self._p(h_out), self._st()))
        self._exit()
        self._keep = (x, ctx)
        return h_out

    def decode_from(self, h_in, want_eps=True):
        """Decoder side (op='dec'): replace the cached mid-block output by `h_in` [1, C, H, W] and re-run the decoder half on the
        skip connections of the last `set_point` (`get_h_to_e`, utils.py:529-635); returns eps [1, c, H, W]."""
        h_in = self._f32(h_in, (-1,))
        assert h_in.numel() == self.n_in, "h shape does not match the planned geometry"
        eps = torch.empty((1,) + self.h_shape, device=self.device) if want_eps else None
        self._enter()
        self._ck(self.L.pb_decode_from(self.h, self._p(h_in), self._p(eps), self._st()))
        self._exit()
        return eps

    def jvp(self, V):
        V = self._f32(V, (-1, self.n_in))
        U = torch.empty(V.shape[0], self.n_out, device=self.device)
        self._enter()
        self._ck(self.L.pb_jvp(self.h, self._p(V), V.shape[0], self._p(U), self._st()))
        self._exit()
        return U

    def vjp(self, U):
        U = self._f32(U, (-1, self.n_out))
        W = torch.empty(U.shape[0], self.n_in, device=self.device)
        self._enter()
        self._ck(self.L.pb_vjp(self.h, self._p(U), U.shape[0], self._p(W), self._st()))
        self._exit()
        return W

    def orthonormalize(self, W, Vprev=None, atol=0.0):
        W = self._f32(W, (-1, self.n_in))
        k = W.shape[0]
        Vprev = self._f32(Vprev, (k, self.n_in)) if Vprev is not None else None
        V, s, met = torch.empty_like(W), torch.empty(k, device=self.device), torch.zeros(2, device=self.device)
        self._enter()
        self._ck(self.L.pb_orthonormalize(self.h, self._p(W), self._p(Vprev), k, float(atol), self._p(V), self._p(s), self._p(met),
                                          self._st()))
        self._exit()
        return s, V, met

    def pullback(self, V0, min_iter, max_iter, tol):
        """utils.py:756-808 on the device: returns (u [k, n_out], s [k], vT [k, n_in], info).  With problem slots V0 is
        [slots * k, n_in] (slot-major) and so are the results."""
        V0 = self._f32(V0, (-1, self.n_in))
        kt = V0.shape[0]
        k = kt // getattr(self, "slots", 1)
        u = torch.empty(kt, self.n_out, device=self.device)
        s = torch.empty(kt, device=self.device)
        vT = torch.empty(kt, self.n_in, device=self.device)
        info = N.PbIterInfo()
        self._enter()
        self._ck(self.L.pb_pullback(self.h, self._p(V0), k, int(min_iter), int(max_iter), float(tol), self._p(u), self._p(s),
                                    self._p(vT), C.byref(info), self._st()))
        self._exit()
        return u, s, vT, info

    def pullback_host(self, x, t, ctx, V0, min_iter, max_iter, tol, out=None):
        """Host-buffer entry (pb_pullback_host): pinned/pageable CPU tensors in, CPU tensors out."""
        k = V0.shape[0]
        assert x.device.type == "cpu" and V0.device.type == "cpu"
        if out is None:
            pin = self.device.type == "cuda"
            out = (torch.empty(k, self.n_out, pin_memory=pin), torch.empty(k, pin_memory=pin), torch.empty(k, self.n_in, pin_memory=pin))
        u, s, vT = out
        info = N.PbIterInfo()
        P = getattr(self, "slots", 1)
        self._enter()
        if P > 1:
            # problem slots: x [P, n_in], t P values, ctx [P, L, D], V0 [P * k, n_in] (slot-major), contiguous fp32 CPU tensors
            th = torch.as_tensor(t, dtype=torch.float32).reshape(P).contiguous()
            self._ck(self.L.pb_pullback_host_slots(self.h, self._p(x), self._p(th), self._p(ctx), self._p(V0), k // P, int(min_iter),
                                                   int(max_iter), float(tol), self._p(u), self._p(s), self._p(vT), C.byref(info),
                                                   self._st()))
            self._exit()
            return u, s, vT, info
        self._ck(self.L.pb_pullback_host(self.h, self._p(x), float(t), self._p(ctx), self._p(V0), k, int(min_iter), int(max_iter),
                                         float(tol), self._p(u), self._p(s), self._p(vT), C.byref(info), self._st()))
        self._exit()
        return u, s, vT, info
