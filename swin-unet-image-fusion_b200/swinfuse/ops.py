"""Tensor-level wrappers over the C ABI (include/swinfuse.h).

Feature maps are logical (B, C, H, W) fp32 CUDA tensors whose *memory* is channels-last
(NHWC) -- the layout every kernel works in.  ``as_fmap`` converts anything else with the
library's own transpose kernel.  PyTorch is used for device memory, streams and autograd
bookkeeping only; all arithmetic happens in libswinfuse.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SF_PREC_BF16, SF_PREC_FP32, SwinFuseError, check

Tensor = torch.Tensor

_PRECISIONS = {"fp32": SF_PREC_FP32, "bf16": SF_PREC_BF16, SF_PREC_FP32: SF_PREC_FP32, SF_PREC_BF16: SF_PREC_BF16}
_default_precision = SF_PREC_FP32


def set_default_precision(p) -> None:
    """'fp32' (FFMA, <=1e-4 rel.) or 'bf16' (tensor cores, <=2e-2 rel.)."""
    global _default_precision
    _default_precision = _PRECISIONS[p]


def get_default_precision() -> int:
    return _default_precision


def _prec(p) -> int:
    return _default_precision if p is None else _PRECISIONS[p]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(t: Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SwinFuseError(f"{what}: expected a CUDA tensor (libswinfuse has no CPU path), got "
                            f"{type(t).__name__} on {getattr(t, 'device', '?')}")
    if t.dtype != torch.float32:
        raise SwinFuseError(f"{what}: expected float32, got {t.dtype}")


def _param(t: Optional[Tensor], what: str) -> Optional[Tensor]:
    """weights: fp32, CUDA, contiguous (exactly how nn.Linear / nn.Conv2d store them)."""
    if t is None:
        return None
    t = t.detach()
    _require_cuda(t, what)
    return t if t.is_contiguous() else t.contiguous()


def is_fmap(x: Tensor) -> bool:
    return x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()


def new_fmap(b: int, c: int, h: int, w: int, like: Tensor) -> Tensor:
    return torch.empty((b, c, h, w), dtype=torch.float32, device=like.device, memory_format=torch.channels_last)


def as_fmap(x: Tensor, what: str = "input") -> Tensor:
    """Return x with NHWC memory (no copy if it already is)."""
    _require_cuda(x, what)
    if x.dim() != 4:
        raise SwinFuseError(f"{what}: expected a (B,C,H,W) tensor, got shape {tuple(x.shape)}")
    x = x.detach() if not x.requires_grad else x
    if is_fmap(x):
        return x
    return _NchwToNhwc.apply(x)


def to_nchw_contiguous(x: Tensor) -> Tensor:
    """NHWC-memory fmap -> plain contiguous NCHW (for callers that need it)."""
    return _NhwcToNchw.apply(x)


# ------------------------------------------------------------------------------------------------
# inference edges of a017_test.py
# ------------------------------------------------------------------------------------------------
def bgr_to_y_crcb(bgr: Tensor) -> Tuple[Tensor, Tensor]:
    """(B,H,W,3) uint8 BGR on the device -> (y (B,1,H,W), crcb (B,2,H,W)) float32: cv2 BGR2YCrCb on uint8 +
    ToImage/ToDtype(scale=True) + the split of a017:68, bit-exact (a015:86-93, a015:56-60)."""
    if not isinstance(bgr, torch.Tensor) or not bgr.is_cuda:
        raise SwinFuseError("bgr_to_y_crcb: expected a CUDA tensor (libswinfuse has no CPU path)")
    if bgr.dtype != torch.uint8 or bgr.dim() != 4 or bgr.shape[-1] != 3:
        raise SwinFuseError(f"bgr_to_y_crcb expects a (B,H,W,3) uint8 tensor, got {tuple(bgr.shape)} {bgr.dtype}")
    bgr = bgr.contiguous()
    b, h, w, _ = bgr.shape
    y = torch.empty((b, 1, h, w), dtype=torch.float32, device=bgr.device)
    crcb = torch.empty((b, 2, h, w), dtype=torch.float32, device=bgr.device)
    check(_lib.load().sf_bgr_to_ycrcb(bgr.data_ptr(), y.data_ptr(), crcb.data_ptr(), b, h, w, _stream()), "sf_bgr_to_ycrcb")
    return y, crcb


def y_crcb_to_rgb(fus_y: Tensor, crcb: Tensor) -> Tensor:
    """clamp(fus_y,0,1), re-attach CrCb, YCrCb -> RGB as cv2 does on float32 (a017:83-88); (B,3,H,W) float32."""
    _require_cuda(fus_y, "fus_y")
    _require_cuda(crcb, "crcb")
    b, c, h, w = fus_y.shape
    if c != 1 or tuple(crcb.shape) != (b, 2, h, w) or fus_y.dtype != torch.float32 or crcb.dtype != torch.float32:
        raise SwinFuseError(f"y_crcb_to_rgb expects fus_y (B,1,H,W) and crcb (B,2,H,W) float32, got {tuple(fus_y.shape)}, {tuple(crcb.shape)}")
    fus_y, crcb = fus_y.contiguous(), crcb.contiguous()
    rgb = torch.empty((b, 3, h, w), dtype=torch.float32, device=fus_y.device)
    check(_lib.load().sf_ycrcb_to_rgb(fus_y.data_ptr(), crcb.data_ptr(), rgb.data_ptr(), b, h, w, _stream()), "sf_ycrcb_to_rgb")
    return rgb


# ------------------------------------------------------------------------------------------------
# two-path concurrency
# ------------------------------------------------------------------------------------------------
# The IR (x) and visible (y) paths of a block are independent kernel sequences (a002:58-82, a003:46-50;
# cross attention reads both inputs but writes its own path).  In no-grad mode the y path is issued on a
# side stream between a fork (side waits for the caller's stream) and a join (the caller's stream waits
# for the side stream), so the narrow late stages -- whose persistent grids do not fill 148 SMs -- run
# two kernels at a time.  Every dual operation is fully ordered at both ends, so tensors that cross
# streams are never reused before their readers were ordered behind the join; the fork/join pattern is
# also what CUDA-graph capture records as parallel branches.
import os as _os
_dual_streams = _os.environ.get("SWINFUSE_DUAL_STREAMS", "1") != "0"
_side_streams = {}


def set_dual_streams(on: bool) -> None:
    global _dual_streams
    _dual_streams = bool(on)


def dual_path(fx, fy):
    """(fx(), fy()) -- fy on a side stream when gradients are off and both run on a CUDA device."""
    if not _dual_streams or torch.is_grad_enabled() or not torch.cuda.is_available():
        return fx(), fy()
    main = torch.cuda.current_stream()
    key = main.device.index
    side = _side_streams.get(key)
    if side is None:
        if torch.cuda.is_current_stream_capturing():
            return fx(), fy()   # a stream cannot be created inside a capture; warm up eagerly first
        side = _side_streams[key] = torch.cuda.Stream(device=main.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        oy = fy()
    ox = fx()
    main.wait_stream(side)
    return ox, oy


_weights_epoch = 0


def invalidate_packed_cache() -> None:
    """Call after weights were modified behind autograd's back (e.g. by sf_adam_step writing through
    raw pointers): in-place torch ops bump ``_version`` and are detected automatically."""
    global _weights_epoch
    _weights_epoch += 1


def _packed_weights(holder, key_tensors, nbytes_fn, pack_fn, p, like: Tensor, what: str, variant=()):
    """bf16 tensor-core operand images of the weights: packed once per weight version, owned by the
    caller side (stored on the parameter object), passed to the library through ``params.packed``.

    The image of a (holder, variant) pair lives in ONE buffer for the life of the holder: a stale image is
    re-packed in place, never re-allocated, so a CUDA graph that recorded the buffer's address stays valid
    whatever runs eagerly between replays.  ``variant`` carries everything besides the weight values that
    decides the image's layout (self / cross plan, geometry class, byte size): a module called both ways
    keeps one image per layout.  During stream capture a stale image is re-packed by kernels recorded INTO
    the graph, so every replay packs the weights of that moment (training: sf_adam_step rewrites them
    between replays)."""
    if p.precision != SF_PREC_BF16 or holder is None:
        return None
    nbytes = nbytes_fn(C.byref(p))
    if nbytes == 0:
        return None
    variant = (tuple(variant), int(nbytes), like.device.index)
    key = (_weights_epoch, tuple((t.data_ptr(), t._version) for t in key_tensors if t is not None))
    store = getattr(holder, "_sf_packed", None)
    if not isinstance(store, dict):
        store = {}
        try:
            holder._sf_packed = store
        except AttributeError:
            pass
    cached = store.get(variant)
    if cached is not None and cached[0] == key:
        return cached[1]
    if cached is not None:
        buf = cached[1]
    else:
        if torch.cuda.is_current_stream_capturing():
            raise SwinFuseError(f"{what}: the packed-weight buffer must exist before stream capture "
                                f"(run one eager warm-up call of the module first)")
        buf = torch.empty(nbytes, dtype=torch.uint8, device=like.device)
    check(pack_fn(C.byref(p), buf.data_ptr(), nbytes, _stream()), what)
    store[variant] = (key, buf)
    return buf


def _workspace(nbytes: int, like: Tensor) -> Tuple[Optional[Tensor], Optional[int]]:
    if nbytes == 0:
        return None, None
    ws = torch.empty(nbytes, dtype=torch.uint8, device=like.device)
    return ws, ws.data_ptr()


# ----------------------------------------------------------------------------------------------
# layout conversion
# ----------------------------------------------------------------------------------------------
def _nchw_to_nhwc_raw(x: Tensor) -> Tensor:
    x = x.contiguous()
    b, c, h, w = x.shape
    out = new_fmap(b, c, h, w, x)
    check(_lib.load().sf_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), b, c, h, w, _stream()), "sf_nchw_to_nhwc")
    return out


def _nhwc_to_nchw_raw(x: Tensor) -> Tensor:
    b, c, h, w = x.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().sf_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), b, c, h, w, _stream()), "sf_nhwc_to_nchw")
    return out


class _NchwToNhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _nchw_to_nhwc_raw(x)

    @staticmethod
    def backward(ctx, g):
        return g  # logical tensor unchanged; only the memory format differs


class _NhwcToNchw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _nhwc_to_nchw_raw(x if is_fmap(x) else _nchw_to_nhwc_raw(x))

    @staticmethod
    def backward(ctx, g):
        return g


def _g(t: Tensor) -> Tensor:
    """incoming gradient -> NHWC-memory fmap"""
    return t if is_fmap(t) else _nchw_to_nhwc_raw(t)


# ----------------------------------------------------------------------------------------------
# index ops
# ----------------------------------------------------------------------------------------------
class _PadReflect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pd, pr):
        b, c, h, w = x.shape
        out = new_fmap(b, c, h + pd, w + pr, x)
        check(_lib.load().sf_pad_reflect(x.data_ptr(), out.data_ptr(), b, h, w, c, pd, pr, _stream()), "sf_pad_reflect")
        ctx.pad = (pd, pr)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _g(g)
        pd, pr = ctx.pad
        b, c, ho, wo = g.shape
        h, w = ho - pd, wo - pr
        gin = new_fmap(b, c, h, w, g)
        check(_lib.load().sf_pad_reflect_bwd(g.data_ptr(), gin.data_ptr(), b, h, w, c, pd, pr, _stream()),
              "sf_pad_reflect_bwd")
        return gin, None, None


def pad_reflect(x: Tensor, pad_down: int, pad_right: int) -> Tensor:
    """MyPadding encoder branch (a006:111-131)."""
    x = as_fmap(x, "pad_reflect")
    if pad_down == 0 and pad_right == 0:
        return x
    return _PadReflect.apply(x, int(pad_down), int(pad_right))


class _Crop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, add, cd, cr):
        b, c, h, w = x.shape
        out = new_fmap(b, c, h - cd, w - cr, x)
        check(_lib.load().sf_crop(x.data_ptr(), _ptr(add), out.data_ptr(), b, h, w, c, cd, cr, _stream()), "sf_crop")
        ctx.geom = (h, w, cd, cr, add is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _g(g)
        h, w, cd, cr, has_add = ctx.geom
        b, c = g.shape[:2]
        gin = new_fmap(b, c, h, w, g)
        check(_lib.load().sf_crop_bwd(g.data_ptr(), gin.data_ptr(), b, h, w, c, cd, cr, _stream()), "sf_crop_bwd")
        return gin, (g if has_add else None), None, None


def crop(x: Tensor, crop_down: int, crop_right: int, add: Optional[Tensor] = None) -> Tensor:
    """MyPadding decoder branch (a006:133-146), optionally fused with the U-Net skip add (a013:222-225)."""
    x = as_fmap(x, "crop")
    if add is not None:
        add = as_fmap(add, "crop.add")
        if tuple(add.shape) != (x.shape[0], x.shape[1], x.shape[2] - crop_down, x.shape[3] - crop_right):
            raise SwinFuseError(f"crop: skip tensor {tuple(add.shape)} does not match the cropped map")
    if crop_down == 0 and crop_right == 0 and add is None:
        return x
    return _Crop.apply(x, add, int(crop_down), int(crop_right))


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        check(_lib.load().sf_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "sf_add")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a: Tensor, b: Tensor) -> Tensor:
    a, b = as_fmap(a, "add"), as_fmap(b, "add")
    if a.shape != b.shape:
        raise SwinFuseError(f"add: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    return _Add.apply(a, b)


def patch_merge(x: Tensor, ms: Tuple[int, int]) -> Tensor:
    """a011:87-93 (index kernel only; no autograd -- the fused patch_layer owns the gradient)."""
    x = as_fmap(x.detach(), "patch_merge")
    b, c, h, w = x.shape
    out = new_fmap(b, c * ms[0] * ms[1], h // ms[0], w // ms[1], x)
    check(_lib.load().sf_patch_merge(x.data_ptr(), out.data_ptr(), b, h, w, c, ms[0], ms[1], _stream()), "sf_patch_merge")
    return out


def patch_unmerge(x: Tensor, ms: Tuple[int, int]) -> Tensor:
    """a011:111-117 (index kernel only)."""
    x = as_fmap(x.detach(), "patch_unmerge")
    b, cc, h, w = x.shape
    c = cc // (ms[0] * ms[1])
    out = new_fmap(b, c, h * ms[0], w * ms[1], x)
    check(_lib.load().sf_patch_unmerge(x.data_ptr(), out.data_ptr(), b, h, w, c, ms[0], ms[1], _stream()),
          "sf_patch_unmerge")
    return out


def window_partition(x: Tensor, ws: Tuple[int, int], shift: bool) -> Tensor:
    """a001:165-172 (+ roll a001:442-445): (B,C,Hp,Wp) -> (B*nW, t, C)."""
    x = as_fmap(x.detach(), "window_partition")
    b, c, h, w = x.shape
    out = torch.empty((b * (h // ws[0]) * (w // ws[1]), ws[0] * ws[1], c), dtype=torch.float32, device=x.device)
    check(_lib.load().sf_window_partition(x.data_ptr(), out.data_ptr(), b, h, w, c, ws[0], ws[1], int(shift), _stream()),
          "sf_window_partition")
    return out


def window_reverse(t: Tensor, ws: Tuple[int, int], shift: bool, b: int, h: int, w: int) -> Tensor:
    """a001:390-398 (+ un-roll a001:471-473): (B*nW, t, C) -> (B,C,Hp,Wp)."""
    _require_cuda(t, "window_reverse")
    t = t.detach().contiguous()
    c = t.shape[-1]
    out = new_fmap(b, c, h, w, t)
    check(_lib.load().sf_window_reverse(t.data_ptr(), out.data_ptr(), b, h, w, c, ws[0], ws[1], int(shift), _stream()),
          "sf_window_reverse")
    return out


def shift_mask(h: int, w: int, ws: Tuple[int, int], device) -> Tensor:
    """a001:217-272: (nW, t, t) bool, True = masked."""
    t = ws[0] * ws[1]
    out = torch.empty(((h // ws[0]) * (w // ws[1]), t, t), dtype=torch.uint8, device=device)
    check(_lib.load().sf_shift_mask(out.data_ptr(), h, w, ws[0], ws[1], _stream()), "sf_shift_mask")
    return out.bool()


def relative_position_bias(table: Tensor, ws: Tuple[int, int]) -> Tensor:
    """a001:127-144: (t, t) bias gathered from the (2wsh-1, 2wsw-1) table."""
    table = _param(table, "relative_position_bias")
    t = ws[0] * ws[1]
    out = torch.empty((t, t), dtype=torch.float32, device=table.device)
    check(_lib.load().sf_relative_position_bias(table.data_ptr(), out.data_ptr(), ws[0], ws[1], _stream()),
          "sf_relative_position_bias")
    return out


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5, act: bool = False) -> Tensor:
    """my_layer_norm (a004:54-72), forward only (standalone use; the fused ops own the gradients)."""
    x = as_fmap(x.detach(), "layernorm")
    b, c, h, w = x.shape
    out = new_fmap(b, c, h, w, x)
    check(_lib.load().sf_layernorm(x.data_ptr(), _param(gamma, "gamma").data_ptr(), _param(beta, "beta").data_ptr(),
                                  out.data_ptr(), b * h * w, c, eps, int(act), _stream()), "sf_layernorm")
    return out


# ----------------------------------------------------------------------------------------------
# fused window attention
# ----------------------------------------------------------------------------------------------
_WA_TENSORS = ("ln_q_gamma", "ln_q_beta", "ln_kv_gamma", "ln_kv_beta", "wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo",
               "bias_table")


_direct_param_grads = False


def set_direct_param_grads(on: bool) -> None:
    """Let the backward kernels accumulate parameter gradients straight into ``param.grad`` (they add into their
    destination with atomics anyway) instead of returning fresh tensors for autograd to add.  Only parameters whose
    ``.grad`` already exists as a contiguous fp32 CUDA tensor take this route -- swinfuse.train.FlatParameters keeps
    every ``.grad`` as a zero-initialised view of its flat gradient buffer, which removes ~1.4k tiny accumulation kernels
    per training step.  Tensor hooks on such parameters do not fire (off by default; plain autograd semantics then)."""
    global _direct_param_grads
    _direct_param_grads = bool(on)


def _param_grad_buffers(saved, tens):
    """-> (buffers the kernels accumulate into, what backward() returns for them).

    One zero-filled flat buffer carved into views shaped like `tens` (None stays None) -- one fill instead of a dozen
    tiny ones per operator; with set_direct_param_grads(True) parameters that already own a .grad use it directly
    (and backward() returns None for them)."""
    bufs, rets = [None] * len(tens), [None] * len(tens)
    offs, total = {}, 0
    for i, (s, t) in enumerate(zip(saved, tens)):
        if t is None:
            continue
        g = getattr(s, "grad", None) if _direct_param_grads else None
        if g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.shape == t.shape:
            bufs[i] = g
        else:
            offs[i] = total
            total += (t.numel() + 63) // 64 * 64      # every view starts on a 256-byte boundary
    if offs:
        like = next(t for t in tens if t is not None)
        flat = torch.zeros(total, dtype=torch.float32, device=like.device)
        for i, o in offs.items():
            bufs[i] = rets[i] = flat[o:o + tens[i].numel()].view(tens[i].shape)
    return bufs, rets


def _fill_wa(p: _lib.WindowAttnParams, q, kv, residual, out, tensors, cfg) -> None:
    nh, d, wsh, wsw, shift, eps, prec = cfg
    b, c, h, w = q.shape
    p.q_src, p.kv_src, p.residual, p.out = q.data_ptr(), kv.data_ptr(), _ptr(residual), out.data_ptr()
    for name, t in zip(_WA_TENSORS, tensors):
        setattr(p, name, _ptr(t))
    p.B, p.Hp, p.Wp, p.C, p.num_heads, p.head_dim, p.wsh, p.wsw = b, h, w, c, nh, d, wsh, wsw
    p.shift, p.ln_eps, p.precision = int(shift), eps, prec


class _WindowAttn(torch.autograd.Function):
    """inputs: q, kv (None = self attention), residual (or None), 13 parameter tensors, cfg"""

    @staticmethod
    def forward(ctx, q, kv, residual, *rest):
        tensors, cfg = rest[:13], rest[13]
        lib = _lib.load()
        kv_t = q if kv is None else kv
        b, c, h, w = q.shape
        out = new_fmap(b, c, h, w, q)
        tens = [_param(t, n) for n, t in zip(_WA_TENSORS, tensors)]
        p = _lib.WindowAttnParams()
        _fill_wa(p, q, kv_t, residual, out, tens, cfg)
        # the image layout depends on self / cross and on the frag plan (7x7 windows below the index-range limit)
        variant = (kv is None, b * h * w >= (2 ** 31 - 1) // 64, cfg[2], cfg[3], cfg[0], cfg[1])
        packed = _packed_weights(tensors[4], tens[4:12], lib.sf_window_attn_packed_bytes, lib.sf_window_attn_pack, p, q,
                                 "sf_window_attn_pack", variant)
        p.packed = _ptr(packed)
        nbytes = lib.sf_window_attn_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, q)
        check(lib.sf_window_attn_fwd(C.byref(p), wsp, nbytes, _stream()), "sf_window_attn_fwd")
        ctx.cfg = cfg
        ctx.self_attn = kv is None
        ctx.has_residual = residual is not None
        # x + Attn(LN(x)): the residual is the query source itself -> its gradient (= gout) is added inside the backward
        ctx.fold_residual = residual is not None and residual.data_ptr() == q.data_ptr() and residual.shape == q.shape
        ctx.save_for_backward(q, kv_t, *tensors)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        gout = _g(gout)
        q, kv_t, *tensors = ctx.saved_tensors
        tens = [_param(t, n) for n, t in zip(_WA_TENSORS, tensors)]
        p = _lib.WindowAttnBwdParams()
        dummy_out = gout  # fwd.out is not read by the backward
        _fill_wa(p.fwd, q, kv_t, None, dummy_out, tens, ctx.cfg)
        gq = torch.empty_like(q)
        gkv = None if ctx.self_attn else torch.empty_like(kv_t)
        bufs, grads = _param_grad_buffers(tensors, tens)
        p.gout, p.g_q_src, p.g_kv_src = gout.data_ptr(), gq.data_ptr(), _ptr(gkv)
        for name, g in zip(_WA_TENSORS, bufs):
            setattr(p, "g_" + name, _ptr(g))
        p.add_to_g_q_src = gout.data_ptr() if ctx.fold_residual else None
        nbytes = lib.sf_window_attn_bwd_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, q)
        check(lib.sf_window_attn_bwd(C.byref(p), wsp, nbytes, _stream()), "sf_window_attn_bwd")
        gres = gout if (ctx.has_residual and not ctx.fold_residual) else None
        return (gq, gkv, gres, *grads, None)


def window_attention(q: Tensor, kv: Optional[Tensor], *, wq, bq, wk, bk, wv, bv, wo, bo, bias_table, num_heads: int,
                     head_dim: int, window_size: Tuple[int, int], shift: bool, ln_q=None, ln_kv=None,
                     residual: Optional[Tensor] = None, eps: float = 1e-5, precision=None) -> Tensor:
    """WindowAttention.forward (a001:448-474) [+ pre-norm / residual of a004:29-38].

    ``kv=None`` (or ``kv is q``) means self attention.  ``ln_q`` / ``ln_kv`` are optional
    (gamma, beta) pairs applied to q / kv inside the kernel; ``residual`` is added to the output.
    """
    q = as_fmap(q, "window_attention.q")
    if kv is not None and kv is not q:
        kv = as_fmap(kv, "window_attention.kv")
        if kv.shape != q.shape:
            raise SwinFuseError(f"window_attention: q {tuple(q.shape)} and k/v {tuple(kv.shape)} differ")
    else:
        kv = None
    if residual is not None:
        residual = as_fmap(residual, "window_attention.residual")
    h, w = q.shape[-2:]
    if h % window_size[0] or w % window_size[1]:
        raise SwinFuseError(f"window_attention: map ({h},{w}) is not a multiple of the window {tuple(window_size)}")
    lq = ln_q if ln_q is not None else (None, None)
    lkv = ln_kv if ln_kv is not None else (None, None)
    cfg = (int(num_heads), int(head_dim), int(window_size[0]), int(window_size[1]), bool(shift), float(eps),
           _prec(precision))
    return _WindowAttn.apply(q, kv, residual, lq[0], lq[1], lkv[0], lkv[1], wq, bq, wk, bk, wv, bv, wo, bo, bias_table,
                             cfg)


# ----------------------------------------------------------------------------------------------
# fused MLP
# ----------------------------------------------------------------------------------------------
_MLP_TENSORS = ("ln_gamma", "ln_beta", "w1", "b1", "w2", "b2")


def _fill_mlp(p: _lib.MlpParams, x, residual, out, tens, eps, prec) -> None:
    b, c, h, w = x.shape
    p.in_, p.residual, p.out = x.data_ptr(), _ptr(residual), out.data_ptr()
    for name, t in zip(_MLP_TENSORS, tens):
        setattr(p, name, _ptr(t))
    p.M, p.C, p.hidden, p.ln_eps, p.precision = b * h * w, c, tens[2].shape[0], eps, prec


class _Mlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, ln_g, ln_b, w1, b1, w2, b2, eps, prec):
        lib = _lib.load()
        out = torch.empty_like(x)
        tens = [_param(t, n) for n, t in zip(_MLP_TENSORS, (ln_g, ln_b, w1, b1, w2, b2))]
        p = _lib.MlpParams()
        _fill_mlp(p, x, residual, out, tens, eps, prec)
        packed = _packed_weights(w1, tens[2:6], lib.sf_mlp_packed_bytes, lib.sf_mlp_pack, p, x, "sf_mlp_pack")
        p.packed = _ptr(packed)
        nbytes = lib.sf_mlp_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_mlp_fwd(C.byref(p), wsp, nbytes, _stream()), "sf_mlp_fwd")
        fold = residual is not None and residual.data_ptr() == x.data_ptr() and residual.shape == x.shape
        ctx.cfg = (eps, prec, residual is not None, fold)
        ctx.save_for_backward(x, ln_g, ln_b, w1, b1, w2, b2)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        gout = _g(gout)
        x, *tensors = ctx.saved_tensors
        eps, prec, has_res, fold = ctx.cfg
        tens = [_param(t, n) for n, t in zip(_MLP_TENSORS, tensors)]
        p = _lib.MlpBwdParams()
        _fill_mlp(p.fwd, x, None, gout, tens, eps, prec)
        gin = torch.empty_like(x)
        bufs, grads = _param_grad_buffers(tensors, tens)
        p.gout, p.g_in = gout.data_ptr(), gin.data_ptr()
        for name, g in zip(_MLP_TENSORS, bufs):
            setattr(p, "g_" + name, _ptr(g))
        p.add_to_g_in = gout.data_ptr() if fold else None
        nbytes = lib.sf_mlp_bwd_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_mlp_bwd(C.byref(p), wsp, nbytes, _stream()), "sf_mlp_bwd")
        return (gin, gout if (has_res and not fold) else None, *grads, None, None)


def mlp(x: Tensor, *, w1, b1, w2, b2, ln=None, residual: Optional[Tensor] = None, eps: float = 1e-5,
        precision=None) -> Tensor:
    """AutoPathMLP.sequence_{x,y} (a003:21-31) [+ pre-norm / residual of a004:29-38]."""
    x = as_fmap(x, "mlp.x")
    if residual is not None:
        residual = as_fmap(residual, "mlp.residual")
    ln = ln if ln is not None else (None, None)
    if w1.dim() == 4 and (w1.shape[2] != 1 or w1.shape[3] != 1):
        raise SwinFuseError("mlp: only 1x1 convolutions are supported (a003:21-22)")
    return _Mlp.apply(x, residual, ln[0], ln[1], w1, b1, w2, b2, float(eps), _prec(precision))


# ----------------------------------------------------------------------------------------------
# patch merging / anti patch merging layer
# ----------------------------------------------------------------------------------------------
_PATCH_TENSORS = ("w", "b", "ln_gamma", "ln_beta")


def _fill_patch(p: _lib.PatchParams, x, out, tens, encoder, ms, cout, eps, prec) -> None:
    b, c, h, w = x.shape
    p.in_, p.out = x.data_ptr(), out.data_ptr()
    for name, t in zip(_PATCH_TENSORS, tens):
        setattr(p, name, _ptr(t))
    p.B, p.H, p.W, p.Cin, p.Cout, p.mh, p.mw = b, h, w, c, cout, ms[0], ms[1]
    p.encoder, p.ln_eps, p.precision = int(encoder), eps, prec


class _Patch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, ln_g, ln_b, encoder, ms, cout, eps, prec):
        lib = _lib.load()
        b, c, h, wd = x.shape
        out = new_fmap(b, cout, h // ms[0], wd // ms[1], x) if encoder else new_fmap(b, cout, h * ms[0], wd * ms[1], x)
        tens = [_param(t, n) for n, t in zip(_PATCH_TENSORS, (w, bias, ln_g, ln_b))]
        p = _lib.PatchParams()
        _fill_patch(p, x, out, tens, encoder, ms, cout, eps, prec)
        packed = _packed_weights(w, tens[0:2], lib.sf_patch_packed_bytes, lib.sf_patch_pack, p, x, "sf_patch_pack")
        p.packed = _ptr(packed)
        nbytes = lib.sf_patch_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_patch_fwd(C.byref(p), wsp, nbytes, _stream()), "sf_patch_fwd")
        ctx.cfg = (encoder, ms, cout, eps, prec)
        ctx.save_for_backward(x, w, bias, ln_g, ln_b)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        gout = _g(gout)
        x, *tensors = ctx.saved_tensors
        encoder, ms, cout, eps, prec = ctx.cfg
        tens = [_param(t, n) for n, t in zip(_PATCH_TENSORS, tensors)]
        p = _lib.PatchBwdParams()
        _fill_patch(p.fwd, x, gout, tens, encoder, ms, cout, eps, prec)
        gin = torch.empty_like(x)
        grads = [torch.zeros_like(t) for t in tens]
        p.gout, p.g_in = gout.data_ptr(), gin.data_ptr()
        for name, g in zip(_PATCH_TENSORS, grads):
            setattr(p, "g_" + name, _ptr(g))
        nbytes = lib.sf_patch_bwd_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_patch_bwd(C.byref(p), wsp, nbytes, _stream()), "sf_patch_bwd")
        grads[0] = grads[0].view_as(tensors[0])
        return (gin, *grads, None, None, None, None, None)


def patch_layer(x: Tensor, *, w, b, ln_gamma, ln_beta, encoder: bool, merging_size: Tuple[int, int], out_dims: int,
                eps: float = 1e-5, precision=None) -> Tensor:
    """PatchMergingAndLinearLayer.forward for one path (a011:236-264)."""
    x = as_fmap(x, "patch_layer.x")
    ms = (int(merging_size[0]), int(merging_size[1]))
    h, wd = x.shape[-2:]
    if encoder and (h % ms[0] or wd % ms[1]):
        raise SwinFuseError(f"patch_layer: map ({h},{wd}) is not a multiple of the merging size {ms}")
    return _Patch.apply(x, w, b, ln_gamma, ln_beta, bool(encoder), ms, int(out_dims), float(eps), _prec(precision))


# ----------------------------------------------------------------------------------------------
# final head
# ----------------------------------------------------------------------------------------------
_HEAD_TENSORS = ("w1", "b1", "bn_gamma", "bn_beta", "w2", "b2")


def _fill_head(p: _lib.HeadParams, x, y, out, tens, rm, rv, sm, si, training, eps, momentum) -> None:
    b, _, h, w = x.shape
    p.x, p.y, p.out = x.data_ptr(), y.data_ptr(), out.data_ptr()
    for name, t in zip(_HEAD_TENSORS, tens):
        setattr(p, name, _ptr(t))
    p.running_mean, p.running_var, p.save_mean, p.save_invstd = rm.data_ptr(), rv.data_ptr(), _ptr(sm), _ptr(si)
    p.B, p.H, p.W, p.ksize, p.training = b, h, w, tens[0].shape[-1], int(training)
    p.bn_eps, p.bn_momentum = eps, momentum


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, w1, b1, bn_g, bn_b, w2, b2, running_mean, running_var, training, eps, momentum):
        lib = _lib.load()
        b, _, h, w = x.shape
        out = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device)
        tens = [_param(t, n) for n, t in zip(_HEAD_TENSORS, (w1, b1, bn_g, bn_b, w2, b2))]
        sm = torch.empty(2, dtype=torch.float32, device=x.device) if training else None
        si = torch.empty(2, dtype=torch.float32, device=x.device) if training else None
        p = _lib.HeadParams()
        _fill_head(p, x, y, out, tens, running_mean, running_var, sm, si, training, eps, momentum)
        nbytes = lib.sf_head_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_head_fwd(C.byref(p), wsp, nbytes, _stream()), "sf_head_fwd")
        ctx.cfg = (training, eps, momentum)
        ctx.save_for_backward(x, y, w1, b1, bn_g, bn_b, w2, b2, running_mean, running_var,
                              sm if sm is not None else running_mean, si if si is not None else running_var)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        gout = gout.contiguous()
        x, y, w1, b1, bn_g, bn_b, w2, b2, rm, rv, sm, si = ctx.saved_tensors
        training, eps, momentum = ctx.cfg
        tens = [_param(t, n) for n, t in zip(_HEAD_TENSORS, (w1, b1, bn_g, bn_b, w2, b2))]
        p = _lib.HeadBwdParams()
        _fill_head(p.fwd, x, y, gout, tens, rm, rv, sm, si, training, eps, momentum)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        grads = [torch.zeros_like(t) for t in tens]
        p.gout, p.g_x, p.g_y = gout.data_ptr(), gx.data_ptr(), gy.data_ptr()
        for name, g in zip(_HEAD_TENSORS, grads):
            setattr(p, "g_" + name, _ptr(g))
        nbytes = lib.sf_head_bwd_workspace_bytes(C.byref(p))
        ws, wsp = _workspace(nbytes, x)
        check(lib.sf_head_bwd(C.byref(p), wsp, nbytes, _stream()), "sf_head_bwd")
        return (gx, gy, *grads, None, None, None, None, None)


def final_head(x: Tensor, y: Tensor, *, w1, b1, bn_gamma, bn_beta, running_mean, running_var, w2, b2, training: bool,
               eps: float = 1e-5, momentum: float = 0.1) -> Tensor:
    """MyModel.do_final_layer (a013:126-152).  x, y: (B,1,H,W).  Updates running stats in place when training."""
    x, y = as_fmap(x, "final_head.x"), as_fmap(y, "final_head.y")
    if x.shape[1] != 1 or y.shape != x.shape:
        raise SwinFuseError(f"final_head: expected two (B,1,H,W) maps, got {tuple(x.shape)} and {tuple(y.shape)}")
    x, y = x.contiguous(), y.contiguous()  # C == 1: no data movement
    rm, rv = running_mean.detach(), running_var.detach()
    _require_cuda(rm, "running_mean")
    return _Head.apply(x, y, w1, b1, bn_gamma, bn_beta, w2, b2, rm, rv, bool(training), float(eps), float(momentum))


def launch_count() -> int:
    return int(_lib.load().sf_launch_count())


def reset_launch_count() -> None:
    _lib.load().sf_reset_launch_count()


def profile_enable(on: bool) -> None:
    """Bracket every libswinfuse kernel with CUDA events (see sf_profile_enable)."""
    check(_lib.load().sf_profile_enable(int(on)), "sf_profile_enable")


def profile_summary() -> dict:
    """{kernel name: dict(launches, total_ms, flops, bytes)} since the last call; synchronises."""
    buf = (_lib.ProfileEntry * 64)()
    n = _lib.load().sf_profile_summary(buf, 64)
    if n < 0:
        check(n, "sf_profile_summary")
    return {buf[i].name.decode(): dict(launches=int(buf[i].launches), total_ms=float(buf[i].total_ms),
                                       flops=float(buf[i].flops), bytes=float(buf[i].bytes)) for i in range(n)}
