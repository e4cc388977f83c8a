"""CPU oracle for the Swin-UNet fusion hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a *restatement* (plain torch-CPU / numpy, functional, driven by a
state_dict) of the forward algorithm of RainbowZL0/swin-unet-image-fusion.  It is the
checker the CUDA path is compared with.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``swin-unet-image-fusion_b200/``) never does.

Parity status: **pinned** for the model path (rows a1-a17 of SURVEY.md section 8): the
fixtures under ``tests/golden/`` were produced by importing the real reference modules
from ``/root/reference`` in the authoring container (``oracle/make_golden.py``) and this
restatement reproduces them (``tests/test_oracle_golden.py``).  The loss (row a19) lives in
third-party ``kornia`` which is absent: ``oracle/kornia_restatement.py`` is **parity
unpinned**.

Every function cites the reference file:line it follows (``a001:448`` means
``/root/reference/a001_WindowAttention.py`` line 448).

Two flavours are provided on purpose:

* ``*_index`` helpers: explicit integer index maps (numpy) -- the specification the
  bit-exact index kernels are tested against;
* tensor functions that follow the reference op-for-op (roll, window partition copies,
  linear, matmul, softmax, ...) so that timing this oracle on host cores is a fair stand-in
  for the reference's own CPU path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# configuration (A000_CONFIG.py:55-69)
# ----------------------------------------------------------------------------------------
@dataclass
class FusionConfig:
    window_size: Tuple[int, int] = (7, 7)
    merging_size: Tuple[int, int] = (2, 2)
    in_dims_list: List[int] = field(default_factory=lambda: [1, 24, 48, 96, 192])
    out_dims_list: List[int] = field(default_factory=lambda: [24, 48, 96, 192, 384])
    att_num_heads: int = 8
    att_dims_per_head_ratio: float = 1 / 8
    mlp_hidden_dims_ratio: int = 4
    final_conv_layer_kernel_size: int = 3

    @property
    def n_stages(self) -> int:
        return len(self.in_dims_list)

    def dims_per_head(self, stage: int) -> int:
        # a013:174 / a013:191  math.floor(out_dims * ratio)
        return math.floor(self.out_dims_list[stage] * self.att_dims_per_head_ratio)

    def mlp_hidden(self, stage: int, encoder: bool) -> int:
        # a013:177 (encoder: out_dims*ratio)  a013:196 (decoder: in_dims*ratio)
        base = self.out_dims_list[stage] if encoder else self.in_dims_list[stage]
        return base * self.mlp_hidden_dims_ratio


# ----------------------------------------------------------------------------------------
# integer index maps  (bit-exact specification)
# ----------------------------------------------------------------------------------------
def padding_size(length: int, window: int) -> int:
    """a006:54-56."""
    return (window - length % window) % window


def reflect_index(i: int, length: int) -> int:
    """Source index of padded position ``i`` for F.pad(mode="reflect") on the high side
    (a006:128-131): out[L+k] = in[L-2-k]."""
    return i if i < length else 2 * (length - 1) - i


def reflect_pad_index(h: int, w: int, pad_down: int, pad_right: int) -> np.ndarray:
    """(hp, wp, 2) int32: source (row, col) of each padded pixel.  a006:111-131."""
    rows = np.array([reflect_index(i, h) for i in range(h + pad_down)], dtype=np.int32)
    cols = np.array([reflect_index(j, w) for j in range(w + pad_right)], dtype=np.int32)
    out = np.empty((h + pad_down, w + pad_right, 2), dtype=np.int32)
    out[..., 0] = rows[:, None]
    out[..., 1] = cols[None, :]
    return out


def shift_source_index(hp: int, wp: int, ws: Tuple[int, int], shifted: bool) -> np.ndarray:
    """(hp, wp, 2) int32: pixel of the un-shifted frame that lands at (r, c) of the frame the
    windows are cut from.  torch.roll(x, (-s,-s)) => shifted[r,c] = x[(r+s)%H,(c+s)%W]
    (a001:432-445)."""
    sh, sw = (ws[0] // 2, ws[1] // 2) if shifted else (0, 0)
    rows = (np.arange(hp, dtype=np.int32) + sh) % hp
    cols = (np.arange(wp, dtype=np.int32) + sw) % wp
    out = np.empty((hp, wp, 2), dtype=np.int32)
    out[..., 0] = rows[:, None]
    out[..., 1] = cols[None, :]
    return out


def window_token_source_index(hp: int, wp: int, ws: Tuple[int, int], shifted: bool) -> np.ndarray:
    """(nW, t, 2) int32: for window w = wh*nWw+ww and token t = i*ws_w+j, the (row, col) in
    the ORIGINAL (un-shifted) feature map.  Combines a001:165-172 with a001:442-445."""
    wsh, wsw = ws
    nwh, nww = hp // wsh, wp // wsw
    src = shift_source_index(hp, wp, ws, shifted)
    out = np.empty((nwh * nww, wsh * wsw, 2), dtype=np.int32)
    for wh in range(nwh):
        for ww in range(nww):
            blk = src[wh * wsh:(wh + 1) * wsh, ww * wsw:(ww + 1) * wsw]
            out[wh * nww + ww] = blk.reshape(wsh * wsw, 2)
    return out


def relative_position_index(ws: Tuple[int, int]) -> np.ndarray:
    """(2, t, t) int32 -- a001:113-125.  [0,i,j] = r_j - r_i + ws_h-1 (key minus query)."""
    wsh, wsw = ws
    t = wsh * wsw
    r = np.arange(t, dtype=np.int32) // wsw
    c = np.arange(t, dtype=np.int32) % wsw
    out = np.empty((2, t, t), dtype=np.int32)
    out[0] = r[None, :] - r[:, None] + (wsh - 1)
    out[1] = c[None, :] - c[:, None] + (wsw - 1)
    return out


def shift_region_id(hp: int, wp: int, ws: Tuple[int, int]) -> np.ndarray:
    """(hp, wp) int32 region id of each pixel of the SHIFTED frame -- a001:222-247."""
    wsh, wsw = ws
    sh, sw = wsh // 2, wsw // 2

    def reg(n: int, w: int, s: int) -> np.ndarray:
        i = np.arange(n)
        return np.where(i < n - w, 0, np.where(i < n - s, 1, 2)).astype(np.int32)

    return 3 * reg(hp, wsh, sh)[:, None] + reg(wp, wsw, sw)[None, :]


def shift_mask(hp: int, wp: int, ws: Tuple[int, int]) -> np.ndarray:
    """(nW, t, t) bool; True = masked (score overwritten with -1e10) -- a001:249-272."""
    wsh, wsw = ws
    nwh, nww = hp // wsh, wp // wsw
    rid = shift_region_id(hp, wp, ws)
    rid = rid.reshape(nwh, wsh, nww, wsw).transpose(0, 2, 1, 3).reshape(nwh * nww, wsh * wsw)
    return rid[:, :, None] != rid[:, None, :]


def patch_merge_index(c_in: int, ms: Tuple[int, int]) -> np.ndarray:
    """(ms_h*ms_w*c_in, 3) int32: merged channel -> (ph, pw, c) -- a011:87-93:
    out[b,(ph*ms_w+pw)*C+c,H,W] = in[b,c,ms_h*H+ph,ms_w*W+pw]."""
    mh, mw = ms
    out = np.empty((mh * mw * c_in, 3), dtype=np.int32)
    for ph in range(mh):
        for pw in range(mw):
            for c in range(c_in):
                out[(ph * mw + pw) * c_in + c] = (ph, pw, c)
    return out


# ----------------------------------------------------------------------------------------
# tensor restatement (follows the reference op for op)
# ----------------------------------------------------------------------------------------
def window_partition(x: Tensor, ws: Tuple[int, int]) -> Tensor:
    """b c (nh wh) (nw ww) -> (b nh nw) (wh ww) c  -- a001:154-172 (materialised copy)."""
    b, c, h, w = x.shape
    wsh, wsw = ws
    x = x.reshape(b, c, h // wsh, wsh, w // wsw, wsw)
    return x.permute(0, 2, 4, 3, 5, 1).reshape(b * (h // wsh) * (w // wsw), wsh * wsw, c)


def window_reverse(t: Tensor, ws: Tuple[int, int], b: int, h: int, w: int) -> Tensor:
    """(b nh nw) (wh ww) c -> b c (nh wh) (nw ww)  -- a001:373-398."""
    wsh, wsw = ws
    c = t.shape[-1]
    t = t.reshape(b, h // wsh, w // wsw, wsh, wsw, c)
    return t.permute(0, 5, 1, 3, 2, 4).reshape(b, c, h, w)


def relative_position_bias(table: Tensor, ws: Tuple[int, int]) -> Tensor:
    """a001:127-144."""
    idx = torch.from_numpy(relative_position_index(ws)).long()
    t = ws[0] * ws[1]
    return table[idx[0].reshape(-1), idx[1].reshape(-1)].reshape(t, t)


def window_attention(q_src: Tensor, kv_src: Tensor, p: Dict[str, Tensor], prefix: str,
                     num_heads: int, dims_per_head: int, ws: Tuple[int, int], shifted: bool) -> Tensor:
    """WindowAttention.forward(q, k, v) with k is v  -- a001:448-474.

    q_src / kv_src: (B, C, Hp, Wp) logical NCHW, Hp % ws_h == 0, Wp % ws_w == 0.
    ``p[prefix + 'q_for_heads.weight']`` etc.
    """
    b, c, h, w = q_src.shape
    sh, sw = ws[0] // 2, ws[1] // 2
    q, k, v = q_src, kv_src, kv_src
    if shifted:  # a001:442-445 -- the reference rolls q, k and v separately
        q = torch.roll(q, shifts=(-sh, -sw), dims=(2, 3))
        k = torch.roll(k, shifts=(-sh, -sw), dims=(2, 3))
        v = torch.roll(v, shifts=(-sh, -sw), dims=(2, 3))
    # a001:210-214
    q, k, v = window_partition(q, ws), window_partition(k, ws), window_partition(v, ws)
    q = F.linear(q, p[prefix + "q_for_heads.weight"], p.get(prefix + "q_for_heads.bias"))
    k = F.linear(k, p[prefix + "k_for_heads.weight"], p.get(prefix + "k_for_heads.bias"))
    v = F.linear(v, p[prefix + "v_for_heads.weight"], p.get(prefix + "v_for_heads.bias"))
    bw, t, _ = q.shape

    def heads(z: Tensor) -> Tensor:  # a001:189-194
        return z.reshape(bw, t, num_heads, dims_per_head).permute(0, 2, 1, 3)

    q, k, v = heads(q), heads(k), heads(v)
    # a001:333-341: scale AFTER the product, then + bias
    scores = torch.matmul(q, k.permute(0, 1, 3, 2)) * (dims_per_head ** -0.5)
    scores = scores + relative_position_bias(p[prefix + "relative_position_bias_table"], ws)[None, None]
    if shifted:  # a001:274-315: overwrite with -1e10
        mask = torch.from_numpy(shift_mask(h, w, ws))  # (nW, t, t)
        nw_img = mask.shape[0]
        scores = scores.reshape(b, nw_img, num_heads, t, t).masked_fill(mask[None, :, None], -1e10)
        scores = scores.reshape(bw, num_heads, t, t)
    weights = F.softmax(scores, dim=-1)  # a001:349
    out = torch.matmul(weights, v)  # a001:353
    out = out.permute(0, 2, 1, 3).reshape(bw, t, num_heads * dims_per_head)  # a001:368-371
    out = F.linear(out, p[prefix + "linear_projection.weight"], p[prefix + "linear_projection.bias"])  # a001:413
    out = window_reverse(out, ws, b, h, w)  # a001:416
    if shifted:  # a001:471-473
        out = torch.roll(out, shifts=(sh, sw), dims=(2, 3))
    return out


def layer_norm_c(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """my_layer_norm -- a004:54-72: LN over C of a (B,C,H,W) tensor."""
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), weight, bias, eps)
    return y.permute(0, 3, 1, 2)


def mlp(x: Tensor, p: Dict[str, Tensor], prefix: str, path: str) -> Tensor:
    """AutoPathMLP.sequence_{x,y} -- a003:21-31: conv1x1 -> ELU -> conv1x1 (dropout p=0)."""
    h = F.conv2d(x, p[f"{prefix}mlp_{path}_1.weight"], p[f"{prefix}mlp_{path}_1.bias"])
    h = F.elu(h)
    return F.conv2d(h, p[f"{prefix}mlp_{path}_2.weight"], p[f"{prefix}mlp_{path}_2.bias"])


def basic_block(x: Tensor, y: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, dims_per_head: int,
                ws: Tuple[int, int], shifted: bool, cross: bool) -> Tuple[Tensor, Tensor]:
    """BasicBlock.forward (dual path) -- a005:127-145, a004:29-38, a002:58-82."""
    if cross and bool((x == y).all()):  # a005:111-118 (the reference prints and exit()s)
        raise ValueError("cross attention needs two different modalities (a005:111-118)")
    # stage_1: x + Attn(LN1(x)), y + Attn(LN2(y))
    nx = layer_norm_c(x, p[prefix + "stage_1.norm_layer_1.weight"], p[prefix + "stage_1.norm_layer_1.bias"])
    ny = layer_norm_c(y, p[prefix + "stage_1.norm_layer_2.weight"], p[prefix + "stage_1.norm_layer_2.bias"])
    wa = prefix + "auto_path_win_att."
    if cross:  # a002:68-73: both use the pre-update normalised tensors
        ax = window_attention(nx, ny, p, wa + "window_attention_x.", num_heads, dims_per_head, ws, shifted)
        ay = window_attention(ny, nx, p, wa + "window_attention_y.", num_heads, dims_per_head, ws, shifted)
    else:
        ax = window_attention(nx, nx, p, wa + "window_attention_x.", num_heads, dims_per_head, ws, shifted)
        ay = window_attention(ny, ny, p, wa + "window_attention_y.", num_heads, dims_per_head, ws, shifted)
    x, y = x + ax, y + ay
    # stage_2: + MLP(LN(.))
    nx = layer_norm_c(x, p[prefix + "stage_2.norm_layer_1.weight"], p[prefix + "stage_2.norm_layer_1.bias"])
    ny = layer_norm_c(y, p[prefix + "stage_2.norm_layer_2.weight"], p[prefix + "stage_2.norm_layer_2.bias"])
    mp = prefix + "auto_path_mlp."
    return x + mlp(nx, p, mp, "x"), y + mlp(ny, p, mp, "y")


def self_and_cross_block_pair(x: Tensor, y: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int,
                              dims_per_head: int, ws: Tuple[int, int]) -> Tuple[Tensor, Tensor]:
    """a012:70-78 -> a009:90-109: self/normal, self/shifted, cross/normal, cross/shifted."""
    for att, cross in (("self_att_block.", False), ("cross_att_block.", True)):
        for blk, shifted in (("normal_window_block.", False), ("shifted_window_block.", True)):
            x, y = basic_block(x, y, p, prefix + att + blk, num_heads, dims_per_head, ws, shifted, cross)
    return x, y


def pad_reflect(x: Tensor, window: Tuple[int, int]) -> Tuple[Tensor, Tuple[int, int]]:
    """MyPadding encoder branch -- a006:111-131,167-177."""
    h, w = x.shape[-2:]
    pd, pr = padding_size(h, window[0]), padding_size(w, window[1])
    if pd == 0 and pr == 0:
        return x, (0, 0)
    return F.pad(x, (0, pr, 0, pd), mode="reflect"), (pd, pr)


def crop(x: Tensor, pad: Tuple[int, int]) -> Tensor:
    """MyPadding decoder branch -- a006:133-146."""
    h, w = x.shape[-2:]
    return x[:, :, :h - pad[0], :w - pad[1]]


def patch_merge(x: Tensor, ms: Tuple[int, int]) -> Tensor:
    """b c (H ph) (W pw) -> b (ph pw c) H W  -- a011:87-93."""
    b, c, h, w = x.shape
    mh, mw = ms
    x = x.reshape(b, c, h // mh, mh, w // mw, mw)
    return x.permute(0, 3, 5, 1, 2, 4).reshape(b, mh * mw * c, h // mh, w // mw)


def patch_unmerge(x: Tensor, ms: Tuple[int, int]) -> Tensor:
    """b (ph pw c) H W -> b c (H ph) (W pw)  -- a011:111-117."""
    b, cc, h, w = x.shape
    mh, mw = ms
    c = cc // (mh * mw)
    x = x.reshape(b, mh, mw, c, h, w)
    return x.permute(0, 3, 4, 1, 5, 2).reshape(b, c, h * mh, w * mw)


def patch_layer(x: Tensor, p: Dict[str, Tensor], prefix: str, path: str, encoder: bool,
                ms: Tuple[int, int]) -> Tensor:
    """PatchMergingAndLinearLayer.forward for one path -- a011:236-264.
    encoder: merge -> conv1x1 -> LN -> ELU; decoder: conv1x1 -> LN -> unmerge -> ELU."""
    wt, bs = p[f"{prefix}mlp_layer_{path}.weight"], p[f"{prefix}mlp_layer_{path}.bias"]
    lw, lb = p[f"{prefix}layer_norm_{path}.weight"], p[f"{prefix}layer_norm_{path}.bias"]
    if encoder:
        x = patch_merge(x, ms)
        x = F.conv2d(x, wt, bs)
        x = layer_norm_c(x, lw, lb)
    else:
        x = F.conv2d(x, wt, bs)
        x = layer_norm_c(x, lw, lb)
        x = patch_unmerge(x, ms)
    return F.elu(x)


def final_head(x: Tensor, y: Tensor, p: Dict[str, Tensor], training: bool = False,
               prefix: str = "final_layer.") -> Tensor:
    """a013:126-152: cat -> conv3x3 reflect -> BatchNorm2d(2) -> ELU -> conv3x3 reflect."""
    f = torch.cat([x, y], dim=1)
    k = p[prefix + "0.weight"].shape[-1]
    pad = k // 2
    f = F.conv2d(F.pad(f, (pad,) * 4, mode="reflect"), p[prefix + "0.weight"], p[prefix + "0.bias"])
    if training:
        f = F.batch_norm(f, None, None, p[prefix + "1.weight"], p[prefix + "1.bias"], True, 0.1, 1e-5)
    else:
        f = F.batch_norm(f, p[prefix + "1.running_mean"], p[prefix + "1.running_var"],
                         p[prefix + "1.weight"], p[prefix + "1.bias"], False, 0.1, 1e-5)
    f = F.elu(f)
    return F.conv2d(F.pad(f, (pad,) * 4, mode="reflect"), p[prefix + "3.weight"], p[prefix + "3.bias"])


def model_forward(p: Dict[str, Tensor], ir: Tensor, vis: Tensor, cfg: Optional[FusionConfig] = None,
                  training: bool = False, taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """MyModel.forward -- a013:209-230.  ``taps`` (optional dict) receives intermediate
    tensors keyed 'enc{i}.x', 'dec{j}.x', ... for stage-level parity checks."""
    cfg = cfg or FusionConfig()
    ws, ms = cfg.window_size, cfg.merging_size
    x, y = ir, vis
    pads: List[Tuple[int, int]] = []  # LIFO shared by all MyPadding instances (a013:56-58)
    skips: List[Tuple[Tensor, Tensor]] = []
    n = cfg.n_stages
    for i in range(n):
        pre = f"encoder_list.{i}."
        x, pd = pad_reflect(x, ms)
        y, _ = pad_reflect(y, ms)
        pads.append(pd)
        x = patch_layer(x, p, pre + "1.", "x", True, ms)
        y = patch_layer(y, p, pre + "1.", "y", True, ms)
        x, pd = pad_reflect(x, ws)
        y, _ = pad_reflect(y, ws)
        pads.append(pd)
        x, y = self_and_cross_block_pair(x, y, p, pre + "3.", cfg.att_num_heads, cfg.dims_per_head(i), ws)
        if taps is not None:
            taps[f"enc{i}.x"], taps[f"enc{i}.y"] = x, y
        if i < n - 1:
            skips.append((x, y))  # a013:219-220
    for j in range(n):
        s = n - 1 - j  # decoder_list[j] was built from stage s (a013:164-205)
        pre = f"decoder_list.{j}."
        if j > 0:  # a013:222-225
            hx, hy = skips.pop()
            x, y = x + hx, y + hy
        x, y = self_and_cross_block_pair(x, y, p, pre + "0.", cfg.att_num_heads, cfg.dims_per_head(s), ws)
        pd = pads.pop()
        x, y = crop(x, pd), crop(y, pd)
        x = patch_layer(x, p, pre + "2.", "x", False, ms)
        y = patch_layer(y, p, pre + "2.", "y", False, ms)
        pd = pads.pop()
        x, y = crop(x, pd), crop(y, pd)
        if taps is not None:
            taps[f"dec{j}.x"], taps[f"dec{j}.y"] = x, y
    return final_head(x, y, p, training)


# ----------------------------------------------------------------------------------------
# state_dict contract and deterministic synthetic weights
# ----------------------------------------------------------------------------------------
def _wa_keys(c: int, inner: int, ws: Tuple[int, int]) -> List[Tuple[str, Tuple[int, ...]]]:
    out = [("relative_position_bias_table", (2 * ws[0] - 1, 2 * ws[1] - 1))]
    for n in ("q_for_heads", "k_for_heads", "v_for_heads"):
        out += [(n + ".weight", (inner, c)), (n + ".bias", (inner,))]
    out += [("linear_projection.weight", (c, inner)), ("linear_projection.bias", (c,))]
    return out


def state_dict_spec(cfg: Optional[FusionConfig] = None) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(key, shape, canonical_key)] in the order of ``MyModel.state_dict()``.  Aliased
    entries (the same tensor registered under several parents -- a005:51-82, a003:21-31)
    share ``canonical_key`` (= the first key under which the tensor appears)."""
    cfg = cfg or FusionConfig()
    ws = cfg.window_size
    mm = cfg.merging_size[0] * cfg.merging_size[1]
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def patch(prefix: str, cin: int, cout: int) -> None:
        # registration order a011:60-67: mlp_layer_x, layer_norm_x, mlp_layer_y, layer_norm_y, buffer
        spec.append((prefix + "buffer_to_show_device", (1,), prefix + "buffer_to_show_device"))
        for path in ("x", "y"):
            for k, s in ((f"mlp_layer_{path}.weight", (cout, cin, 1, 1)), (f"mlp_layer_{path}.bias", (cout,)),
                         (f"layer_norm_{path}.weight", (cout,)), (f"layer_norm_{path}.bias", (cout,))):
                spec.append((prefix + k, s, prefix + k))

    def block(prefix: str, c: int, inner: int, hid: int) -> None:
        # a005:51-82: auto_path_win_att, auto_path_mlp, stage_1, stage_2
        wa = prefix + "auto_path_win_att."
        for path in ("x", "y"):
            for k, s in _wa_keys(c, inner, ws):
                kk = f"{wa}window_attention_{path}.{k}"
                spec.append((kk, s, kk))
        mp = prefix + "auto_path_mlp."
        mlp_shapes = {"1": ((hid, c, 1, 1), (hid,)), "2": ((c, hid, 1, 1), (c,))}
        for path in ("x", "y"):
            # module order a003:21-31: mlp_1, mlp_2, (dropouts), sequence (0 = mlp_1, 3 = mlp_2)
            for n in ("1", "2"):
                spec.append((f"{mp}mlp_{path}_{n}.weight", mlp_shapes[n][0], f"{mp}mlp_{path}_{n}.weight"))
                spec.append((f"{mp}mlp_{path}_{n}.bias", mlp_shapes[n][1], f"{mp}mlp_{path}_{n}.bias"))
            for idx, n in (("0", "1"), ("3", "2")):
                spec.append((f"{mp}sequence_{path}.{idx}.weight", mlp_shapes[n][0], f"{mp}mlp_{path}_{n}.weight"))
                spec.append((f"{mp}sequence_{path}.{idx}.bias", mlp_shapes[n][1], f"{mp}mlp_{path}_{n}.bias"))
        for st, other, src in (("stage_1.", "auto_path_win_att.", wa), ("stage_2.", "auto_path_mlp.", mp)):
            sp = prefix + st
            # a004:12-18: other_module registered first, then norm_layer_1, norm_layer_2
            for key, shape, canon in [e for e in spec if e[0].startswith(src)]:
                spec.append((sp + "other_module." + key[len(src):], shape, canon))
            for nl in ("norm_layer_1", "norm_layer_2"):
                for k in ("weight", "bias"):
                    spec.append((f"{sp}{nl}.{k}", (c,), f"{sp}{nl}.{k}"))

    def four_blocks(prefix: str, c: int, inner: int, hid: int) -> None:
        for att in ("self_att_block.", "cross_att_block."):
            for blk in ("normal_window_block.", "shifted_window_block."):
                block(prefix + att + blk, c, inner, hid)

    n = cfg.n_stages
    for i in range(n):
        cin, cout = cfg.in_dims_list[i], cfg.out_dims_list[i]
        patch(f"encoder_list.{i}.1.", cin * mm, cout)
        four_blocks(f"encoder_list.{i}.3.", cout, cfg.att_num_heads * cfg.dims_per_head(i), cfg.mlp_hidden(i, True))
    for j in range(n):
        s = n - 1 - j
        cin, cout = cfg.out_dims_list[s], cfg.in_dims_list[s]
        four_blocks(f"decoder_list.{j}.0.", cin, cfg.att_num_heads * cfg.dims_per_head(s), cfg.mlp_hidden(s, False))
        patch(f"decoder_list.{j}.2.", cin, cout * mm)
    k = cfg.final_conv_layer_kernel_size
    for key, shape in (("0.weight", (2, 2, k, k)), ("0.bias", (2,)), ("1.weight", (2,)), ("1.bias", (2,)),
                       ("1.running_mean", (2,)), ("1.running_var", (2,)), ("1.num_batches_tracked", ()),
                       ("3.weight", (1, 2, k, k)), ("3.bias", (1,))):
        spec.append(("final_layer." + key, shape, "final_layer." + key))
    return spec


def _crc(s: str) -> int:
    import zlib
    return zlib.crc32(s.encode()) & 0x7FFFFFFF


def synth_tensor(canon_key: str, shape: Tuple[int, ...], seed: int = 0) -> Tensor:
    """Deterministic value for one state_dict entry, independent of construction order.
    Unlike a016:382-390 (zero biases) every bias/affine term is non-trivial so that the
    parity tests exercise them."""
    g = torch.Generator().manual_seed(_crc(canon_key) ^ (seed * 2654435761 & 0x7FFFFFFF))
    leaf = canon_key.rsplit(".", 1)[-1]
    if canon_key.endswith("num_batches_tracked"):
        return torch.tensor(7, dtype=torch.long)
    if canon_key.endswith("buffer_to_show_device"):
        return torch.zeros(1)
    if canon_key.endswith("relative_position_bias_table"):
        return torch.randn(shape, generator=g)
    if canon_key.endswith("running_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if canon_key.endswith("running_var"):
        return 0.5 + torch.rand(shape, generator=g)
    is_norm = ("norm" in canon_key.rsplit(".", 2)[-2]) or canon_key.startswith("final_layer.1.")
    if is_norm:
        base = 1.0 if leaf == "weight" else 0.0
        return base + 0.1 * torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
    return math.sqrt(2.0 / fan_in) * torch.randn(shape, generator=g)  # kaiming-normal scale (a016:384)


def synth_state_dict(cfg: Optional[FusionConfig] = None, seed: int = 0,
                     spec: Optional[List[Tuple[str, Tuple[int, ...], str]]] = None) -> Dict[str, Tensor]:
    """Full state_dict (all alias keys present, aliases share storage)."""
    spec = spec if spec is not None else state_dict_spec(cfg)
    canon: Dict[str, Tensor] = {}
    out: Dict[str, Tensor] = {}
    for key, shape, ck in spec:
        if ck not in canon:
            canon[ck] = synth_tensor(ck, shape, seed)
        out[key] = canon[ck]
    return out


def synth_inputs(b: int, h: int, w: int, seed: int = 1) -> Tuple[Tensor, Tensor]:
    """ir, vis in [0,1) fp32 (a015:59 ToDtype(scale=True)); ir != vis (a005:111-118)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 1, h, w, generator=g), torch.rand(b, 1, h, w, generator=g)
