"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the input / output edges of the reference's inference
script -- SURVEY section 8(f) row 3.  numpy restatement, bit-exact by construction:

* ``bgr_to_y_crcb``: a015_dataset.py:86-93 (``cv2.cvtColor(vis, COLOR_BGR2YCrCb)`` on uint8) followed by
  ``v2.ToImage(); v2.ToDtype(torch.float32, scale=True)`` (a015:56-60) and the channel split of
  a017_test.py:68.  The arithmetic lives in two third-party packages that are not part of
  /root/reference: OpenCV (cv2 4.13 in this image; 8-bit RGB->YCrCb is fixed point with a 14-bit shift:
  Y = (4899 R + 9617 G + 1868 B + 2^13) >> 14, Cr = ((R - Y) 11682 + 128*2^14 + 2^13) >> 14,
  Cb = ((B - Y) 9241 + ...) >> 14, saturated to 8 bits) and torchvision 0.26
  (``to_dtype_image``: ``image.to(float32).mul_(1.0 / 255)``).
* ``y_crcb_to_rgb``: a017_test.py:83-88 -- ``clamp_(fus_y, 0, 1)``, ``concat([fus_y, cr_cb])``,
  ``cv2.cvtColor(float32, COLOR_YCrCb2RGB)``: R = fma(Cr - .5, 1.403, Y), G = fma(Cr - .5, -0.714,
  fma(Cb - .5, -0.344, Y)), B = fma(Cb - .5, 1.773, Y) in float32 (the fused multiply-adds are what
  OpenCV's vector path computes; pinned bit for bit against cv2 in this container).

Pinned: tests/golden/color_edges.npz holds cv2 / torchvision outputs generated here by
oracle/make_golden_color.py; tests/test_color_cpu.py checks this file against them.
"""
import numpy as np

SHIFT = 14
R2Y, G2Y, B2Y, YCRI, YCBI = 4899, 9617, 1868, 11682, 9241
INV255 = np.float32(1.0 / 255)


def bgr_u8_to_ycrcb_u8(bgr: np.ndarray) -> np.ndarray:
    """(..., 3) uint8 BGR -> (..., 3) uint8 YCrCb, OpenCV's 8-bit fixed-point path."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    half, delta = 1 << (SHIFT - 1), 128 << SHIFT
    y = (r * R2Y + g * G2Y + b * B2Y + half) >> SHIFT
    cr = ((r - y) * YCRI + delta + half) >> SHIFT
    cb = ((b - y) * YCBI + delta + half) >> SHIFT
    return np.stack([np.clip(y, 0, 255), np.clip(cr, 0, 255), np.clip(cb, 0, 255)], axis=-1).astype(np.uint8)


def bgr_to_y_crcb(bgr: np.ndarray):
    """(B,H,W,3) uint8 BGR -> y (B,1,H,W) float32 in [0,1], crcb (B,2,H,W) float32 (a015:86-93,56-60; a017:68)."""
    ycc = bgr_u8_to_ycrcb_u8(bgr).astype(np.float32) * INV255
    chw = np.ascontiguousarray(np.moveaxis(ycc, -1, 1))
    return chw[:, 0:1], chw[:, 1:3]


def _fma32(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)   # exact product, one rounding


def y_crcb_to_rgb(fus_y: np.ndarray, crcb: np.ndarray) -> np.ndarray:
    """fus_y (B,1,H,W), crcb (B,2,H,W) float32 -> RGB (B,3,H,W) float32 (a017:83-88)."""
    y = np.clip(fus_y[:, 0].astype(np.float32), np.float32(0), np.float32(1))
    cr = crcb[:, 0].astype(np.float32) - np.float32(0.5)
    cb = crcb[:, 1].astype(np.float32) - np.float32(0.5)
    r = _fma32(cr, np.float32(1.403), y)
    g = _fma32(cr, np.float32(-0.714), _fma32(cb, np.float32(-0.344), y))
    b = _fma32(cb, np.float32(1.773), y)
    return np.stack([r, g, b], axis=1)
