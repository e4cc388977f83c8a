"""Oracle of the a017 inference edges (oracle/color_oracle.py) against the golden vectors produced by the real
third-party code paths (cv2.cvtColor, torchvision ToImage/ToDtype; oracle/make_golden_color.py), and -- when cv2 is
importable -- against cv2 itself on fresh data, including every 8-bit colour of a coarse lattice.  Bit-exact."""
import numpy as np
import pytest

from oracle import color_oracle as co
from tests.util import golden


def test_color_oracle_matches_golden_bit_for_bit():
    g = golden("color_edges.npz")
    y, crcb = co.bgr_to_y_crcb(g["bgr"])
    assert np.array_equal(y, g["y"]) and np.array_equal(crcb, g["crcb"])
    assert np.array_equal(co.y_crcb_to_rgb(g["fus_y"], g["crcb"]), g["rgb"])


def test_color_oracle_against_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    v = np.arange(0, 256, 5, dtype=np.uint8)
    lattice = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(1, -1, 3)          # 52^3 colours
    rng = np.random.default_rng(5)
    for bgr in (lattice, rng.integers(0, 256, (1, 1 << 16, 3), dtype=np.uint8)):
        assert np.array_equal(co.bgr_u8_to_ycrcb_u8(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb))
    ycc = rng.random((64, 64, 3), dtype=np.float32) * 1.2 - 0.1
    ycc[..., 0] = np.clip(ycc[..., 0], 0, 1)
    ref = cv2.cvtColor(ycc, cv2.COLOR_YCrCb2RGB)
    got = co.y_crcb_to_rgb(ycc[None, None, ..., 0], np.moveaxis(ycc[None, ..., 1:3], -1, 1))
    assert np.array_equal(np.moveaxis(got[0], 0, -1), ref)
