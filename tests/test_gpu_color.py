"""Device kernels of the a017 inference edges (sf_bgr_to_ycrcb, sf_ycrcb_to_rgb) against the oracle and the golden
vectors: integer / byte work and the float path are both bit-exact."""
import numpy as np
import pytest
import torch

from oracle import color_oracle as co
from tests.util import dropin, golden

pytestmark = pytest.mark.gpu


def test_color_edges_match_golden_bit_for_bit():
    sw = dropin()
    g = golden("color_edges.npz")
    y, crcb = sw.ops.bgr_to_y_crcb(torch.from_numpy(g["bgr"]).cuda())
    assert np.array_equal(y.cpu().numpy(), g["y"]) and np.array_equal(crcb.cpu().numpy(), g["crcb"])
    rgb = sw.ops.y_crcb_to_rgb(torch.from_numpy(g["fus_y"]).cuda(), torch.from_numpy(g["crcb"]).cuda())
    assert np.array_equal(rgb.cpu().numpy(), g["rgb"])


@pytest.mark.parametrize("b,h,w", [(1, 1, 1), (3, 5, 7), (2, 512, 640), (64, 256, 256), (1, 1023, 1025)])
def test_color_edges_match_oracle_on_random_and_ragged_shapes(b, h, w):
    sw = dropin()
    rng = np.random.default_rng(b * 1000 + h + w)
    bgr = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    y, crcb = sw.ops.bgr_to_y_crcb(torch.from_numpy(bgr).cuda())
    oy, ocrcb = co.bgr_to_y_crcb(bgr)
    assert np.array_equal(y.cpu().numpy(), oy) and np.array_equal(crcb.cpu().numpy(), ocrcb)
    fus = (rng.random((b, 1, h, w), dtype=np.float32) * 1.5 - 0.25).astype(np.float32)
    rgb = sw.ops.y_crcb_to_rgb(torch.from_numpy(fus).cuda(), crcb)
    assert np.array_equal(rgb.cpu().numpy(), co.y_crcb_to_rgb(fus, ocrcb))


def test_every_8bit_colour_of_a_lattice_and_round_trip():
    """All 64^3 colours of a step-4 lattice through the integer kernel (saturation, rounding), and the float
    round trip Y/Cr/Cb -> RGB stays within the quantisation of the 8-bit transform."""
    sw = dropin()
    v = np.arange(0, 256, 4, dtype=np.uint8)
    bgr = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(1, 512, 512, 3)
    y, crcb = sw.ops.bgr_to_y_crcb(torch.from_numpy(bgr).cuda())
    oy, ocrcb = co.bgr_to_y_crcb(bgr)
    assert np.array_equal(y.cpu().numpy(), oy) and np.array_equal(crcb.cpu().numpy(), ocrcb)
    rgb = sw.ops.y_crcb_to_rgb(y, crcb).cpu().numpy()[0]
    back = np.stack([rgb[2], rgb[1], rgb[0]], -1) * 255.0
    assert np.abs(back - bgr[0].astype(np.float32)).max() <= 2.5


def test_color_edges_reject_cpu_and_bad_shapes():
    sw = dropin()
    with pytest.raises(sw.SwinFuseError):
        sw.ops.bgr_to_y_crcb(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))
    with pytest.raises(sw.SwinFuseError):
        sw.ops.bgr_to_y_crcb(torch.zeros(1, 4, 4, 3, dtype=torch.float32, device="cuda"))
    with pytest.raises(sw.SwinFuseError):
        sw.ops.y_crcb_to_rgb(torch.zeros(1, 1, 4, 4, device="cuda"), torch.zeros(1, 3, 4, 4, device="cuda"))
