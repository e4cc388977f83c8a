"""Bit-exact parity of the index / permutation kernels with the oracle's integer index maps
(SURVEY section 8 rows a2, a3, a5, a6, a13, a14/a15 index parts).  Calls go through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from tests.util import dropin, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    return dropin().ops


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("b,c,h,w", [(2, 3, 9, 11), (1, 24, 128, 128), (1, 1, 256, 255), (3, 8, 7, 13)])
def test_reflect_pad_and_crop(ops, b, c, h, w):
    x = _rand(b, c, h, w)
    for win in ((7, 7), (2, 2)):
        ref, pad = fo.pad_reflect(x, win)
        got = ops.pad_reflect(x.cuda(), *pad)
        assert torch.equal(got.cpu(), ref)
        assert torch.equal(ops.crop(got, *pad).cpu(), x)
        idx = fo.reflect_pad_index(h, w, *pad)
        assert torch.equal(got.cpu(), x[:, :, idx[..., 0], idx[..., 1]])


def test_crop_with_skip_add_matches_float_add(ops):
    x, s = _rand(2, 8, 14, 21), _rand(2, 8, 12, 20, seed=1)
    got = ops.crop(x.cuda(), 2, 1, add=s.cuda())
    assert torch.equal(got.cpu(), x[:, :, :12, :20] + s)


def test_reflect_pad_requires_pad_smaller_than_map(ops):
    sw = dropin()
    with pytest.raises(sw.SwinFuseError):
        ops.pad_reflect(_rand(1, 2, 3, 3).cuda(), 4, 0)   # a006: reflect needs pad < size


@pytest.mark.parametrize("b,c,h,w", [(2, 6, 10, 12), (1, 1, 256, 256), (1, 24, 134, 134), (2, 5, 4, 6)])
def test_patch_merge_unmerge(ops, b, c, h, w):
    x = _rand(b, c, h, w)
    ref = fo.patch_merge(x, (2, 2))
    got = ops.patch_merge(x.cuda(), (2, 2))
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(ops.patch_unmerge(got, (2, 2)).cpu(), x)
    assert torch.equal(ops.patch_unmerge(ref.cuda(), (2, 2)).cpu(), fo.patch_unmerge(ref, (2, 2)))
    # channel map restated as integers
    idx = fo.patch_merge_index(c, (2, 2))
    for m in (0, len(idx) // 2, len(idx) - 1):
        ph, pw, ch = idx[m]
        assert torch.equal(got.cpu()[:, m], x[:, ch, ph::2, pw::2])


@pytest.mark.parametrize("b,c,h,w", [(2, 16, 14, 21), (1, 24, 133, 133), (1, 3, 7, 7), (2, 8, 35, 28)])
@pytest.mark.parametrize("shift", [False, True])
def test_window_partition_reverse(ops, b, c, h, w, shift):
    x = _rand(b, c, h, w)
    xs = torch.roll(x, shifts=(-3, -3), dims=(2, 3)) if shift else x
    ref = fo.window_partition(xs, (7, 7))
    got = ops.window_partition(x.cuda(), (7, 7), shift)
    assert torch.equal(got.cpu(), ref)
    back = ops.window_reverse(got, (7, 7), shift, b, h, w)
    assert torch.equal(back.cpu(), x)
    # explicit integer map
    src = fo.window_token_source_index(h, w, (7, 7), shift)
    nw = src.shape[0]
    assert torch.equal(got.cpu()[:nw], x[0][:, src[..., 0], src[..., 1]].permute(1, 2, 0))


def test_partition_against_reference_generated_fixture(ops):
    g = golden("blocks_patch_pad.npz")
    for h, w in [(14, 21), (7, 7), (35, 28)]:
        img = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w).cuda()
        for shift, key in ((False, "partition"), (True, "partition_shifted")):
            got = ops.window_partition(img, (7, 7), shift).squeeze(-1).cpu().to(torch.int32).numpy()
            np.testing.assert_array_equal(got, g[f"idx/{key}_{h}x{w}"])
        np.testing.assert_array_equal(ops.shift_mask(h, w, (7, 7), "cuda").cpu().numpy(), g[f"idx/mask_{h}x{w}"])


@pytest.mark.parametrize("h,w,ws", [(133, 133, (7, 7)), (14, 14, (7, 7)), (518, 259, (7, 7)), (16, 24, (8, 8))])
def test_shift_mask(ops, h, w, ws):
    np.testing.assert_array_equal(ops.shift_mask(h, w, ws, "cuda").cpu().numpy(), fo.shift_mask(h, w, ws))


@pytest.mark.parametrize("ws", [(7, 7), (8, 8), (4, 6)])
def test_relative_position_bias_gather(ops, ws):
    table = _rand(2 * ws[0] - 1, 2 * ws[1] - 1)
    got = ops.relative_position_bias(table.cuda(), ws)
    assert torch.equal(got.cpu(), fo.relative_position_bias(table, ws))


def test_layout_round_trip(ops):
    x = _rand(3, 24, 35, 21)
    f = ops.as_fmap(x.cuda())
    assert f.shape == x.shape and f.permute(0, 2, 3, 1).is_contiguous()
    assert torch.equal(f.cpu(), x)
    assert torch.equal(ops.to_nchw_contiguous(f).cpu(), x) and ops.to_nchw_contiguous(f).is_contiguous()


def test_full_size_round_trips(ops):
    """BASELINE config sizes: properties instead of a CPU comparison."""
    x = torch.rand(64, 24, 128, 128, device="cuda")
    p = ops.pad_reflect(x, 5, 5)
    assert p.shape == (64, 24, 133, 133) and torch.equal(ops.crop(p, 5, 5), x)
    assert torch.equal(p[:, :, 128:, :128], x[:, :, 122:127, :].flip(2))
    for shift in (False, True):
        w = ops.window_partition(p, (7, 7), shift)
        assert w.shape == (64 * 361, 49, 24)
        assert torch.equal(ops.window_reverse(w, (7, 7), shift, 64, 133, 133), p)
    m = ops.patch_merge(x, (2, 2))
    assert m.shape == (64, 96, 64, 64) and torch.equal(ops.patch_unmerge(m, (2, 2)), x)
    assert float(m.sum(dtype=torch.float64)) == pytest.approx(float(x.sum(dtype=torch.float64)), rel=1e-12)
