"""Host-side restatement of the fused core's softmax (swin-unet-image-fusion_b200/csrc/wa_common.cuh: wf_softmax_p, wf_l_bad).

The kernels compute P = 2^s (scores already in the log2 domain) WITHOUT subtracting the row maximum, round P and v to bf16 for
the tensor-core product, take the row sum from a ones column of v in the fp32 accumulator, and repeat a task with the maximum
subtracted when a row sum leaves [2^-100, 2^100].  These tests pin the two claims that make that legal against the reference
softmax (a001:343, torch.softmax in fp64 here): inside the window the result equals the shifted softmax to bf16 accuracy, and every
row the fast pass cannot represent trips the range check (nothing wrong is ever stored silently)."""
import torch

L_MIN, L_MAX = 2.0 ** -100, 2.0 ** 100
LOG2E = 1.4426950408889634


def fast_pass(s2: torch.Tensor, v: torch.Tensor):
    """s2: [rows, keys] fp32 log2-domain scores; v: [keys, d].  Returns (O, bad-row mask) as the kernels compute them."""
    p = torch.exp2(s2.float())                                   # ex2.approx.ftz.f32
    p = torch.where(p.abs() < 2.0 ** -126, torch.zeros_like(p), p)   # flush to zero
    vb = torch.cat([v.float(), torch.ones(v.shape[0], 1)], dim=1).to(torch.bfloat16).float()   # ones column: row sums
    acc = p.to(torch.bfloat16).float() @ vb                      # bf16 operands, fp32 accumulation
    l = acc[:, -1]
    bad = ~((l > L_MIN) & (l < L_MAX))                           # NaN and inf compare false
    return acc[:, :-1] / l[:, None], bad


def reference(s2: torch.Tensor, v: torch.Tensor):
    return torch.softmax(s2.double() / LOG2E, dim=1) @ v.double()


def test_fast_pass_matches_shifted_softmax_inside_the_window():
    g = torch.Generator().manual_seed(0)
    for scale, offset in [(1.0, 0.0), (8.0, 0.0), (12.0, 40.0), (12.0, -60.0), (3.0, 80.0), (3.0, -85.0)]:
        s2 = torch.randn(64, 49, generator=g) * scale + offset
        v = torch.randn(49, 3, generator=g)
        o, bad = fast_pass(s2, v)
        assert not bad.any(), (scale, offset)
        ref = reference(s2, v)
        err = float((o.double() - ref).abs().max() / ref.abs().max())
        assert err <= 1.5e-2, (scale, offset, err)               # two bf16 roundings (P, v): 2^-8 each at worst


def test_rows_outside_the_window_are_flagged_not_stored():
    v = torch.randn(49, 3, generator=torch.Generator().manual_seed(1))
    big = torch.full((4, 49), -20.0)
    big[:, 7] = 130.0                                            # 2^130 overflows fp32
    assert fast_pass(big, v)[1].all()
    tiny = torch.full((4, 49), -140.0)                           # everything flushes to zero: l = 0
    assert fast_pass(tiny, v)[1].all()
    low = torch.full((4, 49), -110.0)                            # representable, but below 2^-100: dominant terms near the denormal range
    assert fast_pass(low, v)[1].all()
    nan = torch.zeros(4, 49)
    nan[:, 3] = float("nan")
    assert fast_pass(nan, v)[1].all()
    masked = torch.randn(4, 49)                                  # masked scores (-1e10 * log2 e, a001:310) are plain zeros of P
    masked[:, 10:30] = -1.4426950e10
    o, bad = fast_pass(masked, v)
    assert not bad.any() and torch.isfinite(o).all()
    keep = torch.ones(49, dtype=torch.bool)
    keep[10:30] = False
    ref = reference(masked[:, keep], v[keep])
    assert float((o.double() - ref).abs().max() / ref.abs().max()) <= 1.5e-2
