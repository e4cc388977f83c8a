"""Training host path on the GPU (VERDICT r1 next #4b/#4c): the flat Adam kernel against
torch.optim.Adam, CUDA-graph replay of DataParallelTrainer against eager steps (the packed bf16 weight images must follow
the optimizer), the a016 checkpoint round trip through the drop-in model, and -- with two devices -- one data-parallel step
over NCCL against the single-GPU step on the concatenated batch."""
import io
import os
import socket

import numpy as np
import pytest
import torch
from torch import nn

from oracle import fusion_oracle as fo
from oracle.make_golden import small_cfg
from tests.util import build_model, dropin, golden

pytestmark = pytest.mark.gpu

_ZERO_GRAD = ("k_for_heads.bias", "final_layer.0.bias")   # analytically zero gradients: Adam turns their round-off into noise


def _named_unique(m):
    seen, out = set(), {}
    for n, p in m.named_parameters():
        if id(p) not in seen:
            seen.add(id(p))
            out[n] = p
    return out


def test_adam_step_kernel_matches_torch_adam():
    """sf_adam_step == torch.optim.Adam (a016:67) over 6 steps, incl. the 1/world gradient scale and an lr change."""
    dropin()
    from swinfuse.train import FlatAdam, FlatParameters
    torch.manual_seed(0)
    shapes = [(24, 24), (24,), (13, 13), (96, 24, 1, 1), (1,), (384, 384)]
    ours = [nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [nn.Parameter(p.detach().clone()) for p in ours]
    flat = FlatParameters(ours)
    opt = FlatAdam(flat, lr=1e-2)
    topt = torch.optim.Adam(ref, lr=1e-2)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        if step == 3:
            opt.lr = 3e-3
            topt.param_groups[0]["lr"] = 3e-3
        flat.zero_grad()
        for p, r in zip(ours, ref):
            gr = torch.randn(p.shape, device="cuda", generator=g) * (10.0 ** (step - 3))
            p.grad.copy_(gr * 4.0)           # as if summed over 4 ranks
            r.grad = gr.clone()
        opt.step(grad_scale=0.25)
        topt.step()
        for p, r in zip(ours, ref):
            # an update is lr-sized whatever the gradient's scale: 1e-6 of it, plus fp32 round-off of the parameter itself
            tol = 1e-6 * topt.param_groups[0]["lr"] + 2.5e-7 * float(r.detach().abs().max())
            assert float((p.detach() - r.detach()).abs().max()) <= tol, (step, tuple(p.shape), float((p - r).abs().max()), tol)
    # state travels to torch.optim.Adam and back (a016:238-250, 306-339)
    t2 = torch.optim.Adam(ref, lr=1.0)
    t2.load_state_dict(opt.state_dict())
    for i, r in enumerate(ref):
        for k in ("exp_avg", "exp_avg_sq"):
            a, b = t2.state[r][k], topt.state[r][k]
            err = float(((a - b).abs() / b.abs().clamp_min(1e-3 * float(b.abs().max()))).max())
            assert err <= 1e-5, (k, tuple(r.shape), err)
        assert float(t2.state[r]["step"]) == 6.0


def _trainer(sw, cfg, precision, use_graph, lr=1e-2, scheduler=None):
    from swinfuse.loss_ops import FusionLoss
    from swinfuse.train import DataParallelTrainer
    sw.set_default_precision(precision)
    m = build_model(cfg, act=nn.ELU()).train()
    m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
    return m, DataParallelTrainer(m, FusionLoss(clamp01=True).cuda(), lr=lr, use_graph=use_graph, scheduler=scheduler)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graph_replayed_training_steps_equal_eager_steps(precision):
    """the CUDA graph of forward + loss + backward must re-pack the bf16 weight images on every replay
    (sf_adam_step rewrites the fp32 weights between replays) and must survive an eager forward between replays."""
    sw = dropin()
    cfg = small_cfg()
    ir, vis = (t.cuda() for t in fo.synth_inputs(2, 37, 45, seed=5))
    n_steps = 6
    try:
        runs = {}
        for use_graph in (False, True):
            m, tr = _trainer(sw, cfg, precision, use_graph, lr=2e-3)
            losses = []
            for s in range(n_steps):
                losses.append(float(tr.step(ir, vis)))
                if s == 3:   # a validation forward between two training steps (a016:199-233): eager, replaces nothing the graph reads
                    m.eval()
                    with torch.no_grad():
                        m(ir, vis)
                    m.train()
            torch.cuda.synchronize()
            assert (tr._graph is not None) == use_graph
            runs[use_graph] = (losses, {n: p.detach().cpu().clone() for n, p in _named_unique(m).items()})
            if use_graph:
                # the mechanism itself, free of training-dynamics noise: one more replay and one eager pass over the
                # SAME weights (no optimizer step in between) must see the same loss.  A graph that replays on the weight
                # images of capture time is off by the whole distance the optimizer has travelled since.
                tr._ir.copy_(ir)
                tr._vis.copy_(vis)
                tr._graph.replay()
                replayed = float(tr._loss)
                eager = float(tr._forward_backward(ir, vis))
                assert replayed == pytest.approx(eager, rel=1e-5), (replayed, eager, losses)
    finally:
        sw.set_default_precision("fp32")
        sw.ops.set_direct_param_grads(False)
    le, lg = runs[False][0], runs[True][0]
    assert le[0] != le[-1]
    assert abs(le[0] - le[-1]) > 1e-3 * abs(le[0]), "the loss must move for this comparison to mean anything"
    # two runs of the same recipe differ by the order of the fp32 atomics in the weight gradients; Adam's normalisation
    # and (bf16 mode) operand rounding amplify that from step to step -- two EAGER runs already differ by 1e-4 at step 1
    traj_tol = 2e-2 if precision == "bf16" else 2e-3
    for a, b in zip(le, lg):
        assert b == pytest.approx(a, rel=traj_tol), (le, lg)
    num = den = 0.0
    for n, pe in runs[False][1].items():
        if n.endswith(_ZERO_GRAD):
            continue
        d = runs[True][1][n] - pe
        num += float((d.double() ** 2).sum())
        den += float((pe.double() ** 2).sum())
    assert (num / den) ** 0.5 <= traj_tol, (num / den) ** 0.5


def test_bf16_mode_gradients_against_the_reference_autograd_fixtures():
    """VERDICT r1 next #4a: the benchmarked SF_PREC_BF16 training path (bf16 forward, tensor-core backward) against the
    REFERENCE's autograd gradients (tests/golden/model_small.npz `grad::*`, written by oracle/make_golden.py from the real
    modules) -- not against this repo's own fp32 kernels.  Tolerance: bf16 forward rounding (2e-2 on the image) seen through
    the backward pass: 2e-2 relative L2 over all parameters, 2e-1 of its own scale for the worst single tensor."""
    sw = dropin()
    g = golden("model_small.npz")
    cfg = small_cfg()
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    sw.set_default_precision("bf16")
    try:
        m = build_model(cfg, act=nn.ELU()).train()
        m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
        out = m(ir.cuda(), vis.cuda())
        ref_out = torch.from_numpy(g["out_train"])
        assert float((out.cpu() - ref_out).abs().max() / ref_out.abs().max()) <= 2e-2
        (out * torch.from_numpy(g["grad_weight"]).cuda()).sum().backward()
    finally:
        sw.set_default_precision("fp32")
    gscale = float(np.median([np.abs(g[k]).max() for k in g.files if k.startswith("grad::")]))
    worst, num, den = (0.0, ""), 0.0, 0.0
    for name, prm in _named_unique(m).items():
        ref = torch.from_numpy(g["grad::" + name])
        if name.endswith(_ZERO_GRAD):
            assert float(prm.grad.abs().max()) <= 2e-2 * gscale, name
            continue
        d = prm.grad.cpu() - ref
        worst = max(worst, (float(d.abs().max()) / max(float(ref.abs().max()), 0.05 * gscale), name))
        num += float((d.double() ** 2).sum())
        den += float((ref.double() ** 2).sum())
    rel_l2 = (num / den) ** 0.5
    print("bf16-mode gradients vs reference autograd: rel L2", rel_l2, "worst tensor", worst)
    assert rel_l2 <= 2e-2, rel_l2
    assert worst[0] <= 2e-1, worst


def test_default_config_bf16_gradients_against_oracle_autograd_64():
    """Same comparison at the DEFAULT widths (C = 24 ... 384, d = 3 ... 48) on a 64x64 pair: autograd over the pinned CPU
    oracle is the comparand (the oracle equals the reference bit for bit, tests/test_oracle_golden.py)."""
    sw = dropin()
    cfg = fo.FusionConfig()
    sd = fo.synth_state_dict(cfg)
    ir, vis = fo.synth_inputs(1, 64, 64)
    gw = torch.Generator().manual_seed(9)
    weight = torch.randn(1, 1, 64, 64, generator=gw)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k and "num_batches" not in k}
    osd = dict(sd)
    osd.update(leaves)
    ref_out = fo.model_forward(osd, ir, vis, cfg, training=True)
    names = list(leaves)
    ref_grads = dict(zip(names, torch.autograd.grad((ref_out * weight).sum(), [leaves[n] for n in names], allow_unused=True)))
    sw.set_default_precision("bf16")
    try:
        m = build_model(cfg, act=nn.ELU()).train()
        m.load_state_dict(sd, strict=True)
        out = m(ir.cuda(), vis.cuda())
        assert float((out.cpu() - ref_out.detach()).abs().max() / ref_out.detach().abs().max()) <= 2e-2
        (out * weight.cuda()).sum().backward()
    finally:
        sw.set_default_precision("fp32")
    # the state dict registers every block parameter under several alias keys; the oracle reads one
    # of them, so a parameter's reference gradient is the one non-empty gradient among its aliases
    aliases = {}
    for n, prm in m.named_parameters(remove_duplicate=False):
        aliases.setdefault(id(prm), (prm, []))[1].append(n)
    uniq = {}
    for prm, names in aliases.values():
        hits = [n for n in names if ref_grads.get(n) is not None]
        assert len(hits) == 1, names
        uniq[hits[0]] = prm
    gscale = float(np.median([float(ref_grads[n].abs().max()) for n in uniq]))
    worst, num, den, n_cmp = (0.0, ""), 0.0, 0.0, 0
    for name, prm in uniq.items():
        ref = ref_grads[name]
        if name.endswith(_ZERO_GRAD):
            continue
        d = prm.grad.cpu() - ref
        worst = max(worst, (float(d.abs().max()) / max(float(ref.abs().max()), 0.05 * gscale), name))
        num += float((d.double() ** 2).sum())
        den += float((ref.double() ** 2).sum())
        n_cmp += 1
    rel_l2 = (num / den) ** 0.5
    print("default-config bf16 gradients vs oracle autograd: rel L2", rel_l2, "worst tensor", worst, "tensors", n_cmp)
    assert n_cmp > 1000
    # 2.9e-2 as measured.  The figure is carried by a few LayerNorm / bias tensors behind saturated softmax rows: re-rounding
    # one LayerNorm mean by an ulp (sum * (1/C) instead of sum / C, tried in round 2) flips a handful of bf16 roundings and
    # moves it to 3.6e-2 -- the bound leaves room for that, the worst-tensor bound below is the sharper check.
    assert rel_l2 <= 4e-2, rel_l2
    assert worst[0] <= 3e-1, worst


def test_a016_checkpoint_round_trip_through_the_dropin_model(tmp_path):
    """SURVEY 8(f) row 4 (a016:238-250 save_my_state, a016:306-339 load_my_state, a017:50-54): model + torch.optim.Adam +
    CosineAnnealingWarmRestarts state saved with torch.save after two steps, loaded strictly into a fresh drop-in model /
    optimizer / scheduler: the restored run continues bit-identically (fp32 path, deterministic kernels aside from
    atomics -> compared at 1e-6) and the eval output of the restored model equals the saved model's."""
    from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts
    sw = dropin()
    from a008_loss import MyLoss
    cfg = small_cfg()
    ir, vis = (t.cuda() for t in fo.synth_inputs(2, 37, 45, seed=5))

    def make():
        m = build_model(cfg, act=nn.ELU()).train()
        m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        sch = CosineAnnealingWarmRestarts(optimizer=opt, T_0=20, eta_min=1e-5)
        return m, opt, sch, MyLoss().cuda()

    def step(m, opt, sch, lossf, it):
        fusion = torch.clamp_(m(ir, vis), min=0, max=1)      # a016:150-167
        loss, _ = lossf.calcu_total_loss(fusion_images=fusion, ir_images=ir, vis_images=vis)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sch.step(0 + it / 10)
        return float(loss.detach())

    m, opt, sch, lossf = make()
    for it in range(2):
        step(m, opt, sch, lossf, it)
    path = os.path.join(tmp_path, "state.pth")
    torch.save({"model_state": m.state_dict(), "optimizer_state": opt.state_dict(), "scheduler_state": sch.state_dict(),
                "current_epoch": 1}, path)
    m.eval()
    with torch.no_grad():
        want_eval = m(ir, vis).cpu()
    m.train()
    want_loss = step(m, opt, sch, lossf, 2)

    m2, opt2, sch2, lossf2 = make()
    state = torch.load(path, map_location="cuda")
    m2.load_state_dict(state["model_state"])                 # strict, 3,139-key contract incl. the aliased entries
    opt2.load_state_dict(state["optimizer_state"])
    sch2.load_state_dict(state["scheduler_state"])
    assert state["current_epoch"] == 1
    m2.eval()
    with torch.no_grad():
        got_eval = m2(ir, vis).cpu()
    assert torch.equal(got_eval, want_eval)
    m2.train()
    got_loss = step(m2, opt2, sch2, lossf2, 2)
    assert got_loss == pytest.approx(want_loss, rel=1e-6)
    assert sch2.get_last_lr() == sch.get_last_lr()

    # the same checkpoint drives the flat trainer (sf_adam_step + host-side cosine schedule)
    from swinfuse.loss_ops import FusionLoss
    from swinfuse.train import CosineWarmRestarts, DataParallelTrainer
    m3 = build_model(cfg, act=nn.ELU()).train()
    tr = DataParallelTrainer(m3, FusionLoss(clamp01=True).cuda(), lr=1e-3, scheduler=CosineWarmRestarts(1e-3, 20, 1e-5))
    try:
        assert tr.load_state_dict({**state, "scheduler_state": None}) == 2
        tr.set_epoch(0 + 1 / 10)      # the learning rate a016's scheduler left behind after iteration 1
        # a torch.optim.Adam over clones of the restored parameters, restored from the same optimizer state
        clones = [nn.Parameter(q.detach().clone()) for q in tr.flat.params]
        ropt = torch.optim.Adam(clones, lr=1e-3)
        ropt.load_state_dict(state["optimizer_state"])
        ropt.param_groups[0]["lr"] = tr.opt.lr
        flat_loss = float(tr.step(ir, vis))
        tr.set_epoch(0 + 2 / 10)
    finally:
        sw.ops.set_direct_param_grads(False)
    assert flat_loss == pytest.approx(want_loss, rel=1e-5)
    assert tr.opt.lr == pytest.approx(sch.get_last_lr()[0], rel=1e-12)
    # (1) restored state + sf_adam_step == torch.optim.Adam on the SAME gradients (those the kernels just produced)
    for c, q in zip(clones, tr.flat.params):
        c.grad = q.grad.detach().clone()
    ropt.step()
    lr_before = ropt.param_groups[0]["lr"]
    for c, q in zip(clones, tr.flat.params):
        assert float((c.detach() - q.detach()).abs().max()) <= 1e-5 * lr_before + 2.5e-7 * float(c.detach().abs().max())
    # (2) those gradients are the ones the torch-optimizer run saw in its third step (order-of-atomics noise aside)
    num = den = 0.0
    for (n, a), b in zip(_named_unique(m3).items(), _named_unique(m).values()):
        if n.endswith(_ZERO_GRAD):
            continue
        num += float(((a.grad - b.grad).double() ** 2).sum())
        den += float((b.grad.double() ** 2).sum())
    assert den > 0 and (num / den) ** 0.5 <= 1e-3, (num / den) ** 0.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sw = dropin()
        cfg = small_cfg()
        torch.manual_seed(100 + rank)                       # different replicas: the trainer must broadcast rank 0's
        m = build_model(cfg, act=nn.ELU()).train()
        if rank == 0:
            m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
        m.final_layer[1].eval()                             # running-stat BatchNorm: no batch coupling ("BN stats aside")
        from swinfuse.loss_ops import FusionLoss
        from swinfuse.train import DataParallelTrainer
        tr = DataParallelTrainer(m, FusionLoss(clamp01=True).cuda(), lr=1e-3)
        ir, vis = fo.synth_inputs(2 * world, 37, 45, seed=5)
        sl = slice(2 * rank, 2 * rank + 2)
        losses = [float(tr.step(ir[sl].cuda(), vis[sl].cuda())) for _ in range(2)]
        torch.cuda.synchronize()
        if rank == 0:
            # by value (numpy): shared-memory tensors are fetched from this process and race with its exit
            out.put((losses, {n: p.detach().cpu().numpy().copy() for n, p in _named_unique(m).items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_data_parallel_step_equals_one_gpu_on_the_concatenated_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    losses2, params2 = out.get(timeout=600)
    params2 = {n: torch.from_numpy(a) for n, a in params2.items()}
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    sw = dropin()
    cfg = small_cfg()
    m = build_model(cfg, act=nn.ELU()).train()
    m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
    m.final_layer[1].eval()
    from swinfuse.loss_ops import FusionLoss
    from swinfuse.train import DataParallelTrainer
    try:
        tr = DataParallelTrainer(m, FusionLoss(clamp01=True).cuda(), lr=1e-3)
        ir, vis = fo.synth_inputs(4, 37, 45, seed=5)
        for _ in range(2):
            tr.step(ir.cuda(), vis.cuda())
    finally:
        sw.ops.set_direct_param_grads(False)
    num = den = 0.0
    for n, p in _named_unique(m).items():
        if n.endswith(_ZERO_GRAD):
            continue
        d = params2[n] - p.detach().cpu()
        num += float((d.double() ** 2).sum())
        den += float((p.detach().cpu().double() ** 2).sum())
    assert (num / den) ** 0.5 <= 1e-4, (num / den) ** 0.5
