"""Host-side data-parallel logic on CPU with gloo, world_size 2 (no GPU needed): the flat
parameter / gradient buffers and the single gradient all-reduce of swinfuse.train."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from tests.util import dropin


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    shared = nn.Linear(6, 6)
    m = nn.Sequential(nn.Linear(5, 6), nn.Tanh(), shared, nn.Tanh(), shared, nn.Linear(6, 3))  # one module used twice
    return m


def _worker(rank, world, port, n_buckets, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dropin()
    from swinfuse.train import FlatParameters
    m = _make_model()
    flat = FlatParameters(m.parameters())
    g = torch.Generator().manual_seed(100)
    x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)
    flat.zero_grad()
    ((m(x[shard]) - y[shard]) ** 2).sum().backward()
    flat.all_reduce_grads(n_buckets=n_buckets)
    if rank == 0:
        # numpy, not a tensor: torch shares tensors between processes by file descriptor, and the receiver must fetch it while
        # this process is still alive -- a race with the exit below (seen as ConnectionResetError / FileNotFoundError)
        out.put(flat.flat_grad.detach().numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_buckets", [1, 3])
def test_flat_gradient_allreduce_world2(n_buckets):
    dropin()
    from swinfuse.train import FlatParameters, unique_parameters
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_buckets, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.from_numpy(out.get(timeout=120))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process reference: gradient of the full batch
    m = _make_model()
    flat = FlatParameters(m.parameters())
    g = torch.Generator().manual_seed(100)
    x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    flat.zero_grad()
    ((m(x) - y) ** 2).sum().backward()
    assert torch.allclose(got, flat.flat_grad, atol=1e-5)
    assert len(flat.params) == len(unique_parameters(m.parameters())) == 6   # the shared Linear counted once


def test_flat_parameters_are_views_and_aligned():
    dropin()
    from swinfuse.train import ALIGN, FlatParameters
    m = _make_model()
    before = [p.detach().clone() for p in m.parameters()]
    flat = FlatParameters(m.parameters())
    for p, b, o in zip(flat.params, before, flat.offsets):
        assert torch.equal(p.detach(), b)
        assert o % ALIGN == 0
        assert p.data_ptr() == flat.flat_param.data_ptr() + 4 * o
        assert p.grad.data_ptr() == flat.flat_grad.data_ptr() + 4 * o
    with torch.no_grad():
        flat.flat_param.add_(1.0)
    assert torch.allclose(next(m.parameters()).detach(), before[0] + 1.0)


def _init_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dropin()
    from swinfuse.train import DataParallelTrainer
    torch.manual_seed(1234 + 77 * rank)      # a016:42 draws kaiming-normal weights from the per-process RNG
    m = nn.Sequential(nn.Conv2d(2, 2, 3), nn.BatchNorm2d(2), nn.Conv2d(2, 1, 3))
    with torch.no_grad():
        m[1].running_mean.fill_(float(rank + 1))
        m[1].running_var.fill_(float(rank + 2))
    tr = DataParallelTrainer(m, loss_fn=None)
    out.put((rank, tr.flat.flat_param.detach().numpy().copy(), m[1].running_mean.numpy().copy(), m[1].running_var.numpy().copy(),
             [p.detach().numpy().copy() for p in m.parameters()]))   # by value (see _worker)
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_broadcasts_initial_parameters_and_buffers_world2():
    """replicas built from different RNG states must start from rank 0's parameters and BatchNorm
    buffers (only gradients are reduced afterwards)."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_init_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r = out.get(timeout=120)
        got[r[0]] = (torch.from_numpy(r[1]), torch.from_numpy(r[2]), torch.from_numpy(r[3]), [torch.from_numpy(a) for a in r[4]])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(got[0][0], got[1][0])
    assert torch.equal(got[1][1], torch.full((2,), 1.0)) and torch.equal(got[1][2], torch.full((2,), 2.0))
    torch.manual_seed(1234)
    ref = nn.Sequential(nn.Conv2d(2, 2, 3), nn.BatchNorm2d(2), nn.Conv2d(2, 1, 3))
    for a, b in zip(got[1][3], ref.parameters()):
        assert torch.equal(a, b.detach())


@pytest.mark.parametrize("T_0,eta_min,T_mult", [(10, 1e-5, 1), (3, 0.0, 1), (4, 1e-4, 2)])
def test_cosine_warm_restarts_matches_torch_scheduler(T_0, eta_min, T_mult):
    """a016:68-72,109-113: CosineAnnealingWarmRestarts stepped with a fractional epoch after every iteration."""
    dropin()
    from swinfuse.train import CosineWarmRestarts
    from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts
    p = nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-2)
    ref = CosineAnnealingWarmRestarts(opt, T_0=T_0, eta_min=eta_min, T_mult=T_mult)
    ours = CosineWarmRestarts(1e-2, T_0=T_0, eta_min=eta_min, T_mult=T_mult)
    iters = 7
    for epoch in range(1, 3 * T_0 + 2):
        for it in range(1, iters + 1):
            t = epoch - 1 + (it - 1) / iters
            ref.step(t)
            assert ours.step(t) == pytest.approx(ref.get_last_lr()[0], rel=1e-12, abs=1e-15)
    sd = ours.state_dict()
    again = CosineWarmRestarts(1.0, T_0=1)
    again.load_state_dict(sd)
    assert again.lr_at(2.5) == ours.lr_at(2.5) and again.last_lr == ours.last_lr


def test_flat_adam_state_dict_has_the_torch_adam_layout():
    """a016:238-250 / 306-339: the optimizer state must travel through torch.save into a torch.optim.Adam and back."""
    dropin()
    from swinfuse.train import FlatAdam, FlatParameters
    m = _make_model()
    flat = FlatParameters(m.parameters())
    opt = FlatAdam(flat, lr=3e-3)
    opt.step_count = 4
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    sd = opt.state_dict()
    ref = torch.optim.Adam(unique(m), lr=1.0)
    ref.load_state_dict(sd)                       # torch accepts it as its own
    assert ref.param_groups[0]["lr"] == 3e-3
    back = ref.state_dict()
    opt2 = FlatAdam(flat, lr=1.0)
    opt2.load_state_dict(back)
    assert opt2.step_count == 4 and opt2.lr == 3e-3
    for p, o in zip(flat.params, flat.offsets):   # padding between views is not part of the state
        assert torch.equal(opt2.exp_avg[o:o + p.numel()], opt.exp_avg[o:o + p.numel()])
        assert torch.equal(opt2.exp_avg_sq[o:o + p.numel()], opt.exp_avg_sq[o:o + p.numel()])


def unique(m):
    from swinfuse.train import unique_parameters
    return unique_parameters(m.parameters())
