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
        out.put(flat.flat_grad.clone())
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
    got = out.get(timeout=120)
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
