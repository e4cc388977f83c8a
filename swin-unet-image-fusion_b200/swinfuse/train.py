"""Data-parallel training plumbing for the drop-in MyModel.

One process per GPU.  All parameters (and their gradients, and the Adam moments) live in single
flat fp32 buffers: autograd accumulates straight into views of the flat gradient buffer, the only
collective of a step is ONE all-reduce of that buffer over NCCL (a016 itself has no DP hooks), and
the optimizer is one sf_adam_step kernel over the flat buffers (torch.optim.Adam semantics, a016:67).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

ALIGN = 64  # floats: every parameter view starts on a 256-byte boundary (the kernels use float4 loads)


def unique_parameters(params: Iterable[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    seen, out = set(), []
    for p in params:
        if id(p) not in seen and p.requires_grad:
            seen.add(id(p))
            out.append(p)
    return out


class FlatParameters:
    """Re-homes parameters and gradients into two flat buffers (views keep their shapes)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = unique_parameters(params)
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.offsets, off = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("all parameters must share one device and dtype")
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat_param = torch.zeros(off, dtype=dt, device=dev)
        self.flat_grad = torch.zeros(off, dtype=dt, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):  # re-attach if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def all_reduce_grads(self, group=None, n_buckets: int = 1) -> None:
        """Sum the flat gradient over the data-parallel group (averaging is folded into the optimizer)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if n_buckets <= 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            return
        step = (self.numel + n_buckets - 1) // n_buckets
        works = [dist.all_reduce(self.flat_grad[i:i + step], op=dist.ReduceOp.SUM, group=group, async_op=True)
                 for i in range(0, self.numel, step)]
        for w in works:
            w.wait()


class FlatAdam:
    """Adam over FlatParameters with the library's sf_adam_step kernel (CUDA) -- one launch per step."""

    def __init__(self, flat: FlatParameters, lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.step_count = 0

    def step(self, grad_scale: float = 1.0) -> None:
        from . import _lib
        if not self.flat.flat_param.is_cuda:
            raise _lib.SwinFuseError("FlatAdam: parameters must be CUDA tensors (no CPU path)")
        self.step_count += 1
        f = self.flat
        _lib.check(_lib.load().sf_adam_step(f.flat_param.data_ptr(), f.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), f.numel, self.lr, self.betas[0], self.betas[1],
                                           self.eps, self.step_count, grad_scale, torch.cuda.current_stream().cuda_stream),
                   "sf_adam_step")
        from . import ops
        ops.invalidate_packed_cache()   # the kernel wrote the weights through raw pointers


class DataParallelTrainer:
    """forward -> clamp (a016:153) -> loss -> backward -> gradient all-reduce -> Adam, per rank.

    ``use_graph``: after two eager warm-up steps the device work of zero-grad + forward + loss + backward
    (about 10k kernel launches) is captured once into a CUDA graph and replayed on static input buffers;
    the gradient all-reduce and the Adam kernel (whose bias correction depends on the step count) stay eager.
    Shapes must then stay fixed; a new shape re-captures."""

    def __init__(self, model: torch.nn.Module, loss_fn, lr: float = 1e-2, group=None, n_buckets: int = 1,
                 use_graph: bool = False):
        self.model, self.loss_fn, self.group, self.n_buckets = model, loss_fn, group, n_buckets
        self.flat = FlatParameters(model.parameters())
        from . import ops
        ops.set_direct_param_grads(True)   # every .grad is a zeroed view of the flat buffer: kernels accumulate into it
        self.opt = FlatAdam(self.flat, lr=lr)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = use_graph
        self._graph, self._shape, self._eager_steps = None, None, 0

    def _forward_backward(self, ir: torch.Tensor, vis: torch.Tensor) -> torch.Tensor:
        self.flat.zero_grad()
        fusion = self.model(ir, vis)
        if not getattr(self.loss_fn, "cfg", {}).get("clamp01", False):   # FusionLoss(clamp01=True) clamps in its kernels
            fusion = torch.clamp(fusion, 0, 1)
        loss = self.loss_fn(fusion, ir, vis)
        loss.backward()
        return loss.detach()

    def _capture(self, ir: torch.Tensor, vis: torch.Tensor) -> None:
        self._ir, self._vis = ir.clone(), vis.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # autograd's stream bookkeeping wants a warm-up on the capturing side stream
            self._forward_backward(self._ir, self._vis)
        torch.cuda.current_stream().wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        # the autograd engine runs backward nodes (and their allocations) on its own thread: thread-local capture mode
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            self._loss = self._forward_backward(self._ir, self._vis)
        self._shape = (tuple(ir.shape), tuple(vis.shape))

    def step(self, ir: torch.Tensor, vis: torch.Tensor) -> torch.Tensor:
        if self.use_graph and ir.is_cuda:
            if self._eager_steps < 2:      # first-call checks, packed-weight buffers, cuDNN plans
                self._eager_steps += 1
                loss = self._forward_backward(ir, vis)
            else:
                if self._graph is None or self._shape != (tuple(ir.shape), tuple(vis.shape)):
                    self._capture(ir, vis)
                self._ir.copy_(ir)
                self._vis.copy_(vis)
                self._graph.replay()
                loss = self._loss
        else:
            loss = self._forward_backward(ir, vis)
        self.flat.all_reduce_grads(self.group, self.n_buckets)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss
