"""Data-parallel training plumbing for the drop-in MyModel.

One process per GPU.  All parameters (and their gradients, and the Adam moments) live in single
flat fp32 buffers: autograd accumulates straight into views of the flat gradient buffer, the only
collective of a step is ONE all-reduce of that buffer over NCCL (a016 itself has no DP hooks), and
the optimizer is one sf_adam_step kernel over the flat buffers (torch.optim.Adam semantics, a016:67).
"""
from __future__ import annotations

import math
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


class CosineWarmRestarts:
    """Host-side learning-rate schedule of a016:68-72 / a016:109-113: torch's CosineAnnealingWarmRestarts(T_0, T_mult=1,
    eta_min) stepped with a fractional epoch ``epoch - 1 + (iter - 1) / iters_per_epoch``:

        lr(t) = eta_min + (base_lr - eta_min) * (1 + cos(pi * (t mod T_0) / T_0)) / 2

    A plain closed form evaluated on the host: the only consumer is the ``lr`` scalar argument of sf_adam_step."""

    def __init__(self, base_lr: float, T_0: int, eta_min: float = 0.0, T_mult: int = 1):
        if T_0 <= 0 or not isinstance(T_0, int):
            raise ValueError(f"Expected positive integer T_0, but got {T_0}")
        if T_mult < 1 or not isinstance(T_mult, int):
            raise ValueError(f"Expected integer T_mult >= 1, but got {T_mult}")
        self.base_lr, self.T_0, self.eta_min, self.T_mult = float(base_lr), T_0, float(eta_min), T_mult
        self.last_epoch = 0.0
        self.last_lr = self.base_lr

    def lr_at(self, epoch: float) -> float:
        if epoch < 0:
            raise ValueError(f"Expected non-negative epoch, but got {epoch}")
        if epoch >= self.T_0:
            if self.T_mult == 1:
                t_cur, t_i = epoch % self.T_0, self.T_0
            else:
                n = int(math.log(epoch / self.T_0 * (self.T_mult - 1) + 1, self.T_mult))
                t_cur = epoch - self.T_0 * (self.T_mult ** n - 1) / (self.T_mult - 1)
                t_i = self.T_0 * self.T_mult ** n
        else:
            t_cur, t_i = epoch, self.T_0
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2

    def step(self, epoch: float) -> float:
        self.last_epoch = float(epoch)
        self.last_lr = self.lr_at(epoch)
        return self.last_lr

    def state_dict(self) -> dict:
        return dict(base_lr=self.base_lr, T_0=self.T_0, eta_min=self.eta_min, T_mult=self.T_mult,
                    last_epoch=self.last_epoch, last_lr=self.last_lr)

    def load_state_dict(self, sd: dict) -> None:
        for k in ("base_lr", "T_0", "eta_min", "T_mult", "last_epoch", "last_lr"):
            setattr(self, k, sd[k])


class FlatAdam:
    """Adam over FlatParameters with the library's sf_adam_step kernel (CUDA) -- one launch per step."""

    def __init__(self, flat: FlatParameters, lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.step_count = 0

    def state_dict(self) -> dict:
        """The layout ``torch.optim.Adam(model.parameters()).state_dict()`` has (a016:238-250 saves that): per-parameter
        ``step`` / ``exp_avg`` / ``exp_avg_sq`` in parameter order, one param group."""
        f = self.flat
        state = {}
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(f.params, f.offsets)):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(f.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts a torch.optim.Adam state dict over the same parameter list (a016:306-339)."""
        f = self.flat
        group = sd["param_groups"][0]
        if len(group["params"]) != len(f.params):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the model has {len(f.params)}")
        if group.get("weight_decay", 0) or group.get("amsgrad", False) or group.get("maximize", False):
            raise ValueError("FlatAdam implements plain Adam only (a016:67): no weight decay / amsgrad / maximize")
        self.lr, self.betas, self.eps = float(group["lr"]), tuple(group["betas"]), float(group["eps"])
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, (p, o) in enumerate(zip(f.params, f.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            steps.add(int(float(st["step"])))
            self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
        if len(steps) > 1:
            raise ValueError(f"FlatAdam keeps one step count for all parameters, the state has {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0

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
                 use_graph: bool = False, scheduler: Optional[CosineWarmRestarts] = None, sync_init: bool = True):
        self.model, self.loss_fn, self.group, self.n_buckets = model, loss_fn, group, n_buckets
        self.flat = FlatParameters(model.parameters())
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if sync_init and self.world > 1:
            # replicas must start identical (a016:42 initialises with kaiming-normal draws from the per-process RNG):
            # rank 0's parameters and buffers (BatchNorm running statistics of the head, a013:133) win, as in DDP
            dist.broadcast(self.flat.flat_param, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            for b in model.buffers():
                dist.broadcast(b, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        from . import ops
        if self.flat.flat_param.is_cuda:
            ops.set_direct_param_grads(True)   # every .grad is a zeroed view of the flat buffer: kernels accumulate into it
            ops.invalidate_packed_cache()      # the parameters moved into the flat buffer (and may have been broadcast over)
        self.opt = FlatAdam(self.flat, lr=lr)
        self.scheduler = scheduler
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
        # Every bf16 operand image is stale from here on, so the capture below RECORDS the pack kernels (into the
        # buffers the warm-up allocated): each replay then packs the weights sf_adam_step left behind.  Without this
        # the captured forward would find fresh images, record no pack kernel and replay on frozen weights.
        from . import ops
        ops.invalidate_packed_cache()
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

    def set_epoch(self, epoch: float) -> float:
        """a016:109-113 (``use_scheduler``): called after the optimizer step with the fractional epoch; the new learning
        rate is what the next sf_adam_step receives."""
        if self.scheduler is None:
            return self.opt.lr
        self.opt.lr = self.scheduler.step(epoch)
        return self.opt.lr

    def state_dict(self, current_epoch: int = 0) -> dict:
        """The checkpoint a016:238-250 writes: model / optimizer / scheduler state and the finished epoch."""
        return {"model_state": self.model.state_dict(), "optimizer_state": self.opt.state_dict(),
                "scheduler_state": None if self.scheduler is None else self.scheduler.state_dict(),
                "current_epoch": current_epoch}

    def load_state_dict(self, state: dict) -> int:
        """a016:306-339.  Parameters are written into the flat buffer in place (the views stay attached)."""
        self.model.load_state_dict(state["model_state"])
        self.opt.load_state_dict(state["optimizer_state"])
        if self.scheduler is not None and state.get("scheduler_state") is not None:
            self.scheduler.load_state_dict(state["scheduler_state"])
        from . import ops
        ops.invalidate_packed_cache()
        return int(state.get("current_epoch", 0)) + 1
