"""Collectives on the hot path — mirror of `utils/distributed.py:38-93` (contrastive feature
all-gather with autograd slice-back) plus the gradient all-reduce that torch DDP performs for the
reference (train_utils.py:232).  One process per GPU, NCCL over NVLink/NVSwitch via
torch.distributed; gloo on CPU for the world_size-2 logic tests.

B200-first differences: pretraining batches are equal-sized on every rank (drop_last=True,
train_utils.py:591), so the reference's size exchange + host `.item()` sync + padding
(utils/distributed.py:41-48) is dropped — one all_gather_into_tensor per feature, no host sync;
gradients live in ONE flat fp32 arena (params.ParamStore), so the DDP bucket machinery collapses
to a single in-place all-reduce(AVG) of that arena.
"""
import torch
import torch.distributed as dist
from torch.autograd import Function


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _all_gather(x):
    ws = world_size()
    x = x.contiguous()
    out = torch.empty((ws * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(x) for _ in range(ws)]
        dist.all_gather(parts, x)
        out = torch.cat(parts, dim=0)
    return out


class ddp_allgather_with_grads(Function):
    """utils/distributed.py:38-72: forward concatenates every rank's rows; backward returns only the
    local rows of the incoming gradient (every rank evaluates the full global loss redundantly and
    the gradient all-reduce averages — semantics kept exactly)."""

    @staticmethod
    def forward(ctx, x):
        ctx.n = x.shape[0]
        if world_size() == 1:
            return x.clone() if False else x.view_as(x)
        return _all_gather(x)

    @staticmethod
    def backward(ctx, grad_output):
        if world_size() == 1:
            return grad_output
        r = rank()
        return grad_output[r * ctx.n:(r + 1) * ctx.n]


def ddp_allgather(x):
    """utils/distributed.py:77-93 (no autograd)."""
    if world_size() == 1:
        return x
    return _all_gather(x)


def allreduce_grads(store):
    """DDP's gradient averaging (train_utils.py:232) over the flat fp32 gradient arena."""
    if world_size() == 1:
        return
    if store.grad.is_cuda:
        dist.all_reduce(store.grad, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(store.grad, op=dist.ReduceOp.SUM)
        store.grad.div_(world_size())
