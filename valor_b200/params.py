"""Flat parameter arenas (HBM layout of the trainable state).

Replaces apex amp O2's per-tensor bookkeeping (apex/apex/amp/_process_optimizer.py:14-63:
fp16 model copy + lazily created fp32 masters) and torch DDP's bucket copies with four
contiguous fp32 arenas (master, grad, exp_avg, exp_avg_sq) and one bf16 working arena, all
indexed by the same element offsets:

    [ weight-decay group ............ | no-decay group ........... ]

* `nn.Parameter.data` of every module becomes a VIEW into `master`, so state_dict keys,
  shapes and values stay exactly the reference's;
* `p.lp` is the bf16 (or, in fp32 parity mode, the master itself) working view GEMMs read;
* `p.main_grad` is the fp32 gradient view the weight-gradient kernels accumulate into;
* the two optimizer groups follow optim/misc.py:14-77 (substring rule on parameter names,
  including its quirks: `relative_position_bias_table` and every `*.bias` are no-decay,
  Swin `norm*.weight` / `layernorm*.weight` ARE decayed);
* parameters keep registration order inside a group, which places query/key/value weights
  (and their biases) back to back — functional.fused_lin relies on that.
"""
import math

import torch

from . import kernels as K

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")
ALIGN = 64  # elements; keeps every bf16 view 128-byte aligned for TMA


def is_no_decay(name):
    return any(nd in name for nd in NO_DECAY)


class ParamStore:
    def __init__(self, model, dtype=torch.bfloat16, device=None):
        named = list(model.named_parameters())  # de-duplicates shared tensors, first name wins
        device = device or named[0][1].device
        self.dtype = dtype
        self.names, self.params = [], []
        groups = {False: [], True: []}
        for n, p in named:
            if p.requires_grad:
                groups[is_no_decay(n)].append((n, p))
        offset = 0
        self.group_ranges = []
        layout = []
        for nd in (False, True):
            start = offset
            for n, p in groups[nd]:
                layout.append((n, p, offset))
                offset += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            self.group_ranges.append((start, offset, nd))
        self.numel = offset
        self.master = torch.zeros(offset, device=device, dtype=torch.float32)
        self.grad = torch.zeros(offset, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(offset, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(offset, device=device, dtype=torch.float32)
        self.lp = self.master if dtype == torch.float32 else torch.zeros(offset, device=device, dtype=dtype)
        self.offsets = {}
        for n, p, off in layout:
            k = p.numel()
            self.master[off:off + k].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = self.master[off:off + k].view(p.shape)
            p.main_grad = self.grad[off:off + k].view(p.shape)
            p.lp = self.lp[off:off + k].view(p.shape)
            self.offsets[n] = (off, k)
            self.names.append(n)
            self.params.append(p)
        # tiny device scratch (graph-capturable optimizer step)
        self.sumsq = torch.zeros(1, device=device, dtype=torch.float32)
        self.norm = torch.zeros(2, device=device, dtype=torch.float32)  # [grad_norm, clip_coef]
        self.hyper = [torch.zeros(8, device=device, dtype=torch.float32) for _ in self.group_ranges]
        self.step = 0
        if device.type == "cuda":
            self.refresh_lp()

    def refresh_lp(self):
        if self.lp is not self.master:
            K.cast_flat(self.master, self.lp)

    def zero_grad(self):
        self.grad.zero_()

    def set_hyper(self, lr_ratio, base_lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01):
        """Host-side per-step scalars (optim/adamw.py:76-83 bias correction, optim/sched.py:37-41);
        written with one small H2D copy so the launches stay CUDA-graph friendly."""
        self.step += 1
        lr = base_lr * lr_ratio
        step_size = lr * math.sqrt(1.0 - betas[1] ** self.step) / (1.0 - betas[0] ** self.step)
        for h, (_, _, nd) in zip(self.hyper, self.group_ranges):
            vals = torch.tensor([lr, betas[0], betas[1], eps, 0.0 if nd else weight_decay, step_size, 0.0, 0.0],
                                dtype=torch.float32)
            h.copy_(vals, non_blocking=True)

    def optimizer_step(self, max_norm=5.0):
        """clip_grad_norm_ (train_utils.py:359) + AdamW (optim/adamw.py:50-101) + bf16 refresh:
        4 launches over the arenas."""
        self.sumsq.zero_()
        K.grad_sumsq(self.grad, self.sumsq)
        K.clip_coef(self.sumsq, max_norm, self.norm)
        for h, (s, e, _) in zip(self.hyper, self.group_ranges):
            if e > s:
                K.adamw(self.master[s:e], self.grad[s:e], self.exp_avg[s:e], self.exp_avg_sq[s:e],
                        None if self.lp is self.master else self.lp[s:e], h, self.norm[1:2])

    def grad_of(self, name):
        off, k = self.offsets[name]
        return self.grad[off:off + k]
