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
        self.step = 0
        self.reducer = None   # distributed.GradReducer when the gradient all-reduce overlaps backward (N > 1)
        # optim/adamw.py:52-53,60-70: a parameter whose .grad is None is skipped (no moment update, no weight
        # decay) and every parameter counts its OWN steps for the bias correction.  Which parameters receive no
        # gradient is a static property of (configuration, task): `set_unused` takes their names (the model's
        # `unused_parameter_names(task)`, the analogue of DDP's find_unused_parameters) and the update runs over
        # the maximal contiguous runs of used parameters that share a decay group and a step count.
        self.param_steps = [0] * len(self.params)
        self._unused = frozenset()
        self._runs = None
        self.hyper_table = torch.zeros(64, 8, device=device, dtype=torch.float32)
        self._hyper_host = torch.zeros(64, 8, dtype=torch.float32)
        if device.type == "cuda":   # (the host table stays pageable: the driver stages the copy, so rewriting it
            self.refresh_lp()       #  next step cannot race a pending transfer)

    def refresh_lp(self):
        if self.lp is not self.master:
            K.cast_flat(self.master, self.lp)

    def zero_grad(self):
        self.grad.zero_()

    # ---- which parameters take part in the update -------------------------------------------
    def set_unused(self, names=()):
        names = frozenset(names)
        unknown = names - set(self.names)
        assert not unknown, f"unknown parameter names: {sorted(unknown)[:4]}"
        if names != self._unused:
            self._unused = names
            self._runs = None

    def _plan(self):
        """-> [(start, end, no_decay, step_count)] contiguous element runs of used parameters."""
        runs = []
        for i, n in enumerate(self.names):
            if n in self._unused:
                continue
            off, k = self.offsets[n]
            end = off + (k + ALIGN - 1) // ALIGN * ALIGN
            nd = is_no_decay(n)
            st = self.param_steps[i]
            if runs and runs[-1][1] == off and runs[-1][2] == nd and runs[-1][3] == st:
                runs[-1] = (runs[-1][0], end, nd, st)
            else:
                runs.append((off, end, nd, st))
        assert len(runs) <= self.hyper_table.shape[0], "too many optimizer runs"
        return runs

    def set_hyper(self, lr_ratio, base_lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01):
        """Host-side per-step scalars (optim/adamw.py:76-83 bias correction with the parameter's own step count,
        optim/sched.py:37-41); written with one small H2D copy so the launches stay CUDA-graph friendly.  The run
        layout only changes when the unused set changes (a captured graph must be re-captured then)."""
        self.step += 1
        for i, n in enumerate(self.names):
            if n not in self._unused:
                self.param_steps[i] += 1
        runs = self._plan()
        if self._runs is None or [(r[0], r[1]) for r in runs] != [(r[0], r[1]) for r in self._runs]:
            self._layout_version = getattr(self, "_layout_version", 0) + 1
        self._runs = runs
        lr = base_lr * lr_ratio
        for j, (_, _, nd, st) in enumerate(runs):
            step_size = lr * math.sqrt(1.0 - betas[1] ** st) / (1.0 - betas[0] ** st)
            self._hyper_host[j] = torch.tensor([lr, betas[0], betas[1], eps, 0.0 if nd else weight_decay, step_size, 0.0, 0.0])
        self.hyper_table.copy_(self._hyper_host, non_blocking=True)

    def optimizer_step(self, max_norm=5.0):
        """clip_grad_norm_ (train_utils.py:359) + AdamW (optim/adamw.py:50-101) + bf16 refresh:
        2 + (number of runs) launches over the arenas."""
        if self._runs is None:
            self._runs = self._plan()
        self.sumsq.zero_()
        K.grad_sumsq(self.grad, self.sumsq)
        K.clip_coef(self.sumsq, max_norm, self.norm)
        for j, (s, e, _, _) in enumerate(self._runs):
            K.adamw(self.master[s:e], self.grad[s:e], self.exp_avg[s:e], self.exp_avg_sq[s:e],
                    None if self.lp is self.master else self.lp[s:e], self.hyper_table[j], self.norm[1:2])

    def grad_of(self, name):
        off, k = self.offsets[name]
        return self.grad[off:off + k]
