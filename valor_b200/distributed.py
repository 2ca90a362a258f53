"""Collectives on the hot path — mirror of `utils/distributed.py:38-93` (contrastive feature
all-gather with autograd slice-back) plus the gradient all-reduce that torch DDP performs for the
reference (train_utils.py:232).  One process per GPU, NCCL over NVLink/NVSwitch via
torch.distributed; gloo on CPU for the world_size-2 logic tests.

B200-first differences: pretraining batches are equal-sized on every rank (drop_last=True,
train_utils.py:591), so the reference's size exchange + host `.item()` sync + padding
(utils/distributed.py:41-48) is dropped — one all_gather_into_tensor per feature, no host sync;
gradients live in ONE flat fp32 arena (params.ParamStore), so DDP's bucket machinery becomes a
handful of in-place all-reduce(AVG) calls over arena ranges.  `GradReducer` starts each range's
all-reduce as soon as backward has left the model segment that owns it (DDP's overlap of
communication with the rest of backward, torch/nn/parallel/distributed.py), so at N > 1 only the
last small range is exposed; with no reducer attached the whole arena is reduced in one call.
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


# ---------------------------------------------------------------------------------------------
# gradient all-reduce overlapped with backward
# ---------------------------------------------------------------------------------------------
# Forward order of the model: swin stages -> AST layers -> everything after the encoders (fusion BERT, heads).  A
# `mark(x, name)` placed where a segment BEGINS is an identity whose backward runs after every autograd node created
# later has finished, i.e. when the gradients of that segment and of all later segments are final.
SPLIT_DEPTH = 18   # a swin stage this deep (stage 3 of VideoSwin-B: 60M parameters) is split into three buckets


def swin_bucket(stage, depth, block):
    """Name of the bucket that swin block `block` of stage `stage` (with `depth` blocks) belongs to."""
    return f"swin.{stage}" if depth < SPLIT_DEPTH else f"swin.{stage}" + "abc"[min(block // (depth // 3), 2)]


def default_segments(names):
    """(mark name, parameter-name prefixes) in forward order, derived from the store's parameter names."""
    import re
    depth = {}
    for n in names:
        m = re.match(r"video_encoder\.layers\.(\d+)\.blocks\.(\d+)\.", n)
        if m:
            depth[int(m.group(1))] = max(depth.get(int(m.group(1)), 0), int(m.group(2)) + 1)
    segs = {}
    for st in sorted(depth):
        for blk in range(depth[st]):
            segs.setdefault(swin_bucket(st, depth[st], blk), []).append(f"video_encoder.layers.{st}.blocks.{blk}.")
        segs[swin_bucket(st, depth[st], depth[st] - 1)].append(f"video_encoder.layers.{st}.downsample.")
    out = [(k, tuple(v)) for k, v in segs.items()]
    out.append(("ast", ("audio_encoder.layer.",)))
    out.append(("post", ("multimodal_encoder.", "cls.", "hidden_trans_", "contra_head_", "text_fine_weight",
                         "video_fine_weight", "audio_fine_weight")))
    return tuple(out)


_active = None   # the reducer backward marks report to (one model per process)


class _Mark(Function):
    @staticmethod
    def forward(ctx, x, name):
        ctx.name = name
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if _active is not None:
            _active.segment_started(ctx.name)
        return g, None


def mark(x, name):
    """Segment boundary in the forward pass (identity; a no-op unless a reducer is active and x needs a gradient)."""
    if _active is None or not torch.is_grad_enabled() or not x.requires_grad:
        return x
    return _Mark.apply(x, name)


class GradReducer:
    """Bucketed all-reduce(AVG) of a ParamStore's gradient arena, each bucket launched from inside backward."""

    def __init__(self, store, segments=None):
        self.store = store
        segments = segments or default_segments(store.names)
        self.order = [name for name, _ in segments]
        self.ranges = {}
        covered = []
        for name, prefixes in segments:
            rs = []
            for lo, hi, _ in store.group_ranges:          # decay / no-decay groups: one contiguous range in each
                offs = [(o, k) for n, (o, k) in store.offsets.items() if lo <= o < hi and n.startswith(prefixes)]
                if not offs:
                    continue
                s0, e0 = min(o for o, _ in offs), max(o + k for o, k in offs)
                inside = [n for n, (o, _) in store.offsets.items() if s0 <= o < e0]
                assert all(n.startswith(prefixes) for n in inside), f"segment {name} is not contiguous in the arena"
                rs.append((s0, e0))
            self.ranges[name] = rs
            covered += rs
        covered.sort()
        self.rest, pos = [], 0                            # whatever no segment owns (embeddings, final norms, ...)
        for s0, e0 in covered:
            if s0 > pos:
                self.rest.append((pos, s0))
            pos = max(pos, e0)
        if pos < store.numel:
            self.rest.append((pos, store.numel))
        self._done, self._works = set(), []
        self.launched_in_backward = 0                     # buckets started from inside backward in the last step

    def attach(self):
        global _active
        _active = self
        return self

    def detach(self):
        global _active
        if _active is self:
            _active = None

    def _launch(self, s0, e0):
        g = self.store.grad[s0:e0]
        if g.is_cuda:
            self._works.append((dist.all_reduce(g, op=dist.ReduceOp.AVG, async_op=True), None))
        else:   # gloo has no AVG
            self._works.append((dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True), g))

    def segment_started(self, name):
        """Backward has reached the point where segment `name` began: it and every later segment are final."""
        if name not in self.ranges:
            return
        for later in self.order[self.order.index(name):]:
            if later not in self._done:
                self._done.add(later)
                self.launched_in_backward += len(self.ranges[later])
                for s0, e0 in self.ranges[later]:
                    self._launch(s0, e0)

    def finish(self):
        """After backward: reduce what no mark covered, then make the current stream wait for every bucket."""
        self.last_launched_in_backward, self.launched_in_backward = self.launched_in_backward, 0
        for name in self.order:
            if name not in self._done:
                for s0, e0 in self.ranges[name]:
                    self._launch(s0, e0)
        for s0, e0 in self.rest:
            self._launch(s0, e0)
        ws = world_size()
        for w, g in self._works:
            w.wait()
            if g is not None:
                g.div_(ws)
        self._done, self._works = set(), []


def allreduce_grads(store):
    """DDP's gradient averaging (train_utils.py:232) over the flat fp32 gradient arena: the buckets still outstanding
    when a GradReducer is attached to the store (store.reducer), otherwise the whole arena in one call."""
    if world_size() == 1:
        return
    red = getattr(store, "reducer", None)
    if red is not None:
        return red.finish()
    if store.grad.is_cuda:
        dist.all_reduce(store.grad, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(store.grad, op=dist.ReduceOp.SUM)
        store.grad.div_(world_size())


def overlap_grad_allreduce(store, enable=True):
    """Turn the overlapped bucketed all-reduce on (N > 1) or off for `store`."""
    old = getattr(store, "reducer", None)
    if old is not None:
        old.detach()
    store.reducer = GradReducer(store).attach() if (enable and world_size() > 1) else None
    return store.reducer
