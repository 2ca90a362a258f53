"""Reference-side drop-in: `valor_b200.dropin.install()` makes the UNMODIFIED reference
(`model/pretrain.py: VALOR`, `train_utils.py: set_parallel_optimizer_and_apex / conduct_train`) run on this
library.  The reference has no plugin API; its seams are (1) `import apex...`, (2) the encoder classes
`model/modeling.py` instantiates by name.  install() therefore

  * registers a replacement `apex` package (the vendored apex/ tree is no longer needed):
      apex.normalization.fused_layer_norm.FusedLayerNorm -> valor_b200's LayerNorm kernels (same ctor, same state dict)
      apex.amp.initialize / scale_loss / master_params   -> bf16 flavour of O2: fp32 masters + bf16 working copies in the
                                                             flat arenas (params.ParamStore), loss scale 1, never skips
      apex.parallel.DistributedDataParallel               -> torch.nn.parallel.DistributedDataParallel
  * swaps the classes the reference builds by name for the B200-native mirrors (same constructors, state-dict keys and
    forward signatures): model.videoswin.SwinTransformer3D, model.transformer.TransformerEncoder, model.bert.BertModel,
    model.modeling.AudioEmbeddings / BERTPredictionHead.

Everything else of the reference (task parsing, pooling, contrastive head, losses, data, logging) keeps running as
written, in torch, on the tensors these modules return.  apex amp O2 semantics reproduced: the model
computes in low precision from fp32 masters; `scale_loss` yields the fp32 loss (scale 1) and on exit leaves every
parameter's gradient in the fp32 arena, aliased as `p.grad`, which is what `clip_grad_norm_(amp.master_params(optimizer))`
and the reference's own AdamW consume; `optimizer.step()` is wrapped to refresh the bf16 working copies.
"""
import contextlib
import sys
import types

import torch
import torch.nn as nn

from . import functional as Fn
from .functional import LN

_STATE = {"installed": False, "dtype": torch.bfloat16}


class FusedLayerNorm(nn.Module):
    """apex.normalization.FusedLayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True)
    (apex/apex/normalization/fused_layer_norm.py:129-161) on valor_layernorm_fwd/bwd."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        assert len(normalized_shape) == 1 and elementwise_affine, "only the last-dim affine form is on VALOR's path"
        self.normalized_shape, self.eps, self.elementwise_affine = tuple(normalized_shape), eps, True
        self.weight = nn.Parameter(torch.ones(*normalized_shape))
        self.bias = nn.Parameter(torch.zeros(*normalized_shape))

    def forward(self, x):
        if not hasattr(self.weight, "main_grad"):
            raise RuntimeError("valor_b200 FusedLayerNorm: call apex.amp.initialize(model, optimizer, ...) first "
                               "(it lays the parameters out in the arenas the kernels read)")
        shp = x.shape
        dt = self.weight.lp.dtype
        y = Fn.layer_norm(x.reshape(-1, shp[-1]).to(dt), LN(self.weight, self.bias, self.eps))
        return y.view(shp)


class _AmpState:
    store = None
    model = None


def _amp_initialize(model, optimizer=None, enabled=True, opt_level="O2", **kw):
    """apex.amp.initialize (frontend.py:124-141, _initialize.py:176-198) -> arenas.  enabled=False keeps fp32."""
    from .params import ParamStore
    dtype = _STATE["dtype"] if enabled else torch.float32
    dev = next(model.parameters()).device
    store = ParamStore(model, dtype=dtype, device=dev)
    model.store = store
    model.compute_dtype = dtype
    model.rng = Fn.RngState(dev, seed=0)
    for m in model.modules():
        m._rng = model.rng
    for p in store.params:
        p.grad = p.main_grad            # autograd of the reference-side torch modules accumulates straight into the arena
    _AmpState.store, _AmpState.model = store, model
    if optimizer is not None:
        inner_step, inner_zero = optimizer.step, optimizer.zero_grad

        def step(*a, **k):
            r = inner_step(*a, **k)
            store.refresh_lp()          # apex O2's master -> model copy (_process_optimizer.py:354)
            return r

        def zero_grad(*a, **k):
            store.zero_grad()           # keep the p.grad <-> arena aliasing (torch would set p.grad = None)

        optimizer.step, optimizer.zero_grad = step, zero_grad
    # O2 patches model.forward to cast floating inputs to the model dtype and outputs back to fp32 (_initialize.py:184-198)
    inner_forward = model.forward

    def forward(batch, *a, **k):
        if isinstance(batch, dict):
            batch = {kk: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for kk, v in batch.items()}
        ctx = torch.autocast(dev.type, dtype=dtype) if dtype != torch.float32 else contextlib.nullcontext()
        with ctx:
            out = inner_forward(batch, *a, **k)
        if isinstance(out, dict):
            out = {kk: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for kk, v in out.items()}
        return out

    model.forward = forward
    return (model, optimizer) if optimizer is not None else model


@contextlib.contextmanager
def _amp_scale_loss(loss, optimizer, delay_unscale=False, **kw):
    """apex.amp.scale_loss (handle.py:107-152) with loss scale 1: nothing to unscale, never an overflow skip."""
    yield loss.float()


def _amp_master_params(optimizer):
    """apex.amp.master_params (_amp_state.py:60-69)"""
    for group in optimizer.param_groups:
        for p in group["params"]:
            yield p


def _module(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install(reference_root=None, dtype=torch.bfloat16):
    """Register the apex replacement and swap the encoder classes of the reference's `model` package.
    `reference_root` (e.g. a checkout of TXH-mercury/VALOR) is put on sys.path when given."""
    _STATE["dtype"] = dtype
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # ---- apex
    apex = _module("apex")
    norm = _module("apex.normalization")
    fln = _module("apex.normalization.fused_layer_norm")
    fln.FusedLayerNorm = norm.FusedLayerNorm = FusedLayerNorm
    norm.fused_layer_norm = fln
    amp = _module("apex.amp")
    amp.initialize, amp.scale_loss, amp.master_params = _amp_initialize, _amp_scale_loss, _amp_master_params
    par = _module("apex.parallel")
    par.DistributedDataParallel = torch.nn.parallel.DistributedDataParallel
    apex.normalization, apex.amp, apex.parallel = norm, amp, par
    # ---- encoder classes
    from . import bert as B, modeling as M, transformer as T, videoswin as V
    import importlib
    swaps = {"model.videoswin": {"SwinTransformer3D": V.SwinTransformer3D},
             "model.transformer": {"TransformerEncoder": T.TransformerEncoder},
             "model.bert": {"BertModel": B.BertModel},
             "model.modeling": {"TransformerEncoder": T.TransformerEncoder, "BertModel": B.BertModel,
                                "AudioEmbeddings": T.AudioEmbeddings, "BERTPredictionHead": M.BERTPredictionHead,
                                "FusedLayerNorm": FusedLayerNorm}}
    done = []
    for modname, names in swaps.items():
        mod = importlib.import_module(modname)
        for n, cls in names.items():
            if hasattr(mod, n):
                setattr(mod, n, cls)
                done.append(f"{modname}.{n}")
    _STATE["installed"] = True
    return done
