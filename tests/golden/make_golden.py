"""Mint golden vectors by running the LIVE reference (/root/reference, read-only) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/golden_<cfg>.json.  The fixtures travel to the GPU box; the reference
does not.  Every number is produced by the reference's own `model/pretrain.py: VALOR.forward`
(+ autograd backward of sum(losses), train_utils.py:306-319) on seeded synthetic weights and
inputs from oracle/synth.py; Dropout/DropPath disabled; TokenMasker draw hoisted.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tools import synth  # noqa: E402

CONFIGS = {
    "tiny": dict(geom=synth.TINY, B=2, F=2, A=1, T=16),
    "c1": dict(geom=synth.BASE, B=2, F=4, A=1, T=32),
    # the benchmarked per-sample shape (BASELINE configs[1]: 8 frames -> 392-token windows, 2 audio clips ->
    # 650 media tokens, merged 3-pass cross-attention over two modalities) at the CPU-affordable batch 2
    "c2shape": dict(geom=synth.BASE, B=2, F=8, A=2, T=32),
    # a task without audio objectives: pins which parameters the reference leaves without a gradient
    "tiny_tv": dict(geom=synth.TINY, B=2, F=2, A=1, T=16, task="pt_contra%tv_caption%tv"),
}
# configs that also record a 3-step trajectory under the reference's own optimizer stack
TRAJECTORY = {"tiny": 3, "c1": 3, "tiny_tv": 2}
TASK = "pt_contra%tva%tv%ta_caption%tva%tv%ta"
GRAD_KEYS = [
    "video_encoder.patch_embed.proj.weight",
    "video_encoder.layers.0.blocks.1.attn.relative_position_bias_table",
    "video_encoder.layers.0.blocks.1.attn.qkv.weight",
    "video_encoder.layers.2.blocks.0.mlp.fc1.bias",
    "video_encoder.layers.1.downsample.reduction.weight",
    "video_encoder.norm.weight",
    "audio_embeddings.cls_token",
    "audio_embeddings.position_embeddings.weight",
    "audio_encoder.layer.0.attention.linears.1.weight",
    "audio_encoder.layer.1.ff_layer.linear2.bias",
    "multimodal_encoder.embeddings.word_embeddings.weight",
    "multimodal_encoder.embeddings.position_embeddings.weight",
    "multimodal_encoder.encoder.layer.0.attention.self.query.weight",
    "multimodal_encoder.encoder.layer.1.cross_attn.cross.key.weight",
    "multimodal_encoder.encoder.layer.1.cross_attn.output.LayerNorm.weight",
    "multimodal_encoder.encoder.layer.0.output.dense.bias",
    "cls.dense.weight", "cls.decoder.bias",
    "hidden_trans_video_multimodal.0.weight",
    "contra_head_t.linear.weight", "contra_head_v.linear.weight", "contra_head_a.linear.weight",
    "text_fine_weight.0.weight", "video_fine_weight.2.weight", "audio_fine_weight.2.bias",
    "video_frame_embedding", "audio_type_embeddings", "contra_temp",
]


def stats(t):
    t = t.detach().float().flatten()
    return {"shape": None, "mean": t.mean().item(), "abs_mean": t.abs().mean().item(),
            "std": t.std().item() if t.numel() > 1 else 0.0, "head": t[:8].tolist()}


def run(name, geom, B, F, A, T, task=TASK):
    sd = synth.make_state_dict(geom, seed=0)
    batch = synth.make_batch(B, F, A, T, geom, seed=123)
    tokens = batch["txt_tokens"]["bert_tokens"]
    txt_input, txt_labels = synth.token_masker(tokens, 0.6, seed=1234)
    model = ref_shim.build_reference_valor(geom, sd)
    model.text_masker = ref_shim.FixedMasker(txt_input, txt_labels)
    hooks = {}
    model.video_encoder.register_forward_hook(lambda m, i, o: hooks.__setitem__("swin_out", o))
    model.audio_encoder.register_forward_hook(lambda m, i, o: hooks.__setitem__("ast_out", o[0]))
    calls = []
    model.multimodal_encoder.register_forward_hook(lambda m, i, o: calls.append(o))
    cls_out = []   # masked-token logits of the tva / tv / ta caption passes (pretrain.py:441-443)
    model.cls.register_forward_hook(lambda m, i, o: cls_out.append(o.detach().clone()))
    with ref_shim.cuda_identity():
        loss_dict = model(batch, task, compute_loss=True)
        loss = sum(loss_dict.values())
        loss.backward()
    named = dict(model.named_parameters())
    names = [n for n in ("tva", "tv", "ta") if n in [t for t in task.split("_") if "caption" in t][0].split("%")[1:]]
    gname = "tiny" if geom is synth.TINY else "c1"
    out = {"config": dict(B=B, F=F, A=A, T=T, geom=gname, task=task, weight_seed=0, batch_seed=123, mask_seed=1234),
           "losses": {k: v.item() for k, v in loss_dict.items()},
           "acts": {"swin_out": stats(hooks["swin_out"]), "txt_out": stats(calls[0])},
           "grads": {}}
    if "ast_out" in hooks:
        out["acts"]["ast_out"] = stats(hooks["ast_out"])
    for i, nm in enumerate(names):
        out["acts"][f"fusion_{nm}_out"] = stats(calls[1 + i])
    # masked-token logits: per caption pass, the first 6 masked rows x (first 8 vocabulary
    # columns + the label column), row-wise logsumexp and the argmax -- enough to pin logits without a 70 MB fixture
    lab = txt_labels[txt_labels != -1]
    out["logits"] = {}
    for nm, sc in zip(names, cls_out):
        rows = min(6, sc.shape[0])
        out["logits"][nm] = {"n_rows": sc.shape[0], "head": sc[:rows, :8].tolist(),
                             "at_label": sc[torch.arange(rows), lab[:rows]].tolist(),
                             "lse": torch.logsumexp(sc[:rows].float(), -1).tolist(),
                             "argmax": sc[:rows].argmax(-1).tolist()}
    total = 0.0
    for k, p in named.items():
        if p.grad is not None:
            total += p.grad.double().pow(2).sum().item()
    out["grad_total_norm"] = total ** 0.5
    for k in GRAD_KEYS:
        kk = k
        if kk not in named:  # shared module is registered under its first name
            continue
        g = named[kk].grad
        # float64 norms: torch's CPU fp32 reduction loses ~1e-3 on the 23 M-element embedding matrix
        out["grads"][k] = None if g is None else {"norm": g.double().norm().item(), "head": g.flatten()[:6].tolist()}
    out["unused_params"] = sorted(k for k, p in named.items() if p.grad is None)
    out["n_params"] = sum(p.numel() for p in named.values())
    if name in TRAJECTORY:
        out["trajectory"] = trajectory(model, batch, TRAJECTORY[name], task)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"golden_{name}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(name, out["losses"], "grad_norm", out["grad_total_norm"], "unused", len(out["unused_params"]))


def trajectory(model, batch, n_steps, task=TASK):
    """n_steps of the reference's own step tail (train_utils.py:306-363): build_optimizer (optim/misc.py:13-77) ->
    optim/adamw.py AdamW, get_lr_sched (optim/sched.py:37-41), clip_grad_norm_(5.0), optimizer.step, zero_grad.
    The gradients of step 0 are already in .grad (the caller ran forward/backward).  amp O2 is the identity in
    fp32 on CPU (loss scale 1, masters = params)."""
    import inspect
    from optim.misc import build_optimizer
    from optim.sched import get_lr_sched
    from torch.nn.utils import clip_grad_norm_
    import optim.adamw as ref_adamw
    opts = ref_shim.default_opts()
    opts.num_train_steps = 1000
    # torch >= 2 removed the (Number, Tensor) overloads the 2019 optimizer uses (adamw.py:72-73,84,95); run its
    # step() source unchanged through thin shims that restore them
    _patch_legacy_overloads()
    optimizer = build_optimizer(model, opts)
    for g in optimizer.param_groups:
        g["init_lr"] = g["lr"]              # train_utils.py:237-238
    named = dict(model.named_parameters())
    rec = {"steps": []}
    for step in range(1, n_steps + 1):
        if step > 1:
            with ref_shim.cuda_identity():
                loss_dict = model(batch, task, compute_loss=True)
                sum(loss_dict.values()).backward()
            cur = {k: v.item() for k, v in loss_dict.items()}
        else:
            cur = None
        lr_ratio = get_lr_sched(step, opts)
        for g in optimizer.param_groups:
            g["lr"] = g["init_lr"] * lr_ratio
        gn = clip_grad_norm_([p for g in optimizer.param_groups for p in g["params"]], opts.grad_norm)
        optimizer.step()
        optimizer.zero_grad()
        rec["steps"].append({"losses_before": cur, "lr_ratio": lr_ratio, "grad_norm": float(gn)})
    with ref_shim.cuda_identity(), torch.no_grad():
        loss_dict = model(batch, task, compute_loss=True)
    rec["final_losses"] = {k: v.item() for k, v in loss_dict.items()}
    rec["params"] = {k: {"norm": named[k].detach().double().norm().item(), "head": named[k].detach().flatten()[:6].tolist()}
                     for k in GRAD_KEYS + ["multimodal_encoder.pooler.dense.weight"] if k in named}
    return rec


def _patch_legacy_overloads():
    T = torch.Tensor
    if getattr(T, "_valor_legacy", False):
        return
    o_add, o_addcmul, o_addcdiv = T.add_, T.addcmul_, T.addcdiv_

    def add_(self, a, b=None, **kw):
        if b is not None and not torch.is_tensor(a):
            return o_add(self, b, alpha=a)
        return o_add(self, a, **kw) if b is None else o_add(self, a, b, **kw)

    def addcmul_(self, a, b, c=None, **kw):
        if c is not None and not torch.is_tensor(a):
            return o_addcmul(self, b, c, value=a)
        return o_addcmul(self, a, b, **kw)

    def addcdiv_(self, a, b, c=None, **kw):
        if c is not None and not torch.is_tensor(a):
            return o_addcdiv(self, b, c, value=a)
        return o_addcdiv(self, a, b, **kw)

    T.add_, T.addcmul_, T.addcdiv_ = add_, addcmul_, addcdiv_
    T._valor_legacy = True


if __name__ == "__main__":
    assert ref_shim.reference_available(), "run in the build container (needs /root/reference)"
    torch.manual_seed(0)
    for name in (sys.argv[1:] or CONFIGS):
        c = CONFIGS[name]
        run(name, c["geom"], c["B"], c["F"], c["A"], c["T"], c.get("task", TASK))
