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
from oracle import ref_shim, synth  # noqa: E402

CONFIGS = {
    "tiny": dict(geom=synth.TINY, B=2, F=2, A=1, T=16),
    "c1": dict(geom=synth.BASE, B=2, F=4, A=1, T=32),
}
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


def run(name, geom, B, F, A, T):
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
    with ref_shim.cuda_identity():
        loss_dict = model(batch, TASK, compute_loss=True)
        loss = sum(loss_dict.values())
        loss.backward()
    named = dict(model.named_parameters())
    out = {"config": dict(B=B, F=F, A=A, T=T, geom=name, task=TASK, weight_seed=0, batch_seed=123, mask_seed=1234),
           "losses": {k: v.item() for k, v in loss_dict.items()},
           "acts": {"swin_out": stats(hooks["swin_out"]), "ast_out": stats(hooks["ast_out"]),
                    "txt_out": stats(calls[0]), "fusion_tva_out": stats(calls[1]),
                    "fusion_tv_out": stats(calls[2]), "fusion_ta_out": stats(calls[3])},
           "grads": {}}
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
        out["grads"][k] = None if g is None else {"norm": g.norm().item(), "head": g.flatten()[:6].tolist()}
    out["unused_params"] = sorted(k for k, p in named.items() if p.grad is None)
    out["n_params"] = sum(p.numel() for p in named.values())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"golden_{name}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(name, out["losses"], "grad_norm", out["grad_total_norm"], "unused", len(out["unused_params"]))


if __name__ == "__main__":
    assert ref_shim.reference_available(), "run in the build container (needs /root/reference)"
    torch.manual_seed(0)
    for name in (sys.argv[1:] or CONFIGS):
        c = CONFIGS[name]
        run(name, c["geom"], c["B"], c["F"], c["A"], c["T"])
