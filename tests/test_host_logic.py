"""Host-logic tests (CPU, no GPU compute): the product's module wiring, hand-written backward
compositions and parameter arenas, driven through tests/cpu_backend.py, must reproduce the golden
vectors minted from the live reference."""
import json
import os

import pytest
import torch

from oracle import synth

HERE = os.path.dirname(os.path.abspath(__file__))


class FixedMasker(torch.nn.Module):
    def __init__(self, a, b):
        super().__init__()
        self.a, self.b = a, b

    def forward(self, tokens, p):
        return self.a.clone(), self.b.clone()


def build(cfg, dtype=torch.float32, device="cpu"):
    from valor_b200.pretrain import VALOR, default_opts
    geom = {"tiny": synth.TINY, "c1": synth.BASE}[cfg["geom"]]
    opts = default_opts(swin_depths=geom.swin_depths, ast_layers=geom.ast_layers, bert_layers=geom.bert_layers)
    sd = synth.make_state_dict(geom, seed=cfg["weight_seed"])
    model = VALOR.from_pretrained(opts, sd)
    model.attach(dtype=dtype, device=device)
    batch = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
    tokens = batch["txt_tokens"]["bert_tokens"]
    ti, tl = synth.token_masker(tokens, 0.6, seed=cfg["mask_seed"])
    model.text_masker = FixedMasker(ti.to(device), tl.to(device))
    batch = {"video_pixels": batch["video_pixels"].to(device), "audio_spectrograms": batch["audio_spectrograms"].to(device),
             "txt_tokens": {"bert_tokens": tokens.to(device)}, "ids": batch["ids"]}
    return model, batch


def test_state_dict_contract():
    from valor_b200.pretrain import VALOR, default_opts
    m = VALOR(default_opts(swin_depths=(2, 2, 2, 2), ast_layers=2, bert_layers=2))
    sd = synth.make_state_dict(synth.TINY)
    own = m.state_dict()
    assert set(own) == set(sd)
    for k in own:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    # shared / tied tensors (modeling.py:241,690)
    assert m.txt_encoder is m.multimodal_encoder
    assert m.cls.decoder.weight is m.multimodal_encoder.embeddings.word_embeddings.weight


def test_forward_backward_matches_reference_golden(cpu_kernels):
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"])
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    for k, v in golden["losses"].items():
        assert abs(losses[k].item() - v) <= 5e-5 * abs(v), (k, losses[k].item(), v)
    model.store.zero_grad()
    sum(losses.values()).backward()
    total = model.store.grad.double().pow(2).sum().sqrt().item()
    assert abs(total - golden["grad_total_norm"]) <= 2e-4 * golden["grad_total_norm"], (total, golden["grad_total_norm"])
    named = dict(model.named_parameters())
    for k, ref in golden["grads"].items():
        g = named[k].main_grad
        if ref is None:
            assert g.abs().sum().item() == 0.0, k
            continue
        assert abs(g.norm().item() - ref["norm"]) <= 1e-3 * ref["norm"] + 1e-9, (k, g.norm().item(), ref["norm"])
        torch.testing.assert_close(g.flatten()[:6], torch.tensor(ref["head"]), rtol=5e-3, atol=1e-6)
