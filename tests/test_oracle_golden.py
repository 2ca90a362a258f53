"""Pins oracle/valor_oracle.py against golden vectors minted from the LIVE reference
(tests/golden/make_golden.py): losses, activations statistics, gradient norms."""
import json
import os

import pytest
import torch

from oracle import synth, valor_oracle as vo

HERE = os.path.dirname(os.path.abspath(__file__))
GEOMS = {"tiny": synth.TINY, "c1": synth.BASE}


def run_oracle(cfg, want_grads=True):
    geom = GEOMS[cfg["geom"]]
    sd = synth.make_state_dict(geom, seed=cfg["weight_seed"], include_buffers=False)
    params = {}
    for k, v in sd.items():
        if k.startswith("txt_encoder.") or k == "cls.decoder.weight":
            continue
        params[k] = v.clone().requires_grad_(want_grads)
    full = dict(params)
    for k in sd:
        if k.startswith("txt_encoder."):
            full[k] = params["multimodal_encoder." + k[len("txt_encoder."):]]
    full["cls.decoder.weight"] = params["multimodal_encoder.embeddings.word_embeddings.weight"]
    batch = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
    tokens = batch["txt_tokens"]["bert_tokens"]
    txt_input, txt_labels = synth.token_masker(tokens, 0.6, seed=cfg["mask_seed"])
    losses, aux = vo.forward_pt(batch, full, geom, txt_input, txt_labels, task=cfg["task"], return_aux=True)
    if want_grads:
        sum(losses.values()).backward()
    return losses, aux, params


def _check_stats(t, ref, rtol=2e-4):
    t = t.detach().float().flatten()
    assert abs(t.mean().item() - ref["mean"]) <= rtol * max(1e-3, abs(ref["abs_mean"]))
    assert abs(t.abs().mean().item() - ref["abs_mean"]) <= rtol * ref["abs_mean"]
    torch.testing.assert_close(t[:8], torch.tensor(ref["head"]), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", ["tiny", pytest.param("c1", marks=pytest.mark.slow)])
def test_oracle_matches_reference_golden(name):
    path = os.path.join(HERE, "golden", f"golden_{name}.json")
    golden = json.load(open(path))
    losses, aux, params = run_oracle(golden["config"])
    for k, v in golden["losses"].items():
        assert abs(losses[k].item() - v) <= 2e-5 * abs(v), (k, losses[k].item(), v)
    _check_stats(aux["video_output"].permute(0, 3, 1, 2), golden["acts"]["swin_out"])
    _check_stats(aux["audio_output"], golden["acts"]["ast_out"])
    _check_stats(aux["txt_output"], golden["acts"]["txt_out"])
    total = sum(p.grad.double().pow(2).sum().item() for p in params.values() if p.grad is not None) ** 0.5
    assert abs(total - golden["grad_total_norm"]) <= 1e-4 * golden["grad_total_norm"]
    for k, ref in golden["grads"].items():
        g = params[k].grad
        if ref is None:
            assert g is None or g.abs().sum().item() == 0.0, k
            continue
        assert abs(g.norm().item() - ref["norm"]) <= 5e-4 * ref["norm"] + 1e-9, (k, g.norm().item(), ref["norm"])
        torch.testing.assert_close(g.flatten()[:6], torch.tensor(ref["head"]), rtol=2e-3, atol=1e-7)
    unused = sorted(k for k, p in params.items() if p.grad is None)
    assert unused == golden["unused_params"], (unused, golden["unused_params"])
