"""Pins oracle/valor_oracle.py against golden vectors minted from the LIVE reference
(tests/golden/make_golden.py): losses, activations statistics, gradient norms."""
import json
import os

import pytest
import torch

from oracle import valor_oracle as vo
from tools import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GEOMS = {"tiny": synth.TINY, "c1": synth.BASE}   # config["geom"]: tiny = reduced depths, c1 = VALOR-base


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


def _check_logits(aux, golden, rtol, atol):
    """masked-token logits of every caption pass against the live reference's"""
    for nm, ref in golden["logits"].items():
        sc = aux[f"caption_scores_{nm}"].detach().float()
        assert sc.shape[0] == ref["n_rows"], (nm, sc.shape, ref["n_rows"])
        rows = len(ref["lse"])
        torch.testing.assert_close(sc[:rows, :8], torch.tensor(ref["head"]), rtol=rtol, atol=atol)
        torch.testing.assert_close(torch.logsumexp(sc[:rows], -1), torch.tensor(ref["lse"]), rtol=rtol, atol=atol)
        assert sc[:rows].argmax(-1).tolist() == ref["argmax"], nm


@pytest.mark.parametrize("name", ["tiny", "tiny_tv", pytest.param("c1", marks=pytest.mark.slow),
                                  pytest.param("c2shape", marks=pytest.mark.slow)])
def test_oracle_matches_reference_golden(name):
    path = os.path.join(HERE, "golden", f"golden_{name}.json")
    golden = json.load(open(path))
    losses, aux, params = run_oracle(golden["config"])
    for k, v in golden["losses"].items():
        assert abs(losses[k].item() - v) <= 2e-5 * abs(v), (k, losses[k].item(), v)
    _check_stats(aux["video_output"].permute(0, 3, 1, 2), golden["acts"]["swin_out"])
    if "ast_out" in golden["acts"]:
        _check_stats(aux["audio_output"], golden["acts"]["ast_out"])
    _check_stats(aux["txt_output"], golden["acts"]["txt_out"])
    _check_logits(aux, golden, rtol=2e-4, atol=2e-4)
    total = sum(p.grad.double().pow(2).sum().item() for p in params.values() if p.grad is not None) ** 0.5
    assert abs(total - golden["grad_total_norm"]) <= 1e-4 * golden["grad_total_norm"]
    for k, ref in golden["grads"].items():
        g = params[k].grad
        if ref is None:
            assert g is None or g.abs().sum().item() == 0.0, k
            continue
        assert abs(g.double().norm().item() - ref["norm"]) <= 5e-4 * ref["norm"] + 1e-9, (k, g.double().norm().item(), ref["norm"])
        torch.testing.assert_close(g.flatten()[:6], torch.tensor(ref["head"]), rtol=2e-3, atol=1e-7)
    unused = sorted(k for k, p in params.items() if p.grad is None)
    assert unused == golden["unused_params"], (unused, golden["unused_params"])


@pytest.mark.parametrize("name", ["tiny", "tiny_tv"])
def test_oracle_optimizer_trajectory_matches_reference(name):
    """k steps of clip_grad_norm_(5.0) + the reference AdamW (skip-if-no-grad, per-parameter step counts,
    optim/adamw.py:50-101) + warmup-linear LR, against the trajectory the live reference recorded."""
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    cfg, traj = golden["config"], golden["trajectory"]
    geom = GEOMS[cfg["geom"]]
    sd = synth.make_state_dict(geom, seed=cfg["weight_seed"], include_buffers=False)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if not k.startswith("txt_encoder.") and k != "cls.decoder.weight"}
    full = dict(params)
    for k in sd:
        if k.startswith("txt_encoder."):
            full[k] = params["multimodal_encoder." + k[len("txt_encoder."):]]
    full["cls.decoder.weight"] = params["multimodal_encoder.embeddings.word_embeddings.weight"]
    batch = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
    ti, tl = synth.token_masker(batch["txt_tokens"]["bert_tokens"], 0.6, seed=cfg["mask_seed"])
    state = {k: [torch.zeros_like(p), torch.zeros_like(p), 0] for k, p in params.items()}
    for i, rec in enumerate(traj["steps"]):
        losses = vo.forward_pt(batch, full, geom, ti, tl, task=cfg["task"])
        if rec["losses_before"]:
            for k, v in rec["losses_before"].items():
                assert abs(losses[k].item() - v) <= 1e-4 * abs(v), (i, k, losses[k].item(), v)
        for p in params.values():
            p.grad = None
        sum(losses.values()).backward()
        lr = 1e-4 * vo.warmup_linear((i + 1) / 1000, 0.1)
        assert abs(vo.warmup_linear((i + 1) / 1000, 0.1) - rec["lr_ratio"]) < 1e-12
        gn = vo.clip_grad_norm_([p.grad for p in params.values() if p.grad is not None], 5.0)
        assert abs(gn.item() - rec["grad_norm"]) <= 2e-4 * rec["grad_norm"]
        with torch.no_grad():
            for k, p in params.items():
                if p.grad is None:
                    continue
                st = state[k]
                st[2] += 1
                vo.adamw_step(p, p.grad, st[0], st[1], st[2], lr, weight_decay=0.0 if vo.is_no_decay(k) else 0.01)
    with torch.no_grad():
        losses = vo.forward_pt(batch, full, geom, ti, tl, task=cfg["task"])
    for k, v in traj["final_losses"].items():
        assert abs(losses[k].item() - v) <= 2e-4 * abs(v), (k, losses[k].item(), v)
    for k, ref in traj["params"].items():
        assert abs(params[k].double().norm().item() - ref["norm"]) <= 1e-5 * ref["norm"] + 1e-9, k
