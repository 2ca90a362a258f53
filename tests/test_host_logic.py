"""Host-logic tests (CPU, no GPU compute): the product's module wiring, hand-written backward
compositions and parameter arenas, driven through tests/cpu_backend.py, must reproduce the golden
vectors minted from the live reference."""
import json
import os

import pytest
import torch

from tools import synth

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
    model.set_stochastic(False)          # parity mode: Dropout / DropPath disabled on both sides (see make_golden.py)
    batch = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
    tokens = batch["txt_tokens"]["bert_tokens"]
    ti, tl = synth.token_masker(tokens, 0.6, seed=cfg["mask_seed"])
    model.text_masker = FixedMasker(ti.to(device), tl.to(device))
    batch = {"video_pixels": batch["video_pixels"].to(device), "audio_spectrograms": batch["audio_spectrograms"].to(device),
             "txt_tokens": {"bert_tokens": tokens.to(device)}, "ids": batch["ids"]}
    return model, batch


def masked_logits(model, golden):
    """rows of our [B*npass*T, V] sample-major logits that the reference's `out[labels != -1]` selects, per pass"""
    cap = model.debug_capture
    npass, names = cap["npass"], cap["names"]
    lab = cap["labels"].view(-1, npass, golden["config"]["T"])          # [B, npass, T]
    lg = cap["logits"].view(lab.shape[0], npass, golden["config"]["T"], -1)
    return {nm: lg[:, i][lab[:, i] != -1].float() for i, nm in enumerate(names)}


def check_logits(model, golden, rtol, atol, exact_argmax=True):
    got = masked_logits(model, golden)
    for nm, ref in golden["logits"].items():
        sc = got[nm].cpu()
        assert sc.shape[0] == ref["n_rows"], (nm, sc.shape, ref["n_rows"])
        rows = len(ref["lse"])
        torch.testing.assert_close(sc[:rows, :8], torch.tensor(ref["head"]), rtol=rtol, atol=atol)
        torch.testing.assert_close(torch.logsumexp(sc[:rows], -1), torch.tensor(ref["lse"]), rtol=rtol, atol=atol)
        if exact_argmax:
            assert sc[:rows].argmax(-1).tolist() == ref["argmax"], nm


def run_trajectory(model, batch, golden, loss_rtol, gn_rtol, param_rtol):
    """the reference's step tail (train_utils.py:344-363) through ParamStore, against the recorded trajectory"""
    from valor_b200.optim import get_lr_sched
    from valor_b200.pretrain import default_opts
    opts = default_opts(num_train_steps=1000)
    traj, task = golden["trajectory"], golden["config"]["task"]
    st = model.store
    for i, rec in enumerate(traj["steps"]):
        losses = model(batch, task, compute_loss=True)
        if rec["losses_before"]:
            for k, v in rec["losses_before"].items():
                assert abs(losses[k].item() - v) <= loss_rtol * abs(v), (i, k, losses[k].item(), v)
        st.zero_grad()
        sum(losses.values()).backward()
        st.set_hyper(get_lr_sched(i + 1, opts), base_lr=opts.learning_rate, betas=tuple(opts.betas),
                     weight_decay=opts.weight_decay)
        st.optimizer_step(max_norm=opts.grad_norm)
        assert abs(st.norm[0].item() - rec["grad_norm"]) <= gn_rtol * rec["grad_norm"], (i, st.norm[0].item(), rec["grad_norm"])
    with torch.no_grad():
        losses = model(batch, task, compute_loss=True)
    for k, v in traj["final_losses"].items():
        assert abs(losses[k].item() - v) <= loss_rtol * abs(v), (k, losses[k].item(), v)
    named = dict(model.named_parameters())
    for k, ref in traj["params"].items():
        assert abs(named[k].data.double().norm().item() - ref["norm"]) <= param_rtol * ref["norm"] + 1e-9, \
            (k, named[k].data.double().norm().item(), ref["norm"])


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


@pytest.mark.parametrize("name", ["tiny", "tiny_tv"])
def test_unused_parameters_match_the_reference(name):
    """the set the reference's autograd leaves at grad=None (DDP find_unused_parameters, AdamW skip)"""
    from valor_b200.pretrain import VALOR, default_opts
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    m = VALOR(default_opts(swin_depths=(2, 2, 2, 2), ast_layers=2, bert_layers=2))
    assert m.unused_parameter_names(golden["config"]["task"]) == golden["unused_params"]


@pytest.mark.parametrize("name", ["tiny", "tiny_tv"])
def test_optimizer_trajectory_matches_reference(cpu_kernels, name):
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"])
    run_trajectory(model, batch, golden, loss_rtol=1e-4, gn_rtol=3e-4, param_rtol=1e-5)


@pytest.mark.parametrize("name", ["tiny", "tiny_tv"])
def test_forward_backward_matches_reference_golden(cpu_kernels, name):
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"])
    model.debug_capture = {}
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    check_logits(model, golden, rtol=3e-4, atol=3e-4)
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
        assert abs(g.double().norm().item() - ref["norm"]) <= 1e-3 * ref["norm"] + 1e-9, (k, g.double().norm().item(), ref["norm"])
        torch.testing.assert_close(g.flatten()[:6], torch.tensor(ref["head"]), rtol=5e-3, atol=1e-6)


def test_evaluation_dict_matches_reference_logits(cpu_kernels):
    """compute_loss=False (pretrain.py:402-406,446,463,480,485): features, tokens, per-pass masked-token scores, labels"""
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"])
    with torch.no_grad():
        ev = model(batch, golden["config"]["task"], compute_loss=False)
    cfg = golden["config"]
    assert ev["feat_t"].shape == (cfg["B"], cfg["T"], 512) and ev["feat_v"].shape == (cfg["B"], cfg["F"], 512)
    assert ev["feat_a"].shape == (cfg["B"], cfg["A"], 512) and torch.equal(ev["txt_tokens"], batch["txt_tokens"]["bert_tokens"])
    torch.testing.assert_close(ev["feat_t"].norm(dim=-1), torch.ones(cfg["B"], cfg["T"]), rtol=1e-5, atol=1e-5)
    for nm, ref in golden["logits"].items():
        sc = ev[f"caption_scores_{nm}"].float()
        assert sc.shape[0] == ref["n_rows"]
        rows = len(ref["lse"])
        torch.testing.assert_close(sc[:rows, :8], torch.tensor(ref["head"]), rtol=3e-4, atol=3e-4)
    assert (ev["txt_labels_caption"] != -1).sum().item() == golden["logits"]["tva"]["n_rows"]


def test_stochastic_mode_gradients_match_finite_differences(cpu_kernels):
    """Dropout / DropPath wiring (functional.DropoutAddFn / the DropPath factor in the GEMM epilogue): with the generator state pinned the masks
    are a fixed function of the call site, so the hand-written backward must agree with central differences."""
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"])
    task = golden["config"]["task"]
    model.set_stochastic(True, seed=3)

    def loss():
        model.rng._host[1] = 0            # same counter range -> same masks in every evaluation
        out = model(batch, task, compute_loss=True)
        return out, sum(out.values())

    out, total = loss()
    for k, v in golden["losses"].items():
        assert abs(out[k].item() - v) > 1e-5 * abs(v)          # the masks are really applied
    model.store.zero_grad()
    total.backward()
    named = dict(model.named_parameters())
    probes = [("video_encoder.layers.2.blocks.1.mlp.fc2.bias", 5), ("audio_encoder.layer.1.attention.linears.3.bias", 9),
              ("multimodal_encoder.encoder.layer.1.cross_attn.output.dense.bias", 17),
              ("multimodal_encoder.embeddings.LayerNorm.bias", 3)]
    for name, idx in probes:
        p, g = named[name], named[name].main_grad.flatten()[idx].item()
        eps = 5e-2
        with torch.no_grad():
            p.data.view(-1)[idx] += eps
            lp = loss()[1].item()
            p.data.view(-1)[idx] -= 2 * eps
            lm = loss()[1].item()
            p.data.view(-1)[idx] += eps
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - g) <= 3e-2 * max(abs(g), abs(fd)) + 1e-4, (name, fd, g)


def test_retrieval_path_matches_the_oracle(cpu_kernels):
    """forward_ret evaluation dict + validate_ret scoring + on-device recall (test.py:249-411,714-775) against the
    oracle's compute_fine_matrix + sort/index ranking on the same features; several captions per clip."""
    from oracle import valor_oracle as vo
    from valor_b200 import retrieval as R
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"])
    geom = synth.TINY
    batches = []
    for i in range(3):     # 3 batches x 2 captions; clips repeat so that some clips own two captions
        b = synth.make_batch(2, 2, 1, 16, geom, seed=200 + i)
        b["ids"] = [f"clip{(2 * i) % 4}", f"clip{(2 * i + 1) % 4}"]
        batches.append(b)
    ev = [model(b, "ret%tva%tv", compute_loss=False) for b in batches]
    assert set(ev[0]) == {"feat_t", "feat_v", "feat_a", "txt_tokens"}
    # candidates = one entry per distinct clip (first occurrence), captions = all 6
    ids_txt = [i for b in batches for i in b["ids"]]
    first = {}
    for n, cid in enumerate(ids_txt):
        first.setdefault(cid, n)
    ids = list(first)
    feat_t = torch.cat([e["feat_t"] for e in ev]); toks = torch.cat([e["txt_tokens"] for e in ev])
    feat_v = torch.cat([e["feat_v"] for e in ev])[[first[c] for c in ids]]
    feat_a = torch.cat([e["feat_a"] for e in ev])[[first[c] for c in ids]]
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    for group in ("tva", "tv"):
        want = vo.retrieval_scores(feat_t.float(), feat_v.float(), feat_a.float(), toks, sd, group)
        got = R._fine_scores(model, feat_t, feat_v, feat_a, toks, group)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        for dual in (False, True):
            ref = vo.compute_metric_ret(want, ids, ids_txt, temp=sd["contra_temp"], dual_softmax=dual, evaluate_ret_text=True)
            log = R.compute_metric_ret(model, got, ids, ids_txt, dual_softmax=dual, evaluate_ret_text=True)
            assert log == ref, (group, dual, log, ref)
    # the loss side of forward_ret: the contrastive loss without the pretraining ratio (pretrain.py:699)
    l_ret = model(batches[0], "ret%tva%tv", compute_loss=True)["contra_loss"].item()
    l_pt = model(batches[0], "pt_contra%tva%tv", compute_loss=True)["contra_loss"].item()
    assert abs(l_ret * 1.5 - l_pt) <= 1e-6 * abs(l_pt)
