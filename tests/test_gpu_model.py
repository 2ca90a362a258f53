"""End-to-end GPU parity: the CUDA path (VALOR.forward + backward through the C ABI) against the
golden vectors minted from the live reference (tests/golden/) and against the oracle.

 * fp32 parity mode: losses within 1e-4 relative, gradient norms within 2e-3;
 * bf16 perf mode:  losses within 1e-3 relative fp32 (the north-star tolerance), total gradient
   norm within 3e-2.
"""
import json
import os

import pytest
import torch

from tests.test_host_logic import build, check_logits, run_trajectory

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run(name, dtype):
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"], dtype=dtype, device="cuda")
    model.debug_capture = {}
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    model.store.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    return golden, model, {k: v.item() for k, v in losses.items()}


@pytest.mark.parametrize("name", ["tiny", "tiny_tv", "c1", "c2shape"])
def test_fp32_parity_mode_matches_reference(name):
    golden, model, losses = run(name, torch.float32)
    check_logits(model, golden, rtol=1e-3, atol=1e-3)
    for k, v in golden["losses"].items():
        assert abs(losses[k] - v) <= 1e-4 * abs(v), (k, losses[k], v)
    total = model.store.grad.double().pow(2).sum().sqrt().item()
    assert abs(total - golden["grad_total_norm"]) <= 2e-3 * golden["grad_total_norm"], (total, golden["grad_total_norm"])
    named = dict(model.named_parameters())
    for k, ref in golden["grads"].items():
        g = named[k].main_grad
        if ref is None:
            assert g.abs().sum().item() == 0.0, k
            continue
        assert abs(g.double().norm().item() - ref["norm"]) <= 5e-3 * ref["norm"] + 1e-8, (k, g.double().norm().item(), ref["norm"])


@pytest.mark.parametrize("name", ["tiny", "c1", "c2shape"])
def test_bf16_perf_mode_per_parameter_gradients(name):
    """bf16 perf mode on the production kernels (tcgen05 GEMM, tensor-core attention): every tracked parameter
    gradient against the live reference's fp32 gradient, relative per tensor (norm and leading elements), and
    the masked-token logits."""
    import valor_b200.kernels as K
    golden, model, losses = run(name, torch.bfloat16)
    named = dict(model.named_parameters())
    worst = 0.0
    total = golden["grad_total_norm"]
    for k, ref in golden["grads"].items():
        g = named[k].main_grad
        if ref is None:
            assert g.abs().sum().item() == 0.0, k
            continue
        dev = abs(g.double().norm().item() - ref["norm"])
        rel = dev / (ref["norm"] + 1e-12)
        worst = max(worst, rel)
        # relative per tensor; tensors whose whole gradient is below 1e-3 of the step's gradient norm (the scalar
        # temperature, the single-slot audio fine weight at A=1: analytically ~0) are held to that absolute floor
        # (the scalar temperature sees every similarity through 1/temp^2 = 204: at the two-sample configuration its bf16
        # noise is a few 1e-3 of the step's gradient norm; torch-eager bf16 shows the same, tools/parity_report.py)
        floor = (3e-3 if g.numel() == 1 else 1e-3) * total
        assert rel <= 5e-2 or dev <= floor, (k, g.double().norm().item(), ref["norm"], total)
        if dev <= floor and ref["norm"] <= 1e-2 * total:
            continue        # analytically (near-)zero gradients: nothing elementwise to compare beyond the floor above
        head = torch.tensor(ref["head"])
        got = g.flatten()[:6].cpu()
        # leading elements: within 5% of the tensor's RMS magnitude (elementwise bf16 noise is absolute, not relative)
        rms = ref["norm"] / max(1.0, g.numel() ** 0.5)
        assert (got - head).abs().max().item() <= 0.3 * max(rms, head.abs().max().item()), (k, got, head)   # gross-error net; bf16 element noise reaches ~20% in the first Swin layers (torch-eager bf16 shows the same)
    print(f"{name}: worst per-parameter gradient-norm deviation {worst:.2e}")
    # logits: absolute tolerance in logit units (|logit| ~ 1e-1..1e0 through 12+ bf16 layers)
    check_logits(model, golden, rtol=3e-2, atol=3e-2, exact_argmax=False)


@pytest.mark.parametrize("name", ["tiny", "tiny_tv", "c1"])
def test_optimizer_trajectory_matches_reference_fp32(name):
    """3 steps of clip + AdamW (skip-if-no-grad) + LR schedule on the GPU path vs the reference's own optimizer"""
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"], dtype=torch.float32, device="cuda")
    run_trajectory(model, batch, golden, loss_rtol=2e-4, gn_rtol=3e-3, param_rtol=1e-5)


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_optimizer_trajectory_matches_reference_bf16(name):
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"], dtype=torch.bfloat16, device="cuda")
    # B = 2: the contrastive mean has four terms divided by temp = 0.07 (single-step bf16 tolerance 3e-3, see
    # test_bf16_perf_mode_matches_reference), bf16 gradient-norm noise reaches 3-4 %, and Adam's first steps move every
    # element by ~lr * sign(g), so gradient noise on small elements shows up in the loss after three steps: a pure
    # re-ordering of fp32 atomics in one LayerNorm backward moved the final contrastive loss by 0.6 %.  The fp32
    # trajectory test above is the tight one (1e-4); this one fences regressions of the bf16 kernels.
    run_trajectory(model, batch, golden, loss_rtol=1e-2, gn_rtol=5e-2, param_rtol=1e-4)


def test_hot_gemms_run_on_the_tensor_backend(monkeypatch):
    """bf16 perf mode must not silently fall back to the SIMT GEMM: force BACKEND_TENSOR on every GEMM of a full
    forward/backward (the ABI refuses ineligible operands) except the handful of shapes that can never be TMA
    operands (the [rows,1] fine-weight heads and fp32 score matrices)."""
    import valor_b200.kernels as K
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"], dtype=torch.bfloat16, device="cuda")
    orig = K.gemm
    stats = {"tensor": 0, "simt": 0, "simt_shapes": set()}

    def forced(a, b, **kw):
        M, Kd = (a.shape if kw.get("a_kmajor", True) else (a.shape[1], a.shape[0]))
        N = b.shape[0] if kw.get("b_kmajor", True) else b.shape[1]
        eligible = (a.dtype == torch.bfloat16 and N >= 8 and Kd >= 8 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0
                    and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)
        if eligible:
            stats["tensor"] += 1
            return orig(a, b, **{**kw, "backend": K.BACKEND_TENSOR})
        stats["simt"] += 1
        stats["simt_shapes"].add((M, N, Kd, str(a.dtype)))
        return orig(a, b, **kw)

    monkeypatch.setattr(K, "gemm", forced)
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    model.store.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    print("tensor GEMMs", stats["tensor"], "other", stats["simt"], sorted(stats["simt_shapes"]))
    assert stats["tensor"] > 100
    # everything that did not run on tcgen05 is a tiny head: N == 1 outputs or their gradients
    for (M, N, Kd, dt) in stats["simt_shapes"]:
        assert min(M, N, Kd) < 8 or dt == "torch.float32", (M, N, Kd, dt)
    for k, v in golden["losses"].items():
        assert abs(losses[k].item() - v) <= 3e-3 * abs(v)


@pytest.mark.parametrize("name", ["tiny", "c1", "c2shape"])
def test_bf16_perf_mode_loss_within_1e3(name):
    golden, model, losses = run(name, torch.bfloat16)
    print("bf16 losses", losses, "golden", golden["losses"])
    # north-star tolerance: the step's loss (sum of the loss dict, train_utils.py:307) within 1e-3
    # relative of the fp32 reference.  Per-term: caption 1e-3; the contrastive term is a mean over only
    # 2*B=4 log-softmax entries divided by temp=0.07 and carries bf16 feature noise: 3e-3 (measured 1.6e-3).
    tot_ref, tot = sum(golden["losses"].values()), sum(losses.values())
    assert abs(tot - tot_ref) <= 1e-3 * abs(tot_ref), (tot, tot_ref)
    assert abs(losses["caption_loss"] - golden["losses"]["caption_loss"]) <= 1e-3 * golden["losses"]["caption_loss"]
    assert abs(losses["contra_loss"] - golden["losses"]["contra_loss"]) <= 3e-3 * golden["losses"]["contra_loss"]
    total = model.store.grad.double().pow(2).sum().sqrt().item()
    assert abs(total - golden["grad_total_norm"]) <= 3e-2 * golden["grad_total_norm"], (total, golden["grad_total_norm"])


def test_optimizer_step_changes_weights_and_is_finite():
    golden, model, losses = run("tiny", torch.bfloat16)
    st = model.store
    before = st.master.clone()
    st.set_hyper(lr_ratio=0.5)
    st.optimizer_step(max_norm=5.0)
    torch.cuda.synchronize()
    assert torch.isfinite(st.master).all()
    assert (st.master - before).abs().max().item() > 0
    assert abs(st.norm[0].item() - st.grad.double().pow(2).sum().sqrt().item()) <= 1e-3 * st.norm[0].item()
    torch.testing.assert_close(st.lp.float(), st.master, rtol=1e-2, atol=1e-3)


def test_training_mode_regularisation_is_stochastic_reproducible_and_trainable():
    """train() with Dropout 0.1 / DropPath on (the reference's nn.Dropout / DropPath, bert.py:353, transformer.py:78,
    videoswin.py:238): the loss moves away from the parity-mode loss, two passes with the same generator state agree
    (the backward regenerates the forward's masks), consecutive steps differ, gradients stay finite."""
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"], dtype=torch.bfloat16, device="cuda")
    task = golden["config"]["task"]
    model.set_stochastic(True, seed=11)
    a = {k: v.item() for k, v in model(batch, task, compute_loss=True).items()}
    model.rng._host[1] = 0
    model.rng.state.copy_(model.rng._host)
    b = {k: v.item() for k, v in model(batch, task, compute_loss=True).items()}
    # same counter range -> same masks (the loss reductions add with atomics, so the last bit may differ; another mask
    # draw moves the losses by ~1e-2)
    assert all(abs(a[k] - b[k]) <= 1e-5 * abs(a[k]) for k in a), (a, b)
    losses = model(batch, task, compute_loss=True)                              # next step: new masks
    c = {k: v.item() for k, v in losses.items()}
    assert c != a
    model.store.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(model.store.grad).all() and model.store.grad.abs().sum().item() > 0
    assert any(abs(c[k2] - v) > 1e-4 * abs(v) for k2, v in golden["losses"].items())   # regularised ...
    assert all(abs(c[k2] - v) < 0.3 * abs(v) for k2, v in golden["losses"].items())    # ... not broken
    model.eval()
    with torch.no_grad():
        d = {k: v.item() for k, v in model(batch, task, compute_loss=True).items()}
    for k2, v in golden["losses"].items():
        assert abs(d[k2] - v) <= 3e-3 * abs(v)                                   # eval(): regularisation off
