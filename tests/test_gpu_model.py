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

from tests.test_host_logic import build

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run(name, dtype):
    golden = json.load(open(os.path.join(HERE, "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"], dtype=dtype, device="cuda")
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    model.store.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    return golden, model, {k: v.item() for k, v in losses.items()}


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_fp32_parity_mode_matches_reference(name):
    golden, model, losses = run(name, torch.float32)
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
        assert abs(g.norm().item() - ref["norm"]) <= 5e-3 * ref["norm"] + 1e-8, (k, g.norm().item(), ref["norm"])


@pytest.mark.parametrize("name", ["tiny", "c1"])
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
