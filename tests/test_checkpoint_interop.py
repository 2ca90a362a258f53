"""Checkpoint interop: files written by valor_b200.checkpoint load into the LIVE reference
(model/pretrain.py VALOR, optim/misc.py build_optimizer + optim/adamw.py AdamW) and back.  Needs /root/reference
(build container); the round-trip half runs anywhere."""
import json
import os

import pytest
import torch

from oracle import ref_shim
from tools import synth
from tests.test_host_logic import build

HERE = os.path.dirname(os.path.abspath(__file__))


def _trained_model(cpu_kernels):
    from valor_b200.optim import get_lr_sched
    from valor_b200.pretrain import default_opts
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    model, batch = build(golden["config"])
    opts = default_opts(num_train_steps=1000)
    for i in range(2):
        losses = model(batch, golden["config"]["task"], compute_loss=True)
        model.store.zero_grad()
        sum(losses.values()).backward()
        model.store.set_hyper(get_lr_sched(i + 1, opts))
        model.store.optimizer_step(max_norm=5.0)
    return golden, model, batch, opts


def test_round_trip_and_frame_embedding_extension(cpu_kernels, tmp_path):
    from valor_b200 import checkpoint as C
    golden, model, batch, opts = _trained_model(cpu_kernels)
    mp, op = C.save_checkpoint(model, opts, str(tmp_path), 2)
    sd = torch.load(mp)
    assert set(sd) == set(model.state_dict())                       # reference key layout incl. txt_encoder.* duplicates
    assert torch.equal(sd["txt_encoder.encoder.layer.0.attention.self.query.weight"],
                       sd["multimodal_encoder.encoder.layer.0.attention.self.query.weight"])
    model2, _ = build(golden["config"])
    before = model2.video_frame_embedding.detach().clone()
    missing, unexpected = C.load_checkpoint(model2, mp, video_sample_num=2, audio_sample_num=1)
    assert not missing and not unexpected
    e = model2.video_frame_embedding.detach()
    assert torch.equal(e[:, :2], sd["video_frame_embedding"][:, :2]) and not torch.equal(e, before)
    assert torch.equal(e[:, 2:], sd["video_frame_embedding"][:, 1:2].expand(-1, 30, -1))    # train_utils.py:143-146
    C.load_optimizer_state(model2, opts, torch.load(op))
    assert model2.store.param_steps == model.store.param_steps
    torch.testing.assert_close(model2.store.exp_avg, model.store.exp_avg)
    torch.testing.assert_close(model2.store.exp_avg_sq, model.store.exp_avg_sq)
    with torch.no_grad():   # same weights -> same loss (frame slots beyond the 2 used ones do not matter at F=2)
        a = model(batch, golden["config"]["task"], compute_loss=True)
        b = model2(batch, golden["config"]["task"], compute_loss=True)
    for k in a:
        assert abs(a[k].item() - b[k].item()) <= 1e-6 * abs(a[k].item())


@pytest.mark.skipif(not ref_shim.reference_available(), reason="needs /root/reference (build container)")
def test_files_load_into_the_live_reference(cpu_kernels, tmp_path):
    from valor_b200 import checkpoint as C
    golden, model, batch, opts = _trained_model(cpu_kernels)
    mp, op = C.save_checkpoint(model, opts, str(tmp_path), 2)
    sd = torch.load(mp)
    ref = ref_shim.build_reference_valor(synth.TINY, synth.make_state_dict(synth.TINY, seed=5))   # different weights
    own = ref.state_dict()
    assert set(own) == set(sd)
    ref.load_state_dict(sd, strict=True)                                                             # reference loader
    ti, tl = synth.token_masker(batch["txt_tokens"]["bert_tokens"], 0.6, seed=golden["config"]["mask_seed"])
    ref.text_masker = ref_shim.FixedMasker(ti, tl)
    with ref_shim.cuda_identity(), torch.no_grad():
        want = ref(batch, golden["config"]["task"], compute_loss=True)
        got = model(batch, golden["config"]["task"], compute_loss=True)
    for k in want:
        assert abs(want[k].item() - got[k].item()) <= 5e-5 * abs(want[k].item()), (k, want[k].item(), got[k].item())
    # optimizer file -> the reference's AdamW (optim/misc.py:13-92 group layout, optim/adamw.py state fields)
    import sys
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from optim.misc import build_optimizer
    ropts = ref_shim.default_opts()
    optimizer = build_optimizer(ref, ropts)
    optimizer.load_state_dict(torch.load(op))
    named = dict(ref.named_parameters())
    st = model.store
    for name in ("video_encoder.layers.0.blocks.0.attn.qkv.weight", "multimodal_encoder.encoder.layer.1.output.LayerNorm.bias",
                 "contra_temp"):
        s = optimizer.state[named[name]]
        off, k = st.offsets[name]
        assert s["step"] == 2
        torch.testing.assert_close(s["exp_avg"].reshape(-1), st.exp_avg[off:off + k])
    assert named["multimodal_encoder.pooler.dense.weight"] not in optimizer.state      # never stepped (no gradient)
