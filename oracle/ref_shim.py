"""TEST INFRASTRUCTURE ONLY — loader for the LIVE reference (TXH-mercury/VALOR) on CPU.

This file never ships with the product path.  It exists so that
`tests/golden/make_golden.py` (run in the build container, where /root/reference is
mounted read-only) can execute the reference's own `model/pretrain.py: VALOR` on CPU and
mint golden vectors, and so that `oracle/valor_oracle.py` (the CPU restatement that DOES
travel to the GPU box) can be pinned against it.

Recipe: fake `apex`/`ipdb`/`tensorboardX`/`easydict`/`boto3` modules,
`Tensor.cuda -> identity`, gloo world_size 1, a synthetic `pretrained_weights/` directory,
`torch.load` patched by file name, repaired opts (`fineweight_type=None`).  Nothing under
/root/reference is modified or copied.
"""
import argparse
import contextlib
import json
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VALOR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def _install_fake_modules():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "apex" not in sys.modules:
        apex = mod("apex")
        norm = mod("apex.normalization")
        fln = mod("apex.normalization.fused_layer_norm")

        class FusedLayerNorm(nn.LayerNorm):
            # same math as apex's own CPU fallback (fused_layer_norm.py:154-156 -> F.layer_norm)
            def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
                super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)

        fln.FusedLayerNorm = FusedLayerNorm
        norm.fused_layer_norm = fln
        norm.FusedLayerNorm = FusedLayerNorm
        apex.normalization = norm
        amp = mod("apex.amp")
        apex.amp = amp
        par = mod("apex.parallel")
        par.DistributedDataParallel = object
        apex.parallel = par
    for name in ("ipdb", "boto3", "botocore", "botocore.exceptions", "tensorboardX", "easydict", "toolz",
                 "toolz.sandbox", "ftfy"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name)
    sys.modules["botocore.exceptions"].ClientError = getattr(
        sys.modules["botocore.exceptions"], "ClientError", Exception)
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    if not hasattr(sys.modules["tensorboardX"], "SummaryWriter"):
        sys.modules["tensorboardX"].SummaryWriter = object
    if not hasattr(sys.modules["easydict"], "EasyDict"):
        sys.modules["easydict"].EasyDict = dict
    # scorer imports a java-backed tokenizer; VALOR never calls it on the pretraining path
    if "scorer" not in sys.modules:
        sc = mod("scorer")
        scs = mod("scorer.scorer")
        scs.Scorer = object
        sc.scorer = scs


def default_opts(**over):
    """argparse defaults (train_utils.py:601-695) + repaired pretrain-VALOR-base.json +
    scripts/pretrain.sh overrides + fineweight_type=None."""
    o = dict(
        video_resolution=224, audio_melbins=64, audio_patch_size=16, audio_frame_shift=10,
        audio_target_length=512, audio_mean=-4.2677393, audio_std=4.5689974,
        video_encoder_type="videoswin_base_k600_22k", txt_encoder_type="bert_base_uncased",
        audio_encoder_type="ast", multimodal_encoder_type="bert_base_uncased",
        share_txt_and_multimodal=True, multimodal_use_cross_attn=True, cross_attn_type="va_concate",
        contra_type="fine", caption_type="unimlm", feature_pooling_type="none",
        initial_multimodal=True, initial_vision=True, checkpointing=False, frozen_vision=False,
        frozen_multimodal=False, max_generation_len=30, beam_size=3, beam_size_qa=1, label_smoothing=0.0,
        evaluate_ret_text=False, scst_finetuning=False, full_masker=False, contra_loss_ratio=1.5,
        fineweight_type=None, use_task_prompt=False, late_fusion=False, init_clip_head=True,
        contra_dim=512, dual_softmax=False, learning_rate=1e-4, weight_decay=0.01, betas=[0.9, 0.98],
        optim="adamw", grad_norm=5.0, warmup_ratio=0.1, new_lr=0.0, new_params_name=[], clip_lr=5e-7,
        clip_lr_text=5e-7, decoder_lr=-1, scheduler="warmup_linear", num_train_steps=1000,
    )
    o.update(over)
    return argparse.Namespace(**o)


@contextlib.contextmanager
def reference_env(swin_sd, ast_sd, bert_sd, bert_config):
    """cwd with synthetic ./pretrained_weights + torch.load patched by file name."""
    _install_fake_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch.distributed as dist
    if not dist.is_initialized():
        # single-process group over a file store: no TCP port, so a pytest parent and a probe subprocess never collide
        store_file = os.path.join(tempfile.mkdtemp(prefix="valor_ref_pg_"), "store")
        dist.init_process_group("gloo", init_method=f"file://{store_file}", rank=0, world_size=1)
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    old_load = torch.load

    def fake_load(f, *a, **k):
        name = os.path.basename(str(f))
        if "videoswin" in name or "video-swin" in name:
            return {k2: v.clone() for k2, v in swin_sd.items()}
        if name == "bert-base-uncased.bin":
            return {k2: v.clone() for k2, v in bert_sd.items()}
        if name.startswith("audioset"):
            return {k2: v.clone() for k2, v in ast_sd.items()}
        return old_load(f, *a, **k)

    torch.load = fake_load
    old_cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="valor_ref_")
    os.makedirs(os.path.join(tmp, "pretrained_weights"))
    with open(os.path.join(tmp, "pretrained_weights", "bert_base_uncased_config.json"), "w") as f:
        json.dump(bert_config, f)
    with open(os.path.join(tmp, "pretrained_weights", "bert-base-uncased-vocab.txt"), "w") as f:
        vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
        vocab += [f"w{i}" for i in range(bert_config["vocab_size"] - len(vocab))]
        f.write("\n".join(vocab) + "\n")
    os.chdir(tmp)
    try:
        yield
    finally:
        os.chdir(old_cwd)
        torch.load = old_load
        torch.Tensor.cuda = old_cuda


def build_reference_valor(geom, state_dict, opts=None):
    """Instantiate the reference VALOR (model/pretrain.py:64) for geometry `geom`
    (oracle.synth.Geometry) and load `state_dict` (reference key names) strictly for every
    key the reference owns.  Returns the nn.Module in train() mode semantics EXCEPT that
    stochastic layers are disabled (parity mode: Dropout p=0, DropPath off)."""
    from tools import synth
    opts = opts or default_opts(audio_melbins=geom.audio_melbins, audio_target_length=geom.audio_frames,
                                video_resolution=geom.resolution, contra_dim=geom.contra_dim)
    swin_sd = {k[len("video_encoder."):]: v for k, v in state_dict.items() if k.startswith("video_encoder.")}
    ast_sd = synth.ast_checkpoint_from_state(state_dict, geom)
    bert_sd = synth.bert_checkpoint_from_state(state_dict, geom)
    with reference_env(swin_sd, ast_sd, bert_sd, geom.bert_config_json()):
        import functools
        import model.videoswin as ref_swin
        import model.modeling as ref_modeling
        orig_swin = ref_swin.SwinTransformer3D
        orig_layers = ref_modeling.base_cfg.num_hidden_layers
        # geometry overrides: the reference hard-codes Swin-B / 12-layer AST by name
        # (modeling.py:576-611); depths/layers are shrunk for the tiny parity config only.
        ref_swin.SwinTransformer3D = functools.partial(orig_swin, depths=list(geom.swin_depths))
        ref_modeling.base_cfg.num_hidden_layers = geom.ast_layers
        try:
            from model.pretrain import VALOR
            model = VALOR.from_pretrained(opts, {})
        finally:
            ref_swin.SwinTransformer3D = orig_swin
            ref_modeling.base_cfg.num_hidden_layers = orig_layers
    own = model.state_dict()
    missing = [k for k in own if k not in state_dict]
    extra = [k for k in state_dict if k not in own]
    assert not extra, f"synthetic state has keys the reference does not own: {extra[:8]}"
    # buffers (relative_position_index) may be absent from synthetic dicts
    assert all("relative_position_index" in k for k in missing), f"missing: {missing[:8]}"
    model.load_state_dict(state_dict, strict=False)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    model.train()
    return model


@contextlib.contextmanager
def cuda_identity():
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = old_cuda


class FixedMasker(nn.Module):
    """Stands in for TokenMasker (modeling.py:122-174) with a precomputed draw so both sides
    see the same masked tokens (the reference draws from Python `random` on the host)."""

    def __init__(self, txt_input, txt_labels):
        super().__init__()
        self.txt_input, self.txt_labels = txt_input, txt_labels

    def forward(self, tokens, mask_prob):
        return self.txt_input.clone(), self.txt_labels.clone()
