"""VALORModel — mirror of `model/modeling.py` (VALORModel :281-352, TokenMasker :122-174,
BERTPredictionHead :232-254, encoder wiring :437-502, contrastive_loss :418-433).

Checkpoints are state dicts with the reference's keys; the constructor does NOT read
./pretrained_weights/* (the reference's hard-coded files, modeling.py:514,592-598,616,621) —
`from_pretrained(opts, state_dict)` is the only weight source, as in modeling.py:107-115.
"""
import argparse
import random

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .bert import BertConfig, BertModel
from .functional import LN, lin_of
from .transformer import AudioEmbeddings, TransformerEncoder
from .videoswin import SwinTransformer3D, _Linear, _Norm


class TokenMasker(nn.Module):
    """modeling.py:122-174, same draw order from Python `random` (host-side, as the reference):
    position 0 never masked; every non-zero token (incl. [SEP]) is a candidate; at least one mask
    per row; 80% [MASK] / 10% random id in [range_start, range_end) / 10% unchanged."""

    def __init__(self, mask_token=-1, range_start=-1, range_end=-1):
        super().__init__()
        self.mask_token = mask_token
        self.range = [range_start, range_end]

    def forward(self, tokens, mask_prob):
        device = tokens.device
        tok = np.array(tokens.cpu().numpy())
        ind = np.zeros(tok.shape, dtype=np.int64)
        for i in range(len(ind)):
            while all(ind[i] == 0):
                for j in range(1, len(ind[0])):
                    if tok[i][j] != 0 and random.random() < mask_prob:
                        ind[i][j] = 1
        labels = -np.ones(tok.shape, dtype=np.int64)
        for i in range(tok.shape[0]):
            for j in range(tok.shape[1]):
                if ind[i][j] == 1:
                    src = tok[i][j]
                    prob = random.random()
                    if prob < 0.8:
                        tok[i][j] = self.mask_token
                    elif prob < 0.9:
                        tok[i][j] = random.choice(range(*self.range))
                    labels[i][j] = src
        return torch.from_numpy(tok).long().to(device), torch.from_numpy(labels).long().to(device)


class BERTPredictionHead(nn.Module):
    """modeling.py:232-254: dense -> erf-GELU -> LayerNorm(1e-12) -> decoder tied to the word
    embeddings (:241)."""

    def __init__(self, embedding_weights):
        super().__init__()
        self.hidden_size = embedding_weights.size(1)
        self.vocab_size = embedding_weights.size(0)
        self.dense = _Linear(self.hidden_size, self.hidden_size)
        self.layernorm = _Norm(self.hidden_size, eps=1e-12)
        self.decoder = nn.Module()
        self.decoder.weight = embedding_weights
        self.decoder.bias = nn.Parameter(torch.zeros(self.vocab_size))

    def forward(self, x):
        """x [M, H] -> logits [M, V] (row pitch padded to a multiple of 8 for the vector epilogue)."""
        x = Fn.linear(x, lin_of(self.dense.weight, self.dense.bias), act=K.ACT_GELU)
        x = Fn.layer_norm(x, LN(self.layernorm.weight, self.layernorm.bias, 1e-12))
        return Fn.linear(x, lin_of(self.decoder.weight, self.decoder.bias))


class _Cfg:
    pass


def _audio_cfg(layers):
    """base_cfg, modeling.py:270-278"""
    c = _Cfg()
    c.attention_dropout, c.hidden_act, c.hidden_dropout = 0.1, "gelu", 0.1
    c.hidden_size, c.initializer_range, c.intermediate_size = 768, 0.02, 3072
    c.num_attention_heads, c.num_hidden_layers = 12, layers
    return c


def default_opts(**over):
    """Defaults of train_utils.get_args (:599-695) merged with config/pretrain-VALOR-base.json and
    scripts/pretrain.sh:3-8 (videoswin_base_k600_22k + bert_base_uncased, contra_loss_ratio 1.5)."""
    o = dict(video_resolution=224, audio_melbins=64, audio_patch_size=16, audio_target_length=512,
             video_encoder_type="videoswin_base_k600_22k", txt_encoder_type="bert_base_uncased",
             audio_encoder_type="ast", multimodal_encoder_type="bert_base_uncased", share_txt_and_multimodal=True,
             multimodal_use_cross_attn=True, cross_attn_type="va_concate", contra_type="fine", caption_type="unimlm",
             checkpointing=False, max_generation_len=30, beam_size=3, beam_size_qa=1, label_smoothing=0.0,
             evaluate_ret_text=False, scst_finetuning=False, full_masker=False, contra_loss_ratio=1.5,
             fineweight_type=None, use_task_prompt=False, late_fusion=False, init_clip_head=True, contra_dim=512,
             dual_softmax=False, frozen_vision=False, frozen_multimodal=False, initial_multimodal=True,
             initial_vision=True, learning_rate=1e-4, weight_decay=0.01, betas=[0.9, 0.98], grad_norm=5.0,
             warmup_ratio=0.1, scheduler="warmup_linear", num_train_steps=1000,
             # geometry overrides (not in the reference: depths are hard-coded there by encoder name)
             swin_depths=(2, 2, 18, 2), ast_layers=12, bert_layers=12, vocab_size=30522)
    o.update(over)
    return argparse.Namespace(**o)


class VALORModel(nn.Module):
    def __init__(self, opts):
        super().__init__()
        config = opts
        self.config = config
        self.video_encoder_type = config.video_encoder_type
        self.txt_encoder_type = config.txt_encoder_type
        self.audio_encoder_type = config.audio_encoder_type
        self.multimodal_encoder_type = config.multimodal_encoder_type
        self.multimodal_use_cross_attn = getattr(config, "multimodal_use_cross_attn", False)
        if not self.video_encoder_type.startswith("videoswin"):
            raise NotImplementedError("CLIP-ViT video tower (VALOR-large, model/clip.py) is the next §8 row; "
                                      "VALOR-base as pretrained uses VideoSwin (scripts/pretrain.sh:3-8)")
        if not (self.txt_encoder_type.startswith("bert") and self.multimodal_encoder_type == "bert_base_uncased"
                and config.share_txt_and_multimodal and self.multimodal_use_cross_attn):
            raise NotImplementedError("only the shipped shared BERT text/fusion configuration is built")
        # ---- video (load_videoswin_model, modeling.py:576-587)
        if self.video_encoder_type.startswith("videoswin_small"):
            self.video_encoder = SwinTransformer3D(time_stride=1, embed_dim=96, num_heads=[3, 6, 12, 24],
                                                   depths=list(getattr(config, "swin_depths", (2, 2, 18, 2))))
            self.video_dim = 768
        else:
            self.video_encoder = SwinTransformer3D(time_stride=1, embed_dim=128, num_heads=[4, 8, 16, 32],
                                                   depths=list(getattr(config, "swin_depths", (2, 2, 18, 2))))
            self.video_dim = 1024
        # ---- audio (load_ast_model, modeling.py:605-611)
        acfg = _audio_cfg(getattr(config, "ast_layers", 12))
        self.audio_embeddings = AudioEmbeddings(acfg, config)
        self.audio_encoder = TransformerEncoder(acfg, mode="prenorm")
        self.audio_dim = 768
        # ---- BERT fusion + shared text encoder + MLM head (load_bert_model :613-673, construct_text_model :685-691)
        bcfg = BertConfig(getattr(config, "vocab_size", 30522), num_hidden_layers=getattr(config, "bert_layers", 12))
        bcfg.has_cross_attn, bcfg.cross_attn_type = True, config.cross_attn_type
        self.multimodal_encoder = BertModel(bcfg)
        self.multimodal_dim = 768
        self.cls = BERTPredictionHead(self.multimodal_encoder.embeddings.word_embeddings.weight)
        self.bos_token, self.eos_token, self.text_mask_token = 101, 102, 103  # bert-base-uncased vocab ids
        self.text_masker = TokenMasker(mask_token=self.text_mask_token, range_start=106, range_end=30522)
        self.txt_encoder = self.multimodal_encoder
        self.txt_dim = self.multimodal_dim
        # ---- embeddings / adapters (modeling.py:341-351)
        self.video_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, self.multimodal_dim))
        self.audio_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, self.multimodal_dim))
        self.video_frame_embedding = nn.Parameter(0.02 * torch.randn(1, 32, self.multimodal_dim))
        self.audio_frame_embedding = nn.Parameter(0.02 * torch.randn(1, 32, self.multimodal_dim))
        self.hidden_trans_video_multimodal = None
        self.hidden_trans_audio_multimodal = None
        if self.video_dim != self.multimodal_dim:
            self.hidden_trans_video_multimodal = nn.Sequential(_Linear(self.video_dim, self.multimodal_dim),
                                                               _Norm(self.multimodal_dim, eps=1e-12))
        self.store = None

    # ------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, opts, state_dict, *inputs, **kwargs):
        """modeling.py:107-115"""
        model = cls(opts, *inputs, **kwargs)
        model.load_state_dict(state_dict, strict=False)
        return model

    def attach(self, dtype=torch.bfloat16, device="cuda"):
        """Move to the device and lay the parameters out in the flat arenas (params.ParamStore)."""
        from .params import ParamStore
        self.to(device)
        self.store = ParamStore(self, dtype=dtype, device=torch.device(device))
        self.compute_dtype = dtype
        # stochastic regularisation (Dropout 0.1 in BERT / AST, DropPath 0 -> 0.2 in VideoSwin) is ON in train() mode like
        # the reference's nn.Dropout / DropPath modules; `set_stochastic(False)` is the parity mode of the tests (both
        # sides disabled: the reference draws its masks from torch's host-seeded generator, not reproducible here)
        self.rng = Fn.RngState(torch.device(device), seed=getattr(self.config, "seed", 0))
        for m in self.modules():
            m._rng = self.rng
        return self.store

    def set_stochastic(self, on=True, seed=None):
        self.rng.active = bool(on)
        if seed is not None:
            self.rng._host[0] = int(seed)
            self.rng.state.copy_(self.rng._host)
        return self

    # ---- encoder wiring (token-matrix forms of modeling.py:449-502) -------------------------
    def forward_video_encoder(self, video_pixels):
        """[B,F,3,H,W] -> [B,F,49,C]   (modeling.py:449-455)"""
        tok, (B, F, h, w) = self.video_encoder.forward_tokens(video_pixels, self.compute_dtype)
        return tok.view(B, F, h * w, -1)

    def forward_audio_encoder(self, audio_spectrograms):
        """[B,A,mel,frames] -> [B,A,129,768]   (modeling.py:468-480)"""
        b, n = audio_spectrograms.shape[:2]
        x = self.audio_embeddings.run(audio_spectrograms.reshape(-1, *audio_spectrograms.shape[2:]), self.compute_dtype)
        L = self.audio_embeddings.token_length_per_frame + 1
        x = self.audio_encoder.run(x, b * n, L)
        return x.view(b, n, L, -1)

    def forward_txt_encoder(self, txt_tokens, task_prompt=None):
        """modeling.py:437-446 — shared BERT, cross sub-layers skipped (bert.py:456-457)."""
        B, T = txt_tokens.shape
        return self.txt_encoder.encode(txt_tokens, [False]).view(B, T, -1)

    def media_tokens(self, video_output, audio_output):
        """get_multimodal_forward_input_video + _audio (modeling.py:485-502) fused into the single
        per-sample cross-attention source [B*(Sv+Sa), 768]; returns (media, Sv, Sa)."""
        B = (video_output if video_output is not None else audio_output).shape[0]
        pv = pa = vx = ax = None
        Sv = Sa = 0
        if video_output is not None:
            _, nf, X, C = video_output.shape
            vx = video_output.reshape(-1, C)
            if self.hidden_trans_video_multimodal is not None:
                l0, l1 = self.hidden_trans_video_multimodal[0], self.hidden_trans_video_multimodal[1]
                vx = Fn.layer_norm(Fn.linear(vx, lin_of(l0.weight, l0.bias)), LN(l1.weight, l1.bias, 1e-12))
            pv = (self.video_frame_embedding, self.video_type_embeddings, nf, X)
            Sv = nf * X
        if audio_output is not None:
            _, nf, X, C = audio_output.shape
            ax = audio_output.reshape(-1, C)
            pa = (self.audio_frame_embedding, self.audio_type_embeddings, nf, X)
            Sa = nf * X
        return Fn.MediaInputFn.apply(vx, ax, pv, pa, B), Sv, Sa
