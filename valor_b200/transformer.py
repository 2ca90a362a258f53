"""Audio-spectrogram transformer — mirror of `model/transformer.py` (TransformerEncoder :146-170,
TransformerLayer.forward_prenorm :74-85, MultiHeadAttention :106-130, FeedForward :133-142) and
`AudioEmbeddings` (model/modeling.py:738-762).  Same state-dict keys; q/k/v linears run as one
packed GEMM over the adjacent `linears.0/1/2` arena views.
"""
import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .functional import LN, lin_of, fused_lin
from .distributed import mark
from .videoswin import _Linear, _Norm


class MultiHeadAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.linears = nn.ModuleList([_Linear(config.hidden_size, config.hidden_size) for _ in range(4)])
        self.head_num = config.num_attention_heads
        self.hidden_size = config.hidden_size


class FeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.linear1 = _Linear(config.hidden_size, config.intermediate_size)
        self.linear2 = _Linear(config.intermediate_size, config.hidden_size)


class TransformerLayer(nn.Module):
    def __init__(self, config, mode):
        super().__init__()
        assert mode == "prenorm", "only the prenorm encoder is on VALOR's path (modeling.py:609)"
        self.attention = MultiHeadAttention(config)
        self.ff_layer = FeedForward(config)
        self.layernorm1 = _Norm(config.hidden_size, eps=1e-12)
        self.layernorm2 = _Norm(config.hidden_size, eps=1e-12)
        self.mode = mode
        self.hidden_dropout = config.hidden_dropout            # nn.Dropout(config.hidden_dropout), transformer.py:61
        self.attention_dropout = config.attention_dropout      # transformer.py:112 (applied by the attention kernel)

    def run(self, x, n_seq, seq_len):
        att = self.attention
        rng = getattr(self, "_rng", None) if self.training else None
        H = att.head_num
        hd = att.hidden_size // H
        h, xr = Fn.layer_norm_residual(x, LN(self.layernorm1.weight, self.layernorm1.bias, 1e-12))
        qkv = Fn.linear(h, fused_lin([l.weight for l in att.linears[:3]], [l.bias for l in att.linears[:3]]))
        spec = dict(P=n_seq, H=H, hd=hd, Nq=seq_len, max_nk=seq_len, scale=1.0 / math.sqrt(hd), rng=rng,
                    attn_drop=self.attention_dropout, dtype_is_fp32=x.dtype == torch.float32 and x.is_cuda)   # transformer.py:112,128
        o = Fn.SelfAttnFn.apply(qkv, spec)
        x = Fn.residual_branch(lambda r: Fn.linear(o, lin_of(att.linears[3].weight, att.linears[3].bias), residual=r), xr, rng,
                               self.hidden_dropout)                                    # residual + dropout(attn), :78
        h, xr = Fn.layer_norm_residual(x, LN(self.layernorm2.weight, self.layernorm2.bias, 1e-12))
        ff = self.ff_layer
        return Fn.residual_branch(
            lambda r: Fn.mlp(h, lin_of(ff.linear1.weight, ff.linear1.bias), lin_of(ff.linear2.weight, ff.linear2.bias),
                             K.ACT_GELU, residual=r), xr, rng, self.hidden_dropout)   # residual + dropout(ff), :83


class TransformerEncoder(nn.Module):
    def __init__(self, config, mode="prenorm"):
        super().__init__()
        self.mode = mode
        self.layer = nn.ModuleList([TransformerLayer(config, mode) for _ in range(config.num_hidden_layers)])
        self.last_layernorm = _Norm(config.hidden_size, eps=1e-12)

    def run(self, x, n_seq, seq_len):
        x = mark(x, "ast")                                      # gradient all-reduce bucket boundary (distributed.py)
        for layer in self.layer:
            x = layer.run(x, n_seq, seq_len)
        return Fn.layer_norm(x, LN(self.last_layernorm.weight, self.last_layernorm.bias, 1e-12))

    def forward(self, input_, attention_mask=None, cross_hidden_states=None, use_cache=False, cache=None,
                cache_first=False, cache_type=None):
        """Reference signature (transformer.py:156-170): [N, L, H] -> ([N, L, H], cache)."""
        assert attention_mask is None, "the AST path passes no mask (modeling.py:474)"
        n, l, h = input_.shape
        return self.run(input_.reshape(n * l, h), n, l).view(n, l, h), cache


class AudioEmbeddings(nn.Module):
    def __init__(self, model_cfg_audio, config):
        super().__init__()
        self.patch_size = config.audio_patch_size
        self.token_length_per_frame = (config.audio_melbins // self.patch_size) * (config.audio_target_length // self.patch_size)
        H = model_cfg_audio.hidden_size
        self.first_conv = nn.Module()
        self.first_conv.weight = nn.Parameter(torch.empty(H, 1, self.patch_size, self.patch_size).normal_(0, 0.02))
        self.first_conv.bias = nn.Parameter(torch.zeros(H))
        self.position_embeddings = nn.Module()
        self.position_embeddings.weight = nn.Parameter(torch.empty(self.token_length_per_frame + 1, H).normal_(0, 0.02))
        self.cls_token = nn.Parameter(0.02 * torch.randn(1, 1, H))
        self.hidden_dropout = model_cfg_audio.hidden_dropout   # modeling.py:748,761
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    def forward(self, audio_spectrograms):
        """Reference signature (modeling.py:750-762): [N, mel, frames] -> [N, P+1, H]."""
        N = audio_spectrograms.shape[0]
        return self.run(audio_spectrograms, self.cls_token.lp.dtype).view(N, self.token_length_per_frame + 1, -1)

    def run(self, spec, dtype):
        """spec [N, mel, frames] -> [N*(P+1), H]   (conv-as-GEMM, cls + position, modeling.py:750-762)."""
        N = spec.shape[0]
        H = self.cls_token.shape[-1]
        cols = K.audio_im2col(spec.contiguous(), self.patch_size, dtype)
        w = self.first_conv.weight
        lin = Fn.Lin(w.lp.view(H, -1), w.main_grad.view(H, -1), self.first_conv.bias.data, self.first_conv.bias.main_grad)
        anchor = self._anchor.requires_grad_(True) if torch.is_grad_enabled() else None
        tok = Fn.linear(cols, lin, anchor=anchor)
        x = Fn.AstAssembleFn.apply(tok, self.cls_token, self.position_embeddings.weight, N, self.token_length_per_frame)
        rng = getattr(self, "_rng", None) if self.training else None
        if rng is not None and rng.active and self.hidden_dropout > 0:
            x = Fn.DropoutAddFn.apply(x, None, self.hidden_dropout, rng)              # modeling.py:761
        return x
