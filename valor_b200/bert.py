"""BERT text / fusion encoder — mirror of `model/bert.py` (BertModel :739-896, BertLayer :423-496,
BertSelfAttention :222-289, BertCrossAttention :292-340, Bert*Output :344-420, BertEmbeddings
:173-218).  Same state-dict keys.

B200-first differences in HOW (not WHAT):
  * the caption passes of one step (tva / tv / ta, pretrain.py:437-471) run as ONE batch of
    3*B sequences: they share every weight and the per-sample media tokens, so the cross-attention
    K/V projection of the 650 media tokens is computed once per layer instead of three times
    (the reference re-projects 650+392+258 tokens); each pass attends to its own
    (kv_row0, kv_len) slice of that projection;
  * q/k/v (self) and k/v (cross) projections are single packed GEMMs; bias, GELU and the
    residual add ride the GEMM epilogues; masks are evaluated inside the attention kernel
    from token ids (no [B,1,T,T] -10000 tensor).
"""
import json
import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .functional import LN, lin_of, fused_lin
from .videoswin import _Linear, _Norm


class BertConfig(object):
    """bert.py:67-152 (subset actually read by the model)."""

    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02):
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as reader:
                for key, value in json.loads(reader.read()).items():
                    self.__dict__[key] = value
        else:
            self.vocab_size = vocab_size_or_config_json_file
            self.hidden_size = hidden_size
            self.num_hidden_layers = num_hidden_layers
            self.num_attention_heads = num_attention_heads
            self.hidden_act = hidden_act
            self.intermediate_size = intermediate_size
            self.hidden_dropout_prob = hidden_dropout_prob
            self.attention_probs_dropout_prob = attention_probs_dropout_prob
            self.max_position_embeddings = max_position_embeddings
            self.type_vocab_size = type_vocab_size
            self.initializer_range = initializer_range
        self.has_cross_attn = getattr(self, "has_cross_attn", False)
        self.cross_attn_type = getattr(self, "cross_attn_type", None)
        self.checkpointing = getattr(self, "checkpointing", False)

    @classmethod
    def from_dict(cls, d):
        c = BertConfig(-1)
        c.__dict__.update(d)
        return c

    @classmethod
    def from_json_file(cls, path):
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.loads(f.read()))


class _Embedding(nn.Module):
    def __init__(self, n, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, dim).normal_(0, 0.02))


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = _Embedding(config.vocab_size, config.hidden_size)
        self.position_embeddings = _Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = _Embedding(config.type_vocab_size, config.hidden_size)
        self.prompt_embedding = _Embedding(1, config.hidden_size)  # unused when use_task_prompt=False
        self.LayerNorm = _Norm(config.hidden_size, eps=1e-12)


class _SelfAttn(nn.Module):
    def __init__(self, config):
        super().__init__()
        H = config.hidden_size
        self.query, self.key, self.value = _Linear(H, H), _Linear(H, H), _Linear(H, H)


class _AttnOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = _Norm(config.hidden_size, eps=1e-12)


class BertAttention(nn.Module):
    def __init__(self, config, attn_type):
        super().__init__()
        self.attn_type = attn_type
        if attn_type == "self":
            self.self = _SelfAttn(config)
        else:
            self.cross = _SelfAttn(config)
        self.output = _AttnOutput(config)


class _Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.intermediate_size)


class _Output(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = _Norm(config.hidden_size, eps=1e-12)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config, "self")
        self.has_cross_attn = config.has_cross_attn
        if self.has_cross_attn:
            assert config.cross_attn_type == "va_concate", "only va_concate is shipped (train_utils.py:631)"
            self.cross_attn = BertAttention(config, "cross")
        self.intermediate = _Intermediate(config)
        self.output = _Output(config)
        self.heads = config.num_attention_heads
        self.hidden_dropout = config.hidden_dropout_prob       # bert.py:348,361,413

    def run(self, h, self_spec, media, cross_spec):
        """h [R*T, H]; media [B*S, H] or None.  (BertLayer.forward, bert.py:440-496)"""
        a = self.attention
        s = a.self
        rng = getattr(self, "_rng", None) if self.training else None
        pd = self.hidden_dropout
        qkv = Fn.linear(h, fused_lin([s.query.weight, s.key.weight, s.value.weight],
                                     [s.query.bias, s.key.bias, s.value.bias]))
        ctx = Fn.SelfAttnFn.apply(qkv, self_spec)
        hin = h
        h = Fn.residual_branch(lambda r: Fn.linear(ctx, lin_of(a.output.dense.weight, a.output.dense.bias), residual=r), hin,
                               rng, pd)                                                # BertSelfOutput, bert.py:351-355
        h = Fn.layer_norm(h, LN(a.output.LayerNorm.weight, a.output.LayerNorm.bias, 1e-12))
        if media is not None:
            c = self.cross_attn
            x = c.cross
            q = Fn.linear(h, lin_of(x.query.weight, x.query.bias))
            kv = Fn.linear(media, fused_lin([x.key.weight, x.value.weight], [x.key.bias, x.value.bias]))
            ctx = Fn.CrossAttnFn.apply(q, kv, cross_spec)
            hin = h
            h = Fn.residual_branch(lambda r: Fn.linear(ctx, lin_of(c.output.dense.weight, c.output.dense.bias), residual=r),
                                   hin, rng, pd)                                      # BertCrossOutput, bert.py:365-371
            h = Fn.layer_norm(h, LN(c.output.LayerNorm.weight, c.output.LayerNorm.bias, 1e-12))
        hin = h
        h = Fn.residual_branch(
            lambda r: Fn.mlp(hin, lin_of(self.intermediate.dense.weight, self.intermediate.dense.bias),
                             lin_of(self.output.dense.weight, self.output.dense.bias), K.ACT_GELU, residual=r), hin, rng, pd)
        return Fn.layer_norm(h, LN(self.output.LayerNorm.weight, self.output.LayerNorm.bias, 1e-12))


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = _Linear(config.hidden_size, config.hidden_size)  # never runs on this path


class BertModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = BertPooler(config)
        self.has_cross_attn = config.has_cross_attn
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    # ------------------------------------------------------------------------------
    def encode(self, tokens, casual, media=None, media_ranges=None, n_media_samples=None):
        """Batched encoder pass.
        tokens [R, T] int64, R = B * n_pass sequences in SAMPLE-major order (row r = b * n_pass + pass): the text rows
        of one sample's passes are adjacent, so its cross-attention is one problem of n_pass*T queries over the
        sample's media tokens; casual: list[bool] per pass or bool;
        media [B*S, H] per-sample media tokens and media_ranges = [(start, len)] per pass, or None.
        Returns hidden states [R*T, H]."""
        R, T = tokens.shape
        cfg = self.config
        H, hd = cfg.num_attention_heads, cfg.hidden_size // cfg.num_attention_heads
        dev = tokens.device
        emb = self.embeddings
        dtype = emb.LayerNorm.weight.lp.dtype
        anchor = self._anchor.requires_grad_(True) if torch.is_grad_enabled() else None
        tables = {"word": emb.word_embeddings.weight, "pos": emb.position_embeddings.weight,
                  "type": emb.token_type_embeddings.weight}
        h = Fn.BertEmbedFn.apply(anchor, tokens.contiguous(), tables, dtype)
        h = Fn.layer_norm(h, LN(emb.LayerNorm.weight, emb.LayerNorm.bias, 1e-12))
        rng = getattr(self, "_rng", None) if self.training else None
        if rng is not None and rng.active and cfg.hidden_dropout_prob > 0:
            h = Fn.DropoutAddFn.apply(h, None, cfg.hidden_dropout_prob, rng)            # bert.py:217
        key_valid = (tokens != 0).to(torch.uint8).contiguous()
        n_pass = len(casual) if isinstance(casual, (list, tuple)) else 1
        cas = list(casual) if isinstance(casual, (list, tuple)) else [casual]
        Bp = R // n_pass
        cache = self.__dict__.setdefault("_spec_cache", {})
        ck = (R, T, tuple(cas), tuple(media_ranges) if media_ranges else None, n_media_samples,
              media.shape[0] if media is not None else 0, str(dev))
        if ck not in cache:
            c_t = torch.tensor([1 if cas[r % n_pass] else 0 for r in range(R)], dtype=torch.uint8, device=dev)
            r_t = l_t = q_t = None
            u0 = ulen = 0
            if media is not None:
                assert n_media_samples == Bp, "one media block per sample"
                S_ = media.shape[0] // n_media_samples
                u0 = min(st for st, ln in media_ranges)                      # union of the passes' key ranges
                ulen = max(st + ln for st, ln in media_ranges) - u0
                r_t = torch.tensor([b * S_ + u0 for b in range(Bp)], dtype=torch.int32, device=dev)
                l_t = torch.full((Bp,), ulen, dtype=torch.int32, device=dev)
                if n_pass > 1 or (u0, ulen) != tuple(media_ranges[0]):
                    q_t = torch.tensor([[media_ranges[i // T][0] - u0, media_ranges[i // T][0] - u0 + media_ranges[i // T][1]]
                                        for i in range(n_pass * T)], dtype=torch.int32, device=dev)
            cache[ck] = (c_t, r_t, l_t, q_t, ulen, media is not None and u0 == 0 and ulen == media.shape[0] // n_media_samples)
        causal, row0_c, lens_c, qrange_c, ulen, kv_full = cache[ck]
        drop = dict(rng=rng, attn_drop=cfg.attention_probs_dropout_prob, dtype_is_fp32=dtype == torch.float32 and tokens.is_cuda)   # bert.py:283,334 (fp32 CUDA = SIMT parity kernels: no dropout there)
        self_spec = dict(P=R, H=H, hd=hd, Nq=T, max_nk=T, scale=1.0 / math.sqrt(hd), key_valid=key_valid, causal=causal, **drop)
        cross_spec = None
        if media is not None:
            cross_spec = dict(P=Bp, H=H, hd=hd, Nq=n_pass * T, max_nk=ulen, scale=1.0 / math.sqrt(hd), kv_row0=row0_c,
                              kv_len=lens_c, q_key_range=qrange_c, kv_exclusive=True, kv_full=kv_full, **drop)
        for layer in self.encoder.layer:
            h = layer.run(h, self_spec, media, cross_spec)
        return h

    def forward(self, tokens, task_prompt=None, video_feat=None, audio_feat=None, casual=False, cache=None,
                use_cache=False, cache_first=False, token_type=None, cache_type="unimlm", use_cross_attn=True,
                full_masker=False):
        """Reference signature (bert.py:750-753), cross-attn branch (:848-896)."""
        assert task_prompt is None and not use_cache and token_type is None and not full_masker, \
            "prompt / KV-cache / full-masker paths are outside the pretraining hot path"
        B, T = tokens.shape
        media, ranges = None, None
        if video_feat is not None or audio_feat is not None:
            parts = [f for f in (video_feat, audio_feat) if f is not None]
            media = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]  # bert.py:450
            ranges = [(0, media.shape[1])]
            media = media.reshape(-1, media.shape[-1]).contiguous()
        h = self.encode(tokens, [casual], media, ranges, B)
        return h.view(B, T, -1)
