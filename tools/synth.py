"""Seeded synthetic weights / batches with the reference's key names (benchmark inputs, smoke test, test fixtures;
no network: there are no real checkpoints or datasets here).  Nothing in the model path imports this module.

Every tensor is drawn from its own CPU generator seeded by crc32(key) ^ seed, so the same
state dict is reproduced bit-for-bit in the build container (where the live reference
consumes it, tests/golden/make_golden.py) and on the GPU box (where only the oracle and
the CUDA path exist).  Shapes follow the reference's `state_dict()` [SURVEY.md §8b].

"Trained-like" statistics: LN gamma ~ U(0.5,1.5), beta ~ N(0,0.1),
biases ~ N(0,0.02) — the reference's own init (all biases 0, gamma 1) would hide bias /
affine bugs.
"""
import dataclasses
import random
import zlib

import numpy as np
import torch


@dataclasses.dataclass(frozen=True)
class Geometry:
    """VALOR-base as pretrained (scripts/pretrain.sh:3-8): VideoSwin-B + AST + BERT-base."""
    swin_embed: int = 128
    swin_depths: tuple = (2, 2, 18, 2)
    swin_heads: tuple = (4, 8, 16, 32)
    swin_window: tuple = (8, 7, 7)
    ast_layers: int = 12
    bert_layers: int = 12
    hidden: int = 768
    heads: int = 12
    ffn: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    contra_dim: int = 512
    resolution: int = 224
    audio_melbins: int = 64
    audio_frames: int = 512
    audio_patch: int = 16

    @property
    def video_dim(self):
        return self.swin_embed * 8

    @property
    def audio_tokens(self):
        return (self.audio_melbins // self.audio_patch) * (self.audio_frames // self.audio_patch) + 1

    def bert_config_json(self):
        return dict(vocab_size=self.vocab, hidden_size=self.hidden, num_hidden_layers=self.bert_layers,
                    num_attention_heads=self.heads, intermediate_size=self.ffn, hidden_act="gelu",
                    hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                    max_position_embeddings=self.max_pos, type_vocab_size=2, initializer_range=0.02)


BASE = Geometry()
# reduced-depth geometry for fast parity runs; widths are pinned by the reference
# (modeling.py:270-278,585-587,622 hard-code 768/1024), only depths may shrink.
TINY = Geometry(swin_depths=(2, 2, 2, 2), ast_layers=2, bert_layers=2)


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(key, seed, shape, std):
    return torch.randn(shape, generator=_gen(key, seed), dtype=torch.float32) * std


def _uniform(key, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32) * (hi - lo) + lo


def relative_position_index(window):
    """Restates WindowAttention3D.__init__ (model/videoswin.py:113-127)."""
    wd, wh, ww = window
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    flat = coords.flatten(1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


def make_state_dict(geom: Geometry = BASE, seed: int = 0, include_buffers: bool = True):
    """Reference-keyed state dict (model/pretrain.py VALOR.state_dict()), fp32 CPU."""
    sd = {}

    def lin(prefix, out_f, in_f, bias=True, std=0.02):
        sd[prefix + ".weight"] = _normal(prefix + ".weight", seed, (out_f, in_f), std)
        if bias:
            sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (out_f,), 0.02)

    def ln(prefix, dim, wname="weight", bname="bias"):
        sd[f"{prefix}.{wname}"] = _uniform(f"{prefix}.{wname}", seed, (dim,), 0.5, 1.5)
        sd[f"{prefix}.{bname}"] = _normal(f"{prefix}.{bname}", seed, (dim,), 0.1)

    # ---- VideoSwin (model/videoswin.py:378-439)
    E = geom.swin_embed
    p = "video_encoder."
    sd[p + "patch_embed.proj.weight"] = _normal(p + "patch_embed.proj.weight", seed, (E, 3, 2, 4, 4), 0.05)
    sd[p + "patch_embed.proj.bias"] = _normal(p + "patch_embed.proj.bias", seed, (E,), 0.02)
    ln(p + "patch_embed.norm", E)
    wd, wh, ww = geom.swin_window
    n_rel = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1)
    rpi = relative_position_index(geom.swin_window)
    for s, depth in enumerate(geom.swin_depths):
        C = E * 2 ** s
        h = geom.swin_heads[s]
        for b in range(depth):
            q = f"{p}layers.{s}.blocks.{b}."
            ln(q + "norm1", C)
            sd[q + "attn.relative_position_bias_table"] = _normal(
                q + "attn.relative_position_bias_table", seed, (n_rel, h), 0.5)
            if include_buffers:
                sd[q + "attn.relative_position_index"] = rpi.clone()
            lin(q + "attn.qkv", 3 * C, C)
            lin(q + "attn.proj", C, C)
            ln(q + "norm2", C)
            lin(q + "mlp.fc1", 4 * C, C)
            lin(q + "mlp.fc2", C, 4 * C)
        if s < len(geom.swin_depths) - 1:
            q = f"{p}layers.{s}.downsample."
            lin(q + "reduction", 2 * C, 4 * C, bias=False)
            ln(q + "norm", 4 * C)
    ln(p + "norm", E * 8)

    # ---- AST (model/modeling.py:738-762, model/transformer.py)
    H = geom.hidden
    sd["audio_embeddings.cls_token"] = _normal("audio_embeddings.cls_token", seed, (1, 1, H), 0.02)
    sd["audio_embeddings.first_conv.weight"] = _normal(
        "audio_embeddings.first_conv.weight", seed, (H, 1, geom.audio_patch, geom.audio_patch), 0.05)
    sd["audio_embeddings.first_conv.bias"] = _normal("audio_embeddings.first_conv.bias", seed, (H,), 0.02)
    sd["audio_embeddings.position_embeddings.weight"] = _normal(
        "audio_embeddings.position_embeddings.weight", seed, (geom.audio_tokens, H), 0.02)
    for i in range(geom.ast_layers):
        q = f"audio_encoder.layer.{i}."
        for j in range(4):
            lin(q + f"attention.linears.{j}", H, H)
        lin(q + "ff_layer.linear1", geom.ffn, H)
        lin(q + "ff_layer.linear2", H, geom.ffn)
        ln(q + "layernorm1", H)
        ln(q + "layernorm2", H)
    ln("audio_encoder.last_layernorm", H)

    # ---- BERT text/fusion (model/bert.py:739-748) — txt_encoder IS multimodal_encoder
    q = "multimodal_encoder."
    sd[q + "embeddings.word_embeddings.weight"] = _normal(q + "embeddings.word_embeddings.weight", seed,
                                                          (geom.vocab, H), 0.02)
    sd[q + "embeddings.position_embeddings.weight"] = _normal(q + "embeddings.position_embeddings.weight",
                                                              seed, (geom.max_pos, H), 0.02)
    sd[q + "embeddings.token_type_embeddings.weight"] = _normal(q + "embeddings.token_type_embeddings.weight",
                                                                seed, (2, H), 0.02)
    sd[q + "embeddings.prompt_embedding.weight"] = _normal(q + "embeddings.prompt_embedding.weight", seed,
                                                           (1, H), 0.02)
    ln(q + "embeddings.LayerNorm", H)
    for i in range(geom.bert_layers):
        r = f"{q}encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            lin(r + "attention.self." + nm, H, H)
        lin(r + "attention.output.dense", H, H)
        ln(r + "attention.output.LayerNorm", H)
        for nm in ("query", "key", "value"):
            lin(r + "cross_attn.cross." + nm, H, H)
        lin(r + "cross_attn.output.dense", H, H)
        ln(r + "cross_attn.output.LayerNorm", H)
        lin(r + "intermediate.dense", geom.ffn, H)
        lin(r + "output.dense", H, geom.ffn)
        ln(r + "output.LayerNorm", H)
    lin(q + "pooler.dense", H, H)
    for k in [k for k in sd if k.startswith("multimodal_encoder.")]:
        sd["txt_encoder." + k[len("multimodal_encoder."):]] = sd[k]

    # ---- MLM head (model/modeling.py:232-254), decoder tied to word embeddings (:241)
    lin("cls.dense", H, H)
    ln("cls.layernorm", H)
    sd["cls.decoder.weight"] = sd["multimodal_encoder.embeddings.word_embeddings.weight"]
    sd["cls.decoder.bias"] = _normal("cls.decoder.bias", seed, (geom.vocab,), 0.02)

    # ---- adapters, embeddings, heads (model/modeling.py:341-351, model/pretrain.py:89-119)
    for nm in ("video_type_embeddings", "audio_type_embeddings"):
        sd[nm] = _normal(nm, seed, (1, 1, H), 0.02)
    for nm in ("video_frame_embedding", "audio_frame_embedding"):
        sd[nm] = _normal(nm, seed, (1, 32, H), 0.02)
    if geom.video_dim != H:
        lin("hidden_trans_video_multimodal.0", H, geom.video_dim)
        ln("hidden_trans_video_multimodal.1", H)
    lin("contra_head_t.linear", geom.contra_dim, H, bias=False, std=0.05)
    lin("contra_head_v.linear", geom.contra_dim, geom.video_dim, bias=False, std=0.05)
    lin("contra_head_a.linear", geom.contra_dim, H, bias=False, std=0.05)
    for nm in ("text", "video", "audio"):
        lin(f"{nm}_fine_weight.0", geom.contra_dim, geom.contra_dim, std=0.05)
        lin(f"{nm}_fine_weight.2", 1, geom.contra_dim, std=0.05)
    sd["contra_temp"] = torch.tensor(0.07)
    return sd


def ast_checkpoint_from_state(sd, geom):
    """Shape-correct stand-in for audioset_10_10_0.4593.pth (model/modeling.py:514-548); the
    values are overwritten by load_state_dict afterwards."""
    H = geom.hidden
    ck = {"module.v.cls_token": sd["audio_embeddings.cls_token"].clone(),
          "module.v.dist_token": torch.zeros(1, 1, H),
          "module.v.patch_embed.proj.weight": sd["audio_embeddings.first_conv.weight"].clone(),
          "module.v.patch_embed.proj.bias": sd["audio_embeddings.first_conv.bias"].clone(),
          "module.v.pos_embed": torch.zeros(1, 1214, H),
          "module.v.norm.weight": torch.ones(H), "module.v.norm.bias": torch.zeros(H)}
    for i in range(12):  # the loader indexes 12 blocks unconditionally (modeling.py:530)
        p = f"module.v.blocks.{i}."
        ck[p + "attn.qkv.weight"] = torch.zeros(3 * H, H)
        ck[p + "attn.qkv.bias"] = torch.zeros(3 * H)
        ck[p + "attn.proj.weight"] = torch.zeros(H, H)
        ck[p + "attn.proj.bias"] = torch.zeros(H)
        ck[p + "mlp.fc1.weight"] = torch.zeros(geom.ffn, H)
        ck[p + "mlp.fc1.bias"] = torch.zeros(geom.ffn)
        ck[p + "mlp.fc2.weight"] = torch.zeros(H, geom.ffn)
        ck[p + "mlp.fc2.bias"] = torch.zeros(H)
        for n in ("norm1", "norm2"):
            ck[p + n + ".weight"] = torch.ones(H)
            ck[p + n + ".bias"] = torch.zeros(H)
    return ck


def bert_checkpoint_from_state(sd, geom):
    """Stand-in for bert-base-uncased.bin: the six cls.predictions.* keys are indexed
    unconditionally (model/modeling.py:646-652)."""
    H = geom.hidden
    return {"cls.predictions.transform.dense.weight": torch.zeros(H, H),
            "cls.predictions.transform.dense.bias": torch.zeros(H),
            "cls.predictions.transform.LayerNorm.weight": torch.ones(H),
            "cls.predictions.transform.LayerNorm.bias": torch.zeros(H),
            "cls.predictions.decoder.weight": torch.zeros(geom.vocab, H),
            "cls.predictions.bias": torch.zeros(geom.vocab)}


def make_batch(B, F, A, T, geom: Geometry = BASE, seed: int = 123, min_len=8):
    """Synthetic batch with the reference's schema (data/data.py:423-428)."""
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, F, 3, geom.resolution, geom.resolution, generator=g)
    audio = torch.randn(B, A, geom.audio_melbins, geom.audio_frames, generator=g)
    tokens = torch.zeros(B, T, dtype=torch.long)
    max_words = T - 2
    lens = torch.randint(min(min_len, max_words), max_words + 1, (B,), generator=g)
    for i in range(B):
        L = int(lens[i])
        tokens[i, 0] = 101
        tokens[i, 1:1 + L] = torch.randint(1000, 30000, (L,), generator=g)
        tokens[i, 1 + L] = 102
    return {"video_pixels": video, "audio_spectrograms": audio, "txt_tokens": {"bert_tokens": tokens},
            "ids": [f"s{i}" for i in range(B)]}


def token_masker(tokens, mask_prob, mask_token=103, range_start=106, range_end=30522, seed=1234):
    """Restates TokenMasker.perform_mask (model/modeling.py:134-174) with an explicit seed:
    position 0 never masked, any token != 0 (incl. [SEP]) maskable, >=1 mask per row,
    80% [MASK] / 10% random id in [range_start, range_end) / 10% keep; labels -1 elsewhere."""
    rng = random.Random(seed)
    tok = np.array(tokens.cpu().numpy())
    ind = np.zeros(tok.shape, dtype=np.int64)
    for i in range(tok.shape[0]):
        while not ind[i].any():
            for j in range(1, tok.shape[1]):
                if tok[i][j] != 0 and rng.random() < mask_prob:
                    ind[i][j] = 1
    labels = -np.ones(tok.shape, dtype=np.int64)
    for i in range(tok.shape[0]):
        for j in range(tok.shape[1]):
            if ind[i][j] == 1:
                src = tok[i][j]
                prob = rng.random()
                if prob < 0.8:
                    tok[i][j] = mask_token
                elif prob < 0.9:
                    tok[i][j] = rng.choice(range(range_start, range_end))
                labels[i][j] = src
    return torch.from_numpy(tok).long(), torch.from_numpy(labels).long()
