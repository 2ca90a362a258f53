"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of VALOR's pretraining hot path.

This is the ORACLE the CUDA path is checked against (tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference).  It is NOT part of the product path and is
never imported by `valor_b200/`.  It restates, function by function, the algorithm of the
reference (TXH-mercury/VALOR, Python/PyTorch) in plain torch CPU ops, taking a state dict
with the reference's own key names.  Each function cites the reference file:line it follows.

Pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
`tests/golden/make_golden.py` (fixtures committed under tests/golden/); see
tests/test_oracle_golden.py.

Parity mode: Dropout / DropPath are disabled on both sides (they are host-RNG driven in the
reference); TokenMasker draws are hoisted (`oracle.synth.token_masker`) and fed to both.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------


def gelu_erf(x):
    """model/transformer.py:32-38, model/bert.py:52-57 (erf-exact GELU); nn.GELU in Swin."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    """apex FusedLayerNorm CPU path = F.layer_norm (apex fused_layer_norm.py:154-156)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear(x, sd, prefix, bias=True):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"] if bias else None)


# --------------------------------------------------------------------------------------
# VideoSwin  (model/videoswin.py)
# --------------------------------------------------------------------------------------


def window_partition(x, ws):
    """videoswin.py:75-79"""
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], C)


def window_reverse(windows, ws, B, D, H, W):
    """videoswin.py:81-84"""
    x = windows.view(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


def get_window_size(x_size, window_size, shift_size=None):
    """videoswin.py:86-99"""
    use_w = list(window_size)
    use_s = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_w[i] = x_size[i]
            if use_s is not None:
                use_s[i] = 0
    if shift_size is None:
        return tuple(use_w)
    return tuple(use_w), tuple(use_s)


def compute_mask(D, H, W, ws, ss):
    """videoswin.py:272-285 — additive mask is -100 (not -inf)."""
    img = torch.zeros((1, D, H, W, 1))
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    mw = window_partition(img, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


# bench.py's "torch eager on the same B200" leg (the competitor SURVEY F10 names: cuBLAS + SDPA) flips this on; the
# parity checker always runs the explicit matmul / softmax / matmul form of the reference.
USE_SDPA = False


def window_attention(x, sd, p, heads, rel_index, mask):
    """WindowAttention3D.forward, videoswin.py:137-163 (q scaled BEFORE q@k^T, :143)."""
    B_, N, C = x.shape
    qkv = linear(x, sd, p + "qkv").reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (C // heads) ** -0.5
    table = sd[p + "relative_position_bias_table"]
    bias = table[rel_index[:N, :N].reshape(-1)].reshape(N, N, -1).permute(2, 0, 1).contiguous()
    if USE_SDPA:
        am = bias.unsqueeze(0).to(q.dtype)
        if mask is not None:
            nW = mask.shape[0]
            am = (am.unsqueeze(0) + mask.unsqueeze(1).unsqueeze(0).to(q.dtype)).expand(B_ // nW, nW, heads, N, N).reshape(-1, heads, N, N)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=am, scale=1.0)
        return linear(o.transpose(1, 2).reshape(B_, N, C), sd, p + "proj")
    attn = q @ k.transpose(-2, -1) + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return linear(x, sd, p + "proj")


def swin_block(x, sd, p, heads, window, shift, rel_index, mask_matrix):
    """SwinTransformerBlock3D.forward, videoswin.py:191-245 (DropPath off in parity mode)."""
    B, D, H, W, C = x.shape
    ws, ss = get_window_size((D, H, W), window, shift)
    shortcut = x
    x = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    pad_d1 = (ws[0] - D % ws[0]) % ws[0]
    pad_b = (ws[1] - H % ws[1]) % ws[1]
    pad_r = (ws[2] - W % ws[2]) % ws[2]
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b, 0, pad_d1))
    _, Dp, Hp, Wp, _ = x.shape
    if any(i > 0 for i in ss):
        shifted = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        mask = mask_matrix
    else:
        shifted, mask = x, None
    xw = window_partition(shifted, ws)
    aw = window_attention(xw, sd, p + "attn.", heads, rel_index, mask)
    aw = aw.view(-1, *(ws + (C,)))
    shifted = window_reverse(aw, ws, B, Dp, Hp, Wp)
    if any(i > 0 for i in ss):
        x = torch.roll(shifted, shifts=(ss[0], ss[1], ss[2]), dims=(1, 2, 3))
    else:
        x = shifted
    if pad_d1 > 0 or pad_r > 0 or pad_b > 0:
        x = x[:, :D, :H, :W, :].contiguous()
    x = shortcut + x
    y = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    y = linear(F.gelu(linear(y, sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    return x + y


def patch_merging(x, sd, p):
    """PatchMerging.forward, videoswin.py:254-270."""
    B, D, H, W, C = x.shape
    if (H % 2 == 1) or (W % 2 == 1):
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, :, 0::2, 0::2, :], x[:, :, 1::2, 0::2, :], x[:, :, 0::2, 1::2, :],
                   x[:, :, 1::2, 1::2, :]], -1)
    x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def swin_forward(video, sd, geom, p="video_encoder."):
    """SwinTransformer3D.forward, videoswin.py:441-458; PatchEmbed3D :361-376;
    BasicLayer.forward :329-345.  video: [B,3,D,H,W] -> [B, 8E, D, H/32, W/32]."""
    from tools.synth import relative_position_index
    x = F.pad(video, (0, 0, 0, 0, 0, 1))
    x = F.conv3d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=(1, 4, 4))
    B, E, D, Wh, Ww = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = layer_norm(x, sd[p + "patch_embed.norm.weight"], sd[p + "patch_embed.norm.bias"], 1e-5)
    x = x.transpose(1, 2).view(-1, E, D, Wh, Ww)
    rel_index = relative_position_index(geom.swin_window).to(video.device)
    window = geom.swin_window
    shift_full = tuple(i // 2 for i in window)
    for s, depth in enumerate(geom.swin_depths):
        B, C, D, H, W = x.shape
        ws, ss = get_window_size((D, H, W), window, shift_full)
        x = x.permute(0, 2, 3, 4, 1)
        Dp = int(math.ceil(D / ws[0])) * ws[0]
        Hp = int(math.ceil(H / ws[1])) * ws[1]
        Wp = int(math.ceil(W / ws[2])) * ws[2]
        mask = compute_mask(Dp, Hp, Wp, ws, ss).to(video.device)
        for b in range(depth):
            shift = (0, 0, 0) if b % 2 == 0 else shift_full
            x = swin_block(x, sd, f"{p}layers.{s}.blocks.{b}.", geom.swin_heads[s], window, shift, rel_index, mask)
        x = x.view(B, D, H, W, -1)
        if s < len(geom.swin_depths) - 1:
            x = patch_merging(x, sd, f"{p}layers.{s}.downsample.")
        x = x.permute(0, 4, 1, 2, 3)
    x = x.permute(0, 2, 3, 4, 1)
    x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return x.permute(0, 4, 1, 2, 3)


def forward_video_encoder(video_pixels, sd, geom):
    """VALORModel.forward_video_encoder (videoswin branch), model/modeling.py:449-455."""
    out = swin_forward(video_pixels.transpose(1, 2), sd, geom)
    out = out.permute(0, 2, 3, 4, 1)
    return out.reshape(out.shape[0], out.shape[1], -1, out.shape[-1])  # [B,F,49,1024]


# --------------------------------------------------------------------------------------
# AST  (model/modeling.py:738-762, model/transformer.py)
# --------------------------------------------------------------------------------------


def audio_embeddings(spec, sd, geom):
    """AudioEmbeddings.forward, modeling.py:750-762.  spec: [N,mel,frames]."""
    x = F.conv2d(spec.unsqueeze(1), sd["audio_embeddings.first_conv.weight"],
                 sd["audio_embeddings.first_conv.bias"], stride=geom.audio_patch)
    b, c = x.shape[:2]
    x = x.permute(0, 2, 3, 1).reshape(b, -1, c)
    x = torch.cat((sd["audio_embeddings.cls_token"].expand(b, -1, -1), x), dim=1)
    return x + sd["audio_embeddings.position_embeddings.weight"][: x.shape[1]].unsqueeze(0)


def ast_mha(x, sd, p, heads):
    """MultiHeadAttention.forward, transformer.py:115-130 (no mask on the AST path)."""
    B, N, H = x.shape
    q, k, v = [linear(x, sd, f"{p}linears.{j}").view(B, -1, heads, H // heads).transpose(1, 2) for j in range(3)]
    if USE_SDPA:
        out = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).contiguous().view(B, -1, H)
        return linear(out, sd, p + "linears.3")
    att = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    att = att.softmax(-1)
    out = torch.matmul(att, v).transpose(1, 2).contiguous().view(B, -1, H)
    return linear(out, sd, p + "linears.3")


def ast_forward(x, sd, geom):
    """TransformerEncoder.forward / TransformerLayer.forward_prenorm, transformer.py:74-85,156-170."""
    for i in range(geom.ast_layers):
        p = f"audio_encoder.layer.{i}."
        h = layer_norm(x, sd[p + "layernorm1.weight"], sd[p + "layernorm1.bias"], 1e-12)
        x = x + ast_mha(h, sd, p + "attention.", geom.heads)
        h = layer_norm(x, sd[p + "layernorm2.weight"], sd[p + "layernorm2.bias"], 1e-12)
        x = x + linear(gelu_erf(linear(h, sd, p + "ff_layer.linear1")), sd, p + "ff_layer.linear2")
    return layer_norm(x, sd["audio_encoder.last_layernorm.weight"], sd["audio_encoder.last_layernorm.bias"], 1e-12)


def forward_audio_encoder(spec, sd, geom):
    """VALORModel.forward_audio_encoder, modeling.py:468-480."""
    b, n = spec.shape[:2]
    x = audio_embeddings(spec.reshape(-1, *spec.shape[2:]), sd, geom)
    x = ast_forward(x, sd, geom)
    return x.reshape(b, n, -1, x.shape[-1])  # [B,A,129,768]


# --------------------------------------------------------------------------------------
# BERT text / fusion encoder  (model/bert.py)
# --------------------------------------------------------------------------------------


def bert_embeddings(tokens, sd, p):
    """BertEmbeddings.forward (token_type None), bert.py:190-218."""
    T = tokens.shape[1]
    e = (sd[p + "word_embeddings.weight"][tokens] + sd[p + "position_embeddings.weight"][:T].unsqueeze(0)
         + sd[p + "token_type_embeddings.weight"][0])
    return layer_norm(e, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)


def bert_attention_core(q, k, v, heads, mask):
    """BertSelfAttention / BertCrossAttention core, bert.py:249-289,319-340 (scale AFTER q@k^T)."""
    B, Tq, H = q.shape
    hd = H // heads

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, hd).permute(0, 2, 1, 3)

    if USE_SDPA:
        am = None if mask is None else mask.to(q.dtype).expand(B, heads, Tq, k.shape[1])
        ctx = F.scaled_dot_product_attention(split(q), split(k), split(v), attn_mask=am)
        return ctx.permute(0, 2, 1, 3).contiguous().view(B, Tq, H)
    s = torch.matmul(split(q), split(k).transpose(-1, -2)) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    ctx = torch.matmul(s.softmax(-1), split(v))
    return ctx.permute(0, 2, 1, 3).contiguous().view(B, Tq, H)


def bert_self_mask(tokens, casual):
    """BertModel.forward cross-branch mask build, bert.py:854-885 (no task prompt)."""
    am = (tokens != 0).long()
    T = am.shape[1]
    am = am.unsqueeze(1).expand(-1, T, -1).clone()
    if casual:
        am[:, :T, :T] = torch.tril(am[:, :T, :T])
    return (1.0 - am.unsqueeze(1).float()) * -10000.0


def bert_layer(h, mask, media, sd, p, heads):
    """BertLayer.forward, bert.py:440-496 (va_concate; cross sub-layer skipped w/o media :456)."""
    a = p + "attention."
    ctx = bert_attention_core(linear(h, sd, a + "self.query"), linear(h, sd, a + "self.key"),
                              linear(h, sd, a + "self.value"), heads, mask)
    h = layer_norm(linear(ctx, sd, a + "output.dense") + h, sd[a + "output.LayerNorm.weight"],
                   sd[a + "output.LayerNorm.bias"], 1e-12)
    if media is not None:
        c = p + "cross_attn."
        ctx = bert_attention_core(linear(h, sd, c + "cross.query"), linear(media, sd, c + "cross.key"),
                                  linear(media, sd, c + "cross.value"), heads, None)
        h = layer_norm(linear(ctx, sd, c + "output.dense") + h, sd[c + "output.LayerNorm.weight"],
                       sd[c + "output.LayerNorm.bias"], 1e-12)
    inter = gelu_erf(linear(h, sd, p + "intermediate.dense"))
    return layer_norm(linear(inter, sd, p + "output.dense") + h, sd[p + "output.LayerNorm.weight"],
                      sd[p + "output.LayerNorm.bias"], 1e-12)


def bert_forward(tokens, sd, geom, video_feat=None, audio_feat=None, casual=False, p="multimodal_encoder."):
    """BertModel.forward cross-attn branch, bert.py:848-896."""
    h = bert_embeddings(tokens, sd, p + "embeddings.")
    mask = bert_self_mask(tokens, casual)
    if video_feat is not None and audio_feat is not None:
        media = torch.cat((video_feat, audio_feat), dim=1)  # bert.py:450
    else:
        media = video_feat if video_feat is not None else audio_feat
    for i in range(geom.bert_layers):
        h = bert_layer(h, mask, media, sd, f"{p}encoder.layer.{i}.", geom.heads)
    return h


def mlm_head(x, sd):
    """BERTPredictionHead.forward, modeling.py:245-254 (decoder tied to word embeddings :241)."""
    x = gelu_erf(linear(x, sd, "cls.dense"))
    x = layer_norm(x, sd["cls.layernorm.weight"], sd["cls.layernorm.bias"], 1e-12)
    return F.linear(x, sd["cls.decoder.weight"], sd["cls.decoder.bias"])


# --------------------------------------------------------------------------------------
# heads / losses  (model/modeling.py, model/pretrain.py)
# --------------------------------------------------------------------------------------


def multimodal_input_video(video_output, sd):
    """get_multimodal_forward_input_video, modeling.py:485-493."""
    b, n, x, c = video_output.shape
    if "hidden_trans_video_multimodal.0.weight" in sd:
        video_output = linear(video_output, sd, "hidden_trans_video_multimodal.0")
        video_output = layer_norm(video_output, sd["hidden_trans_video_multimodal.1.weight"],
                                  sd["hidden_trans_video_multimodal.1.bias"], 1e-12)
    video_output = video_output + sd["video_frame_embedding"][:, :n, :].unsqueeze(-2)
    video_output = video_output.reshape(b, -1, video_output.shape[-1])
    return video_output + sd["video_type_embeddings"]


def multimodal_input_audio(audio_output, sd):
    """get_multimodal_forward_input_audio, modeling.py:495-502."""
    b, n, x, c = audio_output.shape
    audio_output = audio_output + sd["audio_frame_embedding"][:, :n, :].unsqueeze(-2)
    audio_output = audio_output.reshape(b, -1, c)
    return audio_output + sd["audio_type_embeddings"]


def fine_weight(feat, sd, name):
    """fine_weight_mapper MLP, pretrain.py:104-112."""
    return linear(F.relu(linear(feat, sd, f"{name}_fine_weight.0")), sd, f"{name}_fine_weight.2").squeeze(2)


def compute_fine_matrix(featA, featB, maskA, maskB, weightA, weightB):
    """compute_fine_matrix_slice, pretrain.py:191-211.  Masks multiply the logits BEFORE the
    max (:201-205) so a padded slot contributes 0, not -inf."""
    weightA = weightA.masked_fill((1 - maskA).bool(), float("-inf")).softmax(-1)
    weightB = weightB.masked_fill((1 - maskB).bool(), float("-inf")).softmax(-1)
    logits = torch.einsum("atd,bvd->abtv", featA, featB)
    logits = torch.einsum("abtv,at->abtv", logits, maskA.to(logits.dtype))
    logits = torch.einsum("abtv,bv->abtv", logits, maskB.to(logits.dtype))
    a2b = logits.max(dim=-1)[0]
    b2a = logits.max(dim=-2)[0]
    a2b = torch.einsum("abt,at->ab", a2b, weightA)
    b2a = torch.einsum("abv,bv->ab", b2a, weightB)
    return (a2b + b2a) / 2.0


def contrastive_loss(score, temp):
    """VALORModel.contrastive_loss, modeling.py:418-433."""
    s = score / temp
    l1 = (-F.log_softmax(s, dim=1)).diag()
    l2 = (-F.log_softmax(s, dim=0)).diag()
    return torch.cat((l1, l2), dim=0).mean()


def gather_with_grads(x, world=None):
    """ddp_allgather_with_grads, utils/distributed.py:38-72 — world_size-1 restatement is the
    identity; `world` = list of per-rank tensors emulates N ranks on one process (rank 0's
    rows keep their graph, the rest are constants: backward slices the local rows only)."""
    if world is None:
        return x
    return torch.cat([x] + [w.detach() for w in world], dim=0)


def forward_pt(batch, sd, geom, txt_input, txt_labels, task="pt_contra%tva%tv%ta_caption%tva%tv%ta",
               contra_loss_ratio=1.5, return_aux=False):
    """VALOR.forward_pt, model/pretrain.py:214-541, for contra_type='fine', caption_type='unimlm',
    use_task_prompt=False, late_fusion=False, compute_loss=True, world_size 1.
    `txt_input/txt_labels` = hoisted TokenMasker draw (pretrain.py:428)."""
    contra_task, caption_task = [], []
    for t in task.split("_"):
        if "caption" in t:
            caption_task = t.split("%")[1:]
        elif "contra" in t:
            contra_task = t.split("%")[1:]
    tokens = batch["txt_tokens"]["bert_tokens"]
    aux = {}
    video_output = forward_video_encoder(batch["video_pixels"], sd, geom)
    audio_output = forward_audio_encoder(batch["audio_spectrograms"], sd, geom)
    aux["video_output"], aux["audio_output"] = video_output, audio_output
    losses = {}
    if contra_task:
        txt_output = bert_forward(tokens, sd, geom, casual=False, p="txt_encoder.")  # modeling.py:437-446
        aux["txt_output"] = txt_output
        feat_t = F.normalize(F.linear(txt_output, sd["contra_head_t.linear.weight"]), dim=-1)
        feat_v = F.normalize(F.linear(video_output.mean(dim=2), sd["contra_head_v.linear.weight"]), dim=-1)
        feat_a = F.normalize(F.linear(audio_output[:, :, 0], sd["contra_head_a.linear.weight"]), dim=-1)
        aux["feat_t"], aux["feat_v"], aux["feat_a"] = feat_t, feat_v, feat_a
        maskA = (tokens != 0).long()
        wt = fine_weight(feat_t, sd, "text")
        wv = fine_weight(feat_v, sd, "video")
        wa = fine_weight(feat_a, sd, "audio")
        lo = []
        ones = lambda f: torch.ones(*f.shape[:2], dtype=torch.long, device=f.device)
        temp = sd["contra_temp"]
        # order of accumulation follows pretrain.py:397: (tva, tv, ta)
        if "tva" in contra_task:
            feat_va = torch.cat((feat_v, feat_a), dim=1)
            sc = compute_fine_matrix(feat_t, feat_va, maskA, ones(feat_va), wt.clone(), torch.cat((wv, wa), dim=1))
            aux["score_tva"] = sc
            lo.append(contrastive_loss(sc, temp))
        if "tv" in contra_task:
            sc = compute_fine_matrix(feat_t, feat_v, maskA, ones(feat_v), wt.clone(), wv.clone())
            lo.append(contrastive_loss(sc, temp))
        if "ta" in contra_task:
            sc = compute_fine_matrix(feat_t, feat_a, maskA, ones(feat_a), wt.clone(), wa.clone())
            lo.append(contrastive_loss(sc, temp))
        losses["contra_loss"] = sum(lo) / len(lo) * contra_loss_ratio
    video_input = multimodal_input_video(video_output, sd)
    audio_input = multimodal_input_audio(audio_output, sd)
    if caption_task:
        lo = []
        sel = txt_labels != -1
        for name in ("tva", "tv", "ta"):
            if name not in caption_task:
                continue
            out = bert_forward(txt_input, sd, geom, video_feat=video_input if "v" in name else None,
                               audio_feat=audio_input if "a" in name else None, casual=True)
            out = out[:, : txt_input.shape[1], :][sel]
            scores = mlm_head(out, sd)
            aux[f"caption_scores_{name}"] = scores
            lo.append(F.cross_entropy(scores, txt_labels[sel]))
        losses["caption_loss"] = sum(lo) / len(lo)
    if return_aux:
        return losses, aux
    return losses


# --------------------------------------------------------------------------------------
# retrieval evaluation  (test.py:249-411,680-775)
# --------------------------------------------------------------------------------------


def compute_metric_ret(score_matrix, ids, ids_txt, temp=None, dual_softmax=False, evaluate_ret_text=False):
    """compute_metric_ret, test.py:714-775 (+ compute_dualsoftmax_forward/backward :680-713): full sort per query on the
    score matrix [len(ids_txt), len(ids)], then the position of the ground-truth candidate in the sorted index list."""
    fwd = score_matrix
    if dual_softmax:
        fwd = score_matrix * F.softmax(score_matrix / temp, dim=0) * len(score_matrix)
    idx1 = fwd.sort(dim=-1, descending=True)[1].tolist()
    rank = torch.tensor([idx1[i].index(ids.index(ids_txt[i])) for i in range(len(ids_txt))]).float()

    def summary(r, n, prefix):
        r1, r5, r10 = [(r < k).sum().item() / n for k in (1, 5, 10)]
        return {f"{prefix}_recall": f"{round(r1 * 100, 1)}/{round(r5 * 100, 1)}/{round(r10 * 100, 1)}",
                f"{prefix}_ravg": round((r1 + r5 + r10) / 3 * 100, 1),
                f"{prefix}_medianR": torch.median(r).item() + 1, f"{prefix}_meanR": torch.mean(r).item() + 1}

    log = summary(rank, len(ids_txt), "forward")
    if evaluate_ret_text:
        bwd = score_matrix
        if dual_softmax:
            bwd = score_matrix * F.softmax(score_matrix / temp, dim=1) * len(score_matrix[0])
        idx2 = bwd.sort(dim=0, descending=True)[1].permute(1, 0).tolist()
        r2 = []
        for i in range(len(ids)):
            gts = [j for j, t in enumerate(ids_txt) if t == ids[i]]
            r2.append(min(idx2[i].index(j) for j in gts))
        log.update(summary(torch.tensor(r2).float(), len(ids), "backward"))
    return log


def retrieval_scores(feat_t, feat_v, feat_a, tokens, sd, group):
    """the per-group score matrix of validate_ret (test.py:303-345) for contra_type='fine' with learned fine weights"""
    fb = {"tva": torch.cat((feat_v, feat_a), dim=1) if feat_a is not None and feat_v is not None else None, "tv": feat_v, "ta": feat_a}[group]
    wb = {"tva": lambda: torch.cat((fine_weight(feat_v, sd, "video"), fine_weight(feat_a, sd, "audio")), dim=1),
          "tv": lambda: fine_weight(feat_v, sd, "video"), "ta": lambda: fine_weight(feat_a, sd, "audio")}[group]()
    maskA = (tokens != 0).long()
    maskB = torch.ones(*fb.shape[:2], dtype=torch.long)
    return compute_fine_matrix(feat_t, fb, maskA, maskB, fine_weight(feat_t, sd, "text"), wb)


# --------------------------------------------------------------------------------------
# optimizer step  (optim/adamw.py, optim/sched.py, optim/misc.py, train_utils.py:344-363)
# --------------------------------------------------------------------------------------


def warmup_linear(x, warmup_ratio):
    """optim/sched.py:27-32"""
    if x < warmup_ratio:
        return x / warmup_ratio
    return max((x - 1.0) / (warmup_ratio - 1.0), 0)


def is_no_decay(name):
    """optim/misc.py:14 — substring match on 'bias', 'LayerNorm.bias', 'LayerNorm.weight'."""
    return any(nd in name for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))


def clip_grad_norm_(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ as called at train_utils.py:359 (2-norm, eps 1e-6)."""
    total = torch.sqrt(sum((g.detach().float() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.01):
    """optim/adamw.py:50-101 (bias-corrected; decoupled decay applied AFTER the Adam update
    with the un-corrected lr)."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
