"""TEST-ONLY stand-in for `valor_b200.kernels`: the same tensor-level signatures implemented with
plain torch CPU ops.  It exists so the HOST LOGIC of the product (module wiring, state-dict
layout, hand-written backward compositions, parameter arenas, masks/ranges, DDP plumbing) can be
verified against the oracle without a GPU.  It is never importable from `valor_b200/` and the
product path has no way to select it: tests install it with `monkeypatch` (see conftest
`cpu_kernels` fixture).  GPU parity tests (`-m gpu`) run the real C-ABI kernels instead.
"""
import math

import torch
import torch.nn.functional as F

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU, ACT_RELU = 0, 1, 2, 3
BACKEND_AUTO, BACKEND_TENSOR, BACKEND_SIMT = 0, 1, 2
launch_count = 0


def _act(x, act):
    if act == ACT_GELU:
        return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if act == ACT_QUICKGELU:
        return x * torch.sigmoid(1.702 * x)
    if act == ACT_RELU:
        return torch.relu(x)
    return x


def _act_grad(x, act):
    if act == ACT_GELU:
        return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    if act == ACT_QUICKGELU:
        s = torch.sigmoid(1.702 * x)
        return s * (1 + 1.702 * x * (1 - s))
    if act == ACT_RELU:
        return (x > 0).to(x.dtype)
    return torch.ones_like(x)


def gemm(a, b, *, a_kmajor=True, b_kmajor=True, bias=None, act=ACT_NONE, residual=None, act_aux=None,
         want_preact=False, out=None, out_dtype=None, accumulate=False, alpha=1.0, backend=0, force_bn=0,
         force_splits=0, bias_grad=None, row_scale=None, rows_per_group=1):
    A = a.float() if a_kmajor else a.float().t()
    if bias_grad is not None:
        bias_grad += alpha * A.sum(1)
    Bm = b.float() if b_kmajor else b.float().t()
    x = alpha * (A @ Bm.t())
    if bias is not None:
        x = x + bias
    if row_scale is not None:
        x = x * row_scale.repeat_interleave(rows_per_group)[:x.shape[0], None]
    pre = x
    if act_aux is not None:
        x = x * _act_grad(act_aux.float(), act)
    else:
        x = _act(x, act)
    if residual is not None:
        x = x + residual.float()
    dt = out.dtype if out is not None else (out_dtype or a.dtype)
    if out is None:
        out = torch.zeros(x.shape, dtype=dt)
        accumulate = False
    if accumulate:
        out += x.to(dt)
    else:
        out.copy_(x.to(dt))
    return (out, pre.to(dt)) if want_preact else out


def layernorm_fwd(x, gamma, beta, eps):
    xf = x.float()
    mean = xf.mean(-1)
    var = ((xf - mean[:, None]) ** 2).mean(-1)
    rstd = torch.rsqrt(var + eps)
    y = (xf - mean[:, None]) * rstd[:, None] * gamma + beta
    return y.to(x.dtype), mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None):
    dyf, xf = dy.float(), x.float()
    xh = (xf - mean[:, None]) * rstd[:, None]
    if dgamma is not None:
        dgamma += (dyf * xh).sum(0)
    if dbeta is not None:
        dbeta += dyf.sum(0)
    g = dyf * gamma
    dx = (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True)) * rstd[:, None]
    if dres is not None:
        dx = dx + dres.float()
    return dx.to(x.dtype)


def l2norm_fwd(x):
    n = x.float().norm(dim=-1).clamp_min(1e-12)
    return (x.float() / n[:, None]).to(x.dtype), n


def l2norm_bwd(dy, x, nrm):
    y = x.float() / nrm[:, None]
    return ((dy.float() - y * (y * dy.float()).sum(-1, keepdim=True)) / nrm[:, None]).to(x.dtype)


def _mha_problem(p, Nq, max_nk, q_row0, kv_row0, kv_len):
    q0 = int(q_row0[p]) if q_row0 is not None else p * Nq
    k0 = int(kv_row0[p]) if kv_row0 is not None else p * max_nk
    nk = int(kv_len[p]) if kv_len is not None else max_nk
    return q0, k0, nk


def _mha_mask(p, Nq, nk, max_nk, key_valid, causal):
    add = torch.zeros(Nq, nk)
    if key_valid is not None:
        add = add + (key_valid[p, :nk] == 0).float()[None, :] * -10000.0
    if causal is not None and int(causal[p]):
        tri = torch.triu(torch.ones(Nq, nk), diagonal=1).bool()
        add = torch.where(tri, torch.full_like(add, -10000.0), add)
    return add


def _range_mask(add, q_key_range, nk):
    """-inf outside each query's visible key range [lo, hi) (keys that do not exist for that query)."""
    if q_key_range is None:
        return add
    j = torch.arange(nk)[None, :]
    r = q_key_range.cpu().long()
    out = (j < r[:, 0:1]) | (j >= r[:, 1:2])
    return (add + torch.zeros(r.shape[0], nk)).masked_fill(out, float("-inf"))


KEEP_OVERRIDE = None   # tests: {(p, h): keep mask [Nq, nk]} extracted from the CUDA kernels


def _attn_keep(drop, p, h, Nq, nk):
    """keep / (1-p) factor of attention-probability dropout for problem p, head h (None: no dropout)"""
    if drop is None:
        return None
    pd, state, site = drop
    if KEEP_OVERRIDE is not None:
        return KEEP_OVERRIDE[(p, h)][:, :nk].float() / (1.0 - pd)
    g = torch.Generator().manual_seed(int(state[0]) * 1000003 + int(state[1]) * 7919 + int(site) * 131 + p * 17 + h)
    return (torch.rand(Nq, nk, generator=g) >= pd).float() / (1.0 - pd)


def mha_fwd(q, k, v, P_, H, hd, Nq, max_nk, scale, q_row0=None, kv_row0=None, kv_len=None, key_valid=None,
            causal=None, q_key_range=None, backend=0, drop=None):
    o = torch.zeros(q.shape[0], H * hd, dtype=q.dtype)
    lse = torch.zeros(P_, H, Nq)
    for p in range(P_):
        q0, k0, nk = _mha_problem(p, Nq, max_nk, q_row0, kv_row0, kv_len)
        add = _range_mask(_mha_mask(p, Nq, nk, max_nk, key_valid, causal), q_key_range, nk)
        for h in range(H):
            qq = q[q0:q0 + Nq, h * hd:(h + 1) * hd].float()
            kk = k[k0:k0 + nk, h * hd:(h + 1) * hd].float()
            vv = v[k0:k0 + nk, h * hd:(h + 1) * hd].float()
            s = qq @ kk.t() * scale + add
            lse[p, h] = torch.logsumexp(s, -1)
            pr = torch.softmax(s, -1)
            keep = _attn_keep(drop, p, h, Nq, nk)
            if keep is not None:
                pr = pr * keep
            o[q0:q0 + Nq, h * hd:(h + 1) * hd] = (pr @ vv).to(q.dtype)
    return o, lse


def mha_bwd(q, k, v, o, do, lse, dq_out, P_, H, hd, Nq, max_nk, scale, q_row0=None, kv_row0=None, kv_len=None,
            key_valid=None, causal=None, dkv_out=None, q_key_range=None, backend=0, drop=None):
    dkv = torch.zeros(k.shape[0], 2 * H * hd)
    for p in range(P_):
        q0, k0, nk = _mha_problem(p, Nq, max_nk, q_row0, kv_row0, kv_len)
        add = _range_mask(_mha_mask(p, Nq, nk, max_nk, key_valid, causal), q_key_range, nk)
        for h in range(H):
            sl = slice(h * hd, (h + 1) * hd)
            qq, kk, vv = q[q0:q0 + Nq, sl].float(), k[k0:k0 + nk, sl].float(), v[k0:k0 + nk, sl].float()
            dd, oo = do[q0:q0 + Nq, sl].float(), o[q0:q0 + Nq, sl].float()
            pr = torch.exp(qq @ kk.t() * scale + add - lse[p, h][:, None])
            dp = dd @ vv.t()
            keep = _attn_keep(drop, p, h, Nq, nk)
            pdrop = pr
            if keep is not None:       # O = (keep * P) V: dP = keep * (dO V^T), dV = (keep * P)^T dO; delta is unchanged
                dp = dp * keep
                pdrop = pr * keep
            ds = pr * (dp - (dd * oo).sum(-1, keepdim=True)) * scale
            dq_out[q0:q0 + Nq, sl] = (ds @ kk).to(dq_out.dtype)
            dkv[k0:k0 + nk, sl] += ds.t() @ qq
            dkv[k0:k0 + nk, H * hd + h * hd:H * hd + (h + 1) * hd] += pdrop.t() @ dd
    if dkv_out is not None:
        dkv_out[0].copy_(dkv[:, :H * hd].to(dkv_out[0].dtype))
        dkv_out[1].copy_(dkv[:, H * hd:].to(dkv_out[1].dtype))
        return None
    return dkv


def _window_rows(grid, win, shift):
    B, D, H, W = grid
    idx = torch.arange(B * D * H * W).view(B, D, H, W)
    idx = torch.roll(idx, shifts=(-shift[0], -shift[1], -shift[2]), dims=(1, 2, 3))
    idx = idx.view(B, D // win[0], win[0], H // win[1], win[1], W // win[2], win[2])
    return idx.permute(0, 1, 3, 5, 2, 4, 6).reshape(-1, win[0] * win[1] * win[2])


def _window_add(table, grid, win, shift, cfg_win, heads):
    from oracle import valor_oracle as vo
    from tools.synth import relative_position_index
    N = win[0] * win[1] * win[2]
    rpi = relative_position_index(cfg_win)[:N, :N]
    bias = table[rpi.reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)  # [h,N,N]
    mask = None
    if any(shift):
        mask = vo.compute_mask(grid[1], grid[2], grid[3], win, shift)   # [nW,N,N]
    return bias, mask, rpi


def window_attn_fwd(qkv, table, grid, win, shift, cfg_win, heads, hd, scale, backend=0):
    C = heads * hd
    rows = _window_rows(grid, win, shift)
    bias, mask, _ = _window_add(table, grid, win, shift, cfg_win, heads)
    o = torch.zeros(qkv.shape[0], C, dtype=qkv.dtype)
    nW = rows.shape[0] // grid[0]
    lse = torch.zeros(rows.shape[0], heads, rows.shape[1])
    for p in range(rows.shape[0]):
        x = qkv[rows[p]].float()
        for h in range(heads):
            q, k, v = x[:, h * hd:(h + 1) * hd], x[:, C + h * hd:C + (h + 1) * hd], x[:, 2 * C + h * hd:2 * C + (h + 1) * hd]
            s = (q * scale) @ k.t() + bias[h]
            if mask is not None:
                s = s + mask[p % nW]
            lse[p, h] = torch.logsumexp(s, -1)
            o[rows[p], h * hd:(h + 1) * hd] = (torch.softmax(s, -1) @ v).to(qkv.dtype)
    return o, lse


def window_attn_bwd(qkv, o, do, lse, table, dtable, grid, win, shift, cfg_win, heads, hd, scale, backend=0):
    C = heads * hd
    rows = _window_rows(grid, win, shift)
    bias, mask, rpi = _window_add(table, grid, win, shift, cfg_win, heads)
    nW = rows.shape[0] // grid[0]
    dqkv = torch.zeros(qkv.shape, dtype=torch.float32)
    for p in range(rows.shape[0]):
        x = qkv[rows[p]].float()
        dd_all, oo_all = do[rows[p]].float(), o[rows[p]].float()
        for h in range(heads):
            sl = slice(h * hd, (h + 1) * hd)
            q, k, v = x[:, sl], x[:, C + h * hd:C + (h + 1) * hd], x[:, 2 * C + h * hd:2 * C + (h + 1) * hd]
            s = (q * scale) @ k.t() + bias[h]
            if mask is not None:
                s = s + mask[p % nW]
            pr = torch.exp(s - lse[p, h][:, None])
            dd, oo = dd_all[:, sl], oo_all[:, sl]
            ds = pr * (dd @ v.t() - (dd * oo).sum(-1, keepdim=True))
            dtable[:, h].index_add_(0, rpi.reshape(-1), ds.reshape(-1))
            dqkv[rows[p], h * hd:(h + 1) * hd] = ds @ k * scale
            dqkv[rows[p], C + h * hd:C + (h + 1) * hd] = ds.t() @ (q * scale)
            dqkv[rows[p], 2 * C + h * hd:2 * C + (h + 1) * hd] = pr.t() @ dd
    return dqkv.to(qkv.dtype)


def swin_im2col(video, dtype):
    B, Fr, _, Hh, Ww = video.shape
    x = F.pad(video.float().transpose(1, 2), (0, 0, 0, 0, 0, 1))           # [B,3,F+1,H,W]
    x = x.unfold(2, 2, 1).unfold(3, 4, 4).unfold(4, 4, 4)                    # [B,3,F,Ho,Wo,2,4,4]
    x = x.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B * Fr * (Hh // 4) * (Ww // 4), 96)
    return x.to(dtype).contiguous()


def audio_im2col(spec, ps, dtype):
    BA, mel, fr = spec.shape
    x = spec.float().unfold(1, ps, ps).unfold(2, ps, ps)                     # [BA,Pi,Pj,ps,ps]
    return x.reshape(BA * (mel // ps) * (fr // ps), ps * ps).to(dtype).contiguous()


def ast_assemble_fwd(tok, cls, pos, BA, Pn):
    Hd = tok.shape[1]
    x = torch.cat((cls.view(1, 1, Hd).expand(BA, 1, Hd), tok.float().view(BA, Pn, Hd)), dim=1) + pos[None]
    return x.reshape(BA * (Pn + 1), Hd).to(tok.dtype)


def ast_assemble_bwd(dx, dcls, dpos, BA, Pn):
    Hd = dx.shape[1]
    d = dx.float().view(BA, Pn + 1, Hd)
    dcls += d[:, 0].sum(0)
    dpos += d.sum(0)
    return d[:, 1:].reshape(BA * Pn, Hd).to(dx.dtype)


def bert_embed_fwd(tokens, word, pos, type0, dtype):
    R, Tn = tokens.shape
    e = word[tokens] + pos[:Tn][None] + type0
    return e.reshape(R * Tn, -1).to(dtype)


def bert_embed_bwd(de, tokens, dword, dpos, dtype0):
    R, Tn = tokens.shape
    d = de.float().view(R, Tn, -1)
    dword.index_add_(0, tokens.reshape(-1), d.reshape(R * Tn, -1))
    dpos[:Tn] += d.sum(0)
    dtype0 += d.sum((0, 1))


def media_input_fwd(x, frame_emb, type_emb, out, B, nf, X, S_total, row0):
    Hd = x.shape[-1]
    v = x.float().view(B, nf, X, Hd) + frame_emb[:nf][None, :, None, :] + type_emb
    out.view(B, S_total, Hd)[:, row0:row0 + nf * X] = v.reshape(B, nf * X, Hd).to(out.dtype)


def media_input_bwd(dout, dframe, dtype_emb, B, nf, X, S_total, row0):
    Hd = dout.shape[-1]
    d = dout.float().view(B, S_total, Hd)[:, row0:row0 + nf * X].reshape(B, nf, X, Hd)
    dframe[:nf] += d.sum((0, 2))
    dtype_emb += d.sum((0, 1, 2))
    return d.reshape(B * nf * X, Hd).to(dout.dtype)


def patch_merge(src, BD, H, W, C, inverse):
    if not inverse:
        x = src.view(BD, H, W, C)
        y = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
        return y.reshape(-1, 4 * C).contiguous()
    y = src.view(BD, H // 2, W // 2, 4, C)
    x = torch.zeros(BD, H, W, C, dtype=src.dtype)
    x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2] = y[..., 0, :], y[..., 1, :], y[..., 2, :], y[..., 3, :]
    return x.reshape(-1, C)


def mean_pool_fwd(x, R, X):
    return x.float().view(R, X, -1).mean(1).to(x.dtype)


def mean_pool_bwd(dy, R, X):
    return (dy.float()[:, None, :].expand(R, X, dy.shape[-1]) / X).reshape(R * X, -1).to(dy.dtype)


def colsum(dy, db):
    db += dy.float().sum(0)


def _keep_mask(shape, p, rng_state, site):
    g = torch.Generator().manual_seed(int(rng_state[0]) * 1000003 + int(rng_state[1]) * 7919 + int(site))
    return (torch.rand(shape, generator=g) >= p).float()


def dropout(x, residual, p, rng_state, site, out=None):
    y = x.float() * _keep_mask(x.shape, p, rng_state, site) / (1.0 - p)
    if residual is not None:
        y = y + residual.float()
    y = y.to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def droppath_scale(B, p, rng_state, site):
    g = torch.Generator().manual_seed(int(rng_state[0]) * 1000003 + int(rng_state[1]) * 7919 + int(site))
    return torch.floor(1.0 - p + torch.rand(B, generator=g)) / (1.0 - p)


def row_scale(x, scale, rows_per_group, residual=None):
    y = x.float() * scale.repeat_interleave(rows_per_group)[:, None]
    if residual is not None:
        y = y + residual.float()
    return y.to(x.dtype)


def retrieval_rank(S, gt, by_column=False):
    M = S.t() if by_column else S
    ref = M.gather(1, gt.long()[:, None])
    return (M > ref).sum(1).int()


def dual_softmax(S, temp, dim):
    return S * torch.softmax(S / temp, dim=dim) * S.shape[dim]


def split_bf16x3(x, side):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat((hi, hi, lo) if side == 0 else (hi, lo, hi), dim=1)


def cast2d(src, dst):
    dst.copy_(src.to(dst.dtype))


def cast_flat(src, dst):
    dst.copy_(src.to(dst.dtype))


def act_bwd(dy, h, act):
    return (dy.float() * _act_grad(h.float(), act)).to(h.dtype)


def strided_rows(src, dst, accumulate=False):
    if accumulate:
        dst += src
    else:
        dst.copy_(src)


def xent_fwd(logits, labels):
    lf = logits.float()
    valid = labels >= 0
    lse = torch.logsumexp(lf, -1) * valid
    picked = lf.gather(1, labels.clamp_min(0)[:, None])[:, 0]
    acc = torch.stack(((lse - picked)[valid].sum(), valid.sum().float()))
    return (acc[0] / acc[1].clamp_min(1)).view(1), lse, acc


def xent_bwd(logits, labels, lse, acc, g, gmul=1.0):
    lf = logits.float()
    valid = labels >= 0
    p = torch.exp(lf - lse[:, None])
    p[torch.arange(len(labels))[valid], labels[valid]] -= 1.0
    p = p * valid[:, None] * (g.view(()) * gmul / acc[1].clamp_min(1))
    return p.to(logits.dtype)  # (the CUDA kernel overwrites `logits` in place; same values)


def masked_softmax_fwd(w, mask):
    wf = w.float()
    if mask is not None:
        wf = wf.masked_fill(mask == 0, float("-inf"))
    return torch.softmax(wf, -1)


def masked_softmax_bwd(ws, dws):
    return ws * (dws - (ws * dws).sum(-1, keepdim=True))


def _fine_logits(L, mA, Na, Nb, T, Vt, v0, nv):
    x = L.view(Na, T, Nb, Vt)[:, :, :, v0:v0 + nv].permute(0, 2, 1, 3)        # [a,b,t,v]
    if mA is not None:
        x = x * mA.float()[:, None, :, None]
    return x


def fine_reduce_fwd(L, mA, wsA, wsB, Na, Nb, T, Vt, v0, nv):
    x = _fine_logits(L, mA, Na, Nb, T, Vt, v0, nv)
    a2b, av = x.max(-1)
    b2a, at = x.max(-2)
    score = 0.5 * ((a2b * wsA[:, None, :]).sum(-1) + (b2a * wsB[None]).sum(-1))
    return score, av.to(torch.uint8), at.to(torch.uint8)


def fine_reduce_bwd(L, mA, wsA, wsB, dscore, arg_v, arg_t, dL, dwsA, dwsB, Na, Nb, T, Vt, v0, nv):
    x = _fine_logits(L, mA, Na, Nb, T, Vt, v0, nv)
    g = 0.5 * dscore
    a2b = x.gather(-1, arg_v.long()[..., None])[..., 0]                        # [a,b,t]
    b2a = x.gather(-2, arg_t.long()[:, :, None, :])[:, :, 0, :]                # [a,b,v]
    dwsA += (g[:, :, None] * a2b).sum(1)
    dwsB += (g[:, :, None] * b2a).sum(0)
    dx = torch.zeros_like(x)
    dx.scatter_add_(-1, arg_v.long()[..., None], (g[:, :, None] * wsA[:, None, :])[..., None])
    dx.scatter_add_(-2, arg_t.long()[:, :, None, :], (g[:, :, None] * wsB[None])[:, :, None, :])
    if mA is not None:
        dx = dx * mA.float()[:, None, :, None]
    dL.view(Na, T, Nb, Vt)[:, :, :, v0:v0 + nv] += dx.permute(0, 2, 1, 3)


def contrastive_fwd(S, temp):
    s = S / temp
    rl, cl = torch.logsumexp(s, 1), torch.logsumexp(s, 0)
    d = s.diag()
    return (((rl - d) + (cl - d)).sum() / (2 * S.shape[0])).view(1), rl, cl


def contrastive_bwd(S, temp, row_lse, col_lse, g, dtemp, gmul=1.0):
    N = S.shape[0]
    s = S / temp
    ds = (torch.exp(s - row_lse[:, None]) + torch.exp(s - col_lse[None, :]) - 2 * torch.eye(N)) * (g.view(()) * gmul / (2 * N))
    dtemp += (ds * (-S / (temp * temp))).sum()
    return ds / temp


def grad_sumsq(g, out):
    out += (g.double() ** 2).sum().float()


def clip_coef(sumsq, max_norm, norm_out):
    n = sumsq.sqrt()
    norm_out[0] = n
    norm_out[1] = torch.clamp(max_norm / (n + 1e-6), max=1.0) if max_norm > 0 else 1.0


def adamw(p, g, m, v, p_lp, hyper, coef):
    lr, b1, b2, eps, wd, step_size = [float(x) for x in hyper[:6]]
    gi = g * (float(coef[0]) if coef is not None else 1.0)
    m.mul_(b1).add_(gi, alpha=1 - b1)
    v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
    p.addcdiv_(m, v.sqrt() + eps, value=-step_size)
    if wd > 0:
        p.add_(p, alpha=-lr * wd)
    if p_lp is not None:
        p_lp.copy_(p.to(p_lp.dtype))
