"""Tensor-level wrappers over the C ABI (one Python function per entry point of
include/valor_b200.h).  PyTorch is plumbing here: it owns device memory and the stream;
all arithmetic happens inside libvalor_b200.so.  CPU tensors are rejected — there is no
fallback path.
"""
import ctypes

import torch

from . import _lib

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU, ACT_RELU = 0, 1, 2, 3
BACKEND_AUTO, BACKEND_TENSOR, BACKEND_SIMT, BACKEND_MMA_SYNC = 0, 1, 2, 3

launch_count = 0  # kernels-launched-through-the-ABI counter (bench.py's gpu_launches)


def DT(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def P(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("valor_b200 kernels need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def ST():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _call(name, *args):
    global launch_count
    launch_count += 1
    _lib.call(name, *args)


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, f"need a row-major 2-D view, got {tuple(t.shape)} {t.stride()}"
    return t.stride(0)


# ------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------
# tile policy applied when a call does not force one: 0 = the library's choice, 1000 = two-CTA tiles wherever they are
# legal, 2000 = never (measurement / bisecting hook; valor_gemm's force_bn argument)
GEMM_TILE_POLICY = 0


def gemm(a, b, *, a_kmajor=True, b_kmajor=True, bias=None, act=ACT_NONE, residual=None, act_aux=None,
         want_preact=False, out=None, out_dtype=None, accumulate=False, alpha=1.0, backend=BACKEND_AUTO,
         force_bn=0, force_splits=0, bias_grad=None, row_scale=None, rows_per_group=1):
    """C[M,N] (+)= epi(alpha * A . B^T); a: [M,K] (k-major) or [K,M]; b: [N,K] (k-major) or [K,N].
    row_scale [ceil(M / rows_per_group)] fp32: C = residual + row_scale[row // rows_per_group] * (alpha A.B^T + bias)."""
    M, K = (a.shape if a_kmajor else (a.shape[1], a.shape[0]))
    N, Kb = (b.shape if b_kmajor else (b.shape[1], b.shape[0]))
    assert K == Kb, f"gemm: contraction mismatch {K} vs {Kb}"
    assert a.dtype == b.dtype
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype or a.dtype)
    assert out.shape == (M, N)
    preact = torch.empty(M, N, device=a.device, dtype=out.dtype) if want_preact else None
    ep = _lib.ValorGemmEpilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.ldr = _ld(residual) if residual is not None else 0
    ep.res_dtype = DT(residual) if residual is not None else 0
    ep.act_aux = act_aux.data_ptr() if act_aux is not None else None
    ep.ld_aux = _ld(act_aux) if act_aux is not None else 0
    ep.aux_dtype = DT(act_aux) if act_aux is not None else 0
    ep.preact_out = preact.data_ptr() if preact is not None else None
    ep.ld_pre = _ld(preact) if preact is not None else 0
    ep.act = act
    ep.out_dtype = DT(out)
    ep.accumulate = 1 if accumulate else 0
    ep.alpha = alpha
    ep.bias_grad = bias_grad.data_ptr() if bias_grad is not None else None
    if bias_grad is not None:
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() == M and not a_kmajor and not b_kmajor and accumulate
    ep.row_scale = row_scale.data_ptr() if row_scale is not None else None
    ep.rows_per_group = int(rows_per_group)
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous() and row_scale.numel() * rows_per_group >= M
        assert act == ACT_NONE and act_aux is None and not want_preact and not accumulate
    _call("valor_gemm", DT(a), P(a), _ld(a), int(a_kmajor), P(b), _ld(b), int(b_kmajor), P(out), _ld(out), M, N, K,
          ctypes.byref(ep), backend, force_bn or GEMM_TILE_POLICY, force_splits, ST())
    return (out, preact) if want_preact else out


# ------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps):
    M, N = x.shape
    assert x.is_contiguous()
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    _call("valor_layernorm_fwd", DT(x), P(x), P(gamma), P(beta), P(y), P(mean), P(rstd), M, N, float(eps), ST())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None):
    M, N = x.shape
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    if dres is not None:
        dres = dres.contiguous()
    _call("valor_layernorm_bwd", DT(x), P(dy), P(x), P(gamma), P(mean), P(rstd), P(dres), P(dx), P(dgamma), P(dbeta),
          M, N, ST())
    return dx


def l2norm_fwd(x):
    M, N = x.shape
    y = torch.empty_like(x)
    nrm = torch.empty(M, device=x.device, dtype=torch.float32)
    _call("valor_l2norm_fwd", DT(x), P(x), P(y), P(nrm), M, N, ST())
    return y, nrm


def l2norm_bwd(dy, x, nrm):
    M, N = x.shape
    dx = torch.empty_like(x)
    _call("valor_l2norm_bwd", DT(x), P(dy.contiguous()), P(x), P(nrm), P(dx), M, N, ST())
    return dx


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def mha_fwd(q, k, v, P_, H, hd, Nq, max_nk, scale, q_row0=None, kv_row0=None, kv_len=None, key_valid=None,
            causal=None, q_key_range=None, backend=BACKEND_AUTO, drop=None):
    """q: [rows_q, >=H*hd] view, k/v: [rows_kv, >=H*hd] views (may alias one fused buffer).
    q_key_range: int32 [Nq, 2], query i only sees keys lo <= j < hi of its problem.
    drop = (p, rng_state, site): attention-probability dropout (regenerated in mha_bwd from the same triple)."""
    dp, drs, dsite = drop if drop is not None else (0.0, None, 0)
    o = torch.empty(q.shape[0], H * hd, device=q.device, dtype=q.dtype)
    lse = torch.empty(P_, H, Nq, device=q.device, dtype=torch.float32)
    _call("valor_mha_fwd", DT(q), P(q), P(k), P(v), _ld(q), _ld(k), _ld(v), P(o), _ld(o), P(lse), P_, H, hd, Nq,
          max_nk, P(q_row0), P(kv_row0), P(kv_len), P(key_valid), P(causal), P(q_key_range), float(scale), float(dp), P(drs),
          int(dsite), backend, ST())
    return o, lse


def mha_bwd(q, k, v, o, do, lse, dq_out, P_, H, hd, Nq, max_nk, scale, q_row0=None, kv_row0=None, kv_len=None,
            key_valid=None, causal=None, dkv_out=None, q_key_range=None, backend=BACKEND_AUTO, drop=None):
    """dq_out: [rows_q, >=H*hd] view receiving dQ.
    dkv_out=None  -> returns an fp32 [rows_kv, 2*H*hd] buffer with dK|dV accumulated (shared K/V rows: cross-attention);
    dkv_out=(dk_view, dv_view) in the compute dtype -> written directly (each K/V row owned by one problem)."""
    do = do.contiguous()
    dp, drs, dsite = drop if drop is not None else (0.0, None, 0)
    delta = torch.empty_like(lse)
    if dkv_out is None:
        dkv = torch.zeros(k.shape[0], 2 * H * hd, device=q.device, dtype=torch.float32)
        dk, dv = dkv[:, : H * hd], dkv[:, H * hd:]
        args = (P(dk), P(dv), _ld(dk), _ld(dv), None, None, 0)
    else:
        dkv = None
        dk, dv = dkv_out
        if q.dtype == torch.float32:   # parity mode: the row kernel accumulates -> needs zeros
            dk.zero_()
            dv.zero_()
        args = (None, None, 0, 0, P(dk), P(dv), _ld(dk))
    _call("valor_mha_bwd", DT(q), P(q), P(k), P(v), P(o), P(do), _ld(q), _ld(k), _ld(v), _ld(o), P(lse), P(delta), P(dq_out),
          _ld(dq_out), *args, P_, H, hd, Nq, max_nk, P(q_row0), P(kv_row0), P(kv_len), P(key_valid), P(causal),
          P(q_key_range), float(scale), float(dp), P(drs), int(dsite), backend, ST())
    return dkv


def window_attn_fwd(qkv, table, grid, win, shift, cfg_win, heads, hd, scale, backend=BACKEND_AUTO):
    B, D, H, W = grid
    o = torch.empty(qkv.shape[0], heads * hd, device=qkv.device, dtype=qkv.dtype)
    nwin = B * (D // win[0]) * (H // win[1]) * (W // win[2])
    lse = torch.empty(nwin, heads, win[0] * win[1] * win[2], device=qkv.device, dtype=torch.float32)
    _call("valor_window_attn_fwd", DT(qkv), P(qkv), _ld(qkv), P(o), _ld(o), P(lse), P(table), B, D, H, W, *win, *shift,
          *cfg_win, heads, hd, float(scale), backend, ST())
    return o, lse


def window_attn_bwd(qkv, o, do, lse, table, dtable, grid, win, shift, cfg_win, heads, hd, scale,
                    backend=BACKEND_AUTO):
    """returns dqkv [tokens, 3C] in qkv.dtype; dtable (fp32) accumulated in place."""
    B, D, H, W = grid
    dqkv = torch.empty_like(qkv)
    do = do.contiguous()
    nbytes = _lib.load().valor_window_attn_bwd_scratch_bytes(DT(qkv), qkv.shape[0], heads, hd, _ld(qkv), backend)
    scratch = torch.zeros(nbytes // 4, device=qkv.device, dtype=torch.float32) if nbytes else None
    delta = torch.empty_like(lse)
    _call("valor_window_attn_bwd", DT(qkv), P(qkv), _ld(qkv), P(o), P(do), _ld(o), P(lse), P(delta), P(table), P(dqkv),
          _ld(dqkv), P(scratch), P(dtable), B, D, H, W, *win, *shift, *cfg_win, heads, hd, float(scale), backend, ST())
    return dqkv


# ------------------------------------------------------------------------------------------
# data movement
# ------------------------------------------------------------------------------------------
def swin_im2col(video, dtype):
    B, F, C3, Hh, Ww = video.shape
    assert C3 == 3 and video.is_contiguous()
    cols = torch.empty(B * F * (Hh // 4) * (Ww // 4), 96, device=video.device, dtype=dtype)
    _call("valor_swin_im2col", DT(video), DT(cols), P(video), P(cols), B, F, Hh, Ww, ST())
    return cols


def audio_im2col(spec, ps, dtype):
    BA, mel, frames = spec.shape
    assert spec.is_contiguous()
    cols = torch.empty(BA * (mel // ps) * (frames // ps), ps * ps, device=spec.device, dtype=dtype)
    _call("valor_audio_im2col", DT(spec), DT(cols), P(spec), P(cols), BA, mel, frames, ps, ST())
    return cols


def ast_assemble_fwd(tok, cls, pos, BA, Pn):
    Hd = tok.shape[1]
    x = torch.empty(BA * (Pn + 1), Hd, device=tok.device, dtype=tok.dtype)
    _call("valor_ast_assemble_fwd", DT(tok), P(tok), P(cls), P(pos), P(x), BA, Pn, Hd, ST())
    return x


def ast_assemble_bwd(dx, dcls, dpos, BA, Pn):
    Hd = dx.shape[1]
    dx = dx.contiguous()
    dtok = torch.empty(BA * Pn, Hd, device=dx.device, dtype=dx.dtype)
    _call("valor_ast_assemble_bwd", DT(dx), P(dx), P(dtok), P(dcls), P(dpos), BA, Pn, Hd, ST())
    return dtok


def bert_embed_fwd(tokens, word, pos, type0, dtype):
    R, Tn = tokens.shape
    Hd = word.shape[1]
    e = torch.empty(R * Tn, Hd, device=tokens.device, dtype=dtype)
    _call("valor_bert_embed_fwd", DT(e), P(tokens), P(word), P(pos), P(type0), P(e), R, Tn, Hd, ST())
    return e


def bert_embed_bwd(de, tokens, dword, dpos, dtype0):
    R, Tn = tokens.shape
    Hd = de.shape[1]
    de = de.contiguous()
    _call("valor_bert_embed_bwd", DT(de), P(de), P(tokens), P(dword), P(dpos), P(dtype0), R, Tn, Hd, ST())


def media_input_fwd(x, frame_emb, type_emb, out, B, nf, X, S_total, row0):
    Hd = x.shape[-1]
    _call("valor_media_input_fwd", DT(x), P(x), P(frame_emb), P(type_emb), P(out), B, nf, X, Hd, S_total, row0, ST())


def media_input_bwd(dout, dframe, dtype_emb, B, nf, X, S_total, row0):
    Hd = dout.shape[-1]
    din = torch.empty(B * nf * X, Hd, device=dout.device, dtype=dout.dtype)
    _call("valor_media_input_bwd", DT(dout), P(dout), P(din), P(dframe), P(dtype_emb), B, nf, X, Hd, S_total, row0, ST())
    return din


def patch_merge(src, BD, H, W, C, inverse):
    src = src.contiguous()
    if inverse:
        dst = torch.empty(BD * H * W, C, device=src.device, dtype=src.dtype)
    else:
        dst = torch.empty(BD * (H // 2) * (W // 2), 4 * C, device=src.device, dtype=src.dtype)
    _call("valor_patch_merge", DT(src), P(src), P(dst), BD, H, W, C, int(inverse), ST())
    return dst


def mean_pool_fwd(x, R, X):
    C = x.shape[-1]
    y = torch.empty(R, C, device=x.device, dtype=x.dtype)
    _call("valor_mean_pool_fwd", DT(x), P(x.contiguous()), P(y), R, X, C, ST())
    return y


def mean_pool_bwd(dy, R, X):
    C = dy.shape[-1]
    dx = torch.empty(R * X, C, device=dy.device, dtype=dy.dtype)
    _call("valor_mean_pool_bwd", DT(dy), P(dy.contiguous()), P(dx), R, X, C, ST())
    return dx


def colsum(dy, db):
    M, N = dy.shape
    _call("valor_colsum", DT(dy), P(dy), _ld(dy), P(db), M, N, ST())


def cast2d(src, dst):
    """dst[...] = cast(src[...]) for 2-D row-major (possibly strided) views of equal shape."""
    assert src.shape == dst.shape
    R, C = src.shape
    _call("valor_cast2d", DT(src), DT(dst), P(src), _ld(src), P(dst), _ld(dst), R, C, ST())


def cast_flat(src, dst):
    n = src.numel()
    assert dst.numel() == n and src.is_contiguous() and dst.is_contiguous()
    _call("valor_cast2d", DT(src), DT(dst), P(src), n, P(dst), n, 1, n, ST())


def split_bf16x3(x, side):
    """x [R,C] fp32 -> [R,3C] bf16 two-term expansion ([hi|hi|lo] for side 0, [hi|lo|hi] for side 1)."""
    R, C = x.shape
    assert x.dtype == torch.float32
    out = torch.empty(R, 3 * C, device=x.device, dtype=torch.bfloat16)
    _call("valor_split_bf16x3", P(x), _ld(x), P(out), R, C, side, ST())
    return out


def dropout(x, residual, p, rng_state, site, out=None):
    """out = residual + x * keep / (1-p)  (residual None: plain dropout); mask = f(rng_state, site, element index)"""
    R, C = x.shape
    out = torch.empty(R, C, device=x.device, dtype=x.dtype) if out is None else out
    _call("valor_dropout", DT(x), P(x), _ld(x), P(residual), _ld(residual) if residual is not None else 0, P(out), _ld(out), R, C,
          float(p), P(rng_state), int(site), ST())
    return out


def droppath_scale(B, p, rng_state, site):
    scale = torch.empty(B, device=rng_state.device, dtype=torch.float32)
    _call("valor_droppath_scale", P(scale), B, float(p), P(rng_state), int(site), ST())
    return scale


def row_scale(x, scale, rows_per_group, residual=None):
    """out[r] = residual[r] + x[r] * scale[r // rows_per_group]"""
    R, C = x.shape
    out = torch.empty(R, C, device=x.device, dtype=x.dtype)
    _call("valor_row_scale", DT(x), P(x), _ld(x), P(scale), int(rows_per_group), P(residual),
          _ld(residual) if residual is not None else 0, P(out), _ld(out), R, C, ST())
    return out


def act_bwd(dy, h, act):
    dy = dy.contiguous()
    dh = torch.empty_like(h)
    _call("valor_act_bwd", DT(h), P(dy), P(h), P(dh), h.numel(), act, ST())
    return dh


def strided_rows(src, dst, accumulate=False):
    R, C = src.shape
    assert dst.shape == src.shape
    _call("valor_strided_rows", DT(src), P(src), _ld(src), P(dst), _ld(dst), R, C, int(accumulate), ST())


# ------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------
def xent_fwd(logits, labels):
    M, V = logits.shape
    lse = torch.empty(M, device=logits.device, dtype=torch.float32)
    acc = torch.empty(2, device=logits.device, dtype=torch.float32)
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    _call("valor_xent_fwd", DT(logits), P(logits), _ld(logits), P(labels), P(lse), P(acc), P(loss), M, V, ST())
    return loss, lse, acc


def xent_bwd(logits, labels, lse, acc, g, gmul=1.0):
    """in place: logits <- d loss / d logits * g"""
    M, V = logits.shape
    _call("valor_xent_bwd", DT(logits), P(logits), _ld(logits), P(labels), P(lse), P(acc), P(g), float(gmul),
          P(logits), _ld(logits), M, V, ST())
    return logits


def masked_softmax_fwd(w, mask):
    R, L = w.shape
    ws = torch.empty_like(w)
    _call("valor_masked_softmax_fwd", P(w.contiguous()), P(mask), P(ws), R, L, ST())
    return ws


def masked_softmax_bwd(ws, dws):
    R, L = ws.shape
    dw = torch.empty_like(ws)
    _call("valor_masked_softmax_bwd", P(ws), P(dws.contiguous()), P(dw), R, L, ST())
    return dw


def fine_reduce_fwd(L, mA, wsA, wsB, Na, Nb, T, Vt, v0, nv):
    score = torch.empty(Na, Nb, device=L.device, dtype=torch.float32)
    arg_v = torch.empty(Na, Nb, T, device=L.device, dtype=torch.uint8)
    arg_t = torch.empty(Na, Nb, nv, device=L.device, dtype=torch.uint8)
    _call("valor_fine_reduce_fwd", P(L), _ld(L), P(mA), P(wsA), P(wsB), P(score), P(arg_v), P(arg_t), Na, Nb, T, Vt,
          v0, nv, ST())
    return score, arg_v, arg_t


def fine_reduce_bwd(L, mA, wsA, wsB, dscore, arg_v, arg_t, dL, dwsA, dwsB, Na, Nb, T, Vt, v0, nv):
    _call("valor_fine_reduce_bwd", P(L), _ld(L), P(mA), P(wsA), P(wsB), P(dscore.contiguous()), P(arg_v), P(arg_t),
          P(dL), P(dwsA), P(dwsB), Na, Nb, T, Vt, v0, nv, ST())


def contrastive_fwd(S, temp):
    N = S.shape[0]
    row_lse = torch.empty(N, device=S.device, dtype=torch.float32)
    col_lse = torch.empty(N, device=S.device, dtype=torch.float32)
    loss = torch.empty(1, device=S.device, dtype=torch.float32)
    _call("valor_contrastive_fwd", P(S), P(temp), P(row_lse), P(col_lse), P(loss), N, ST())
    return loss, row_lse, col_lse


def contrastive_bwd(S, temp, row_lse, col_lse, g, dtemp, gmul=1.0):
    N = S.shape[0]
    dS = torch.empty_like(S)
    _call("valor_contrastive_bwd", P(S), P(temp), P(row_lse), P(col_lse), P(g), float(gmul), P(dS), P(dtemp), N, ST())
    return dS


def retrieval_rank(S, gt, by_column=False):
    """rank[i] = #{candidates scoring above query i's ground truth}.  S [Nt, Nv] fp32 row-major; by_column: queries are
    the columns (video -> text direction)."""
    Nt, Nv = S.shape
    Nq, Nc = (Nv, Nt) if by_column else (Nt, Nv)
    sq, sc = (1, S.stride(0)) if by_column else (S.stride(0), 1)
    rank = torch.empty(Nq, device=S.device, dtype=torch.int32)
    _call("valor_retrieval_rank", P(S), sq, sc, P(gt), P(rank), Nq, Nc, ST())
    return rank


def dual_softmax(S, temp, dim):
    """S * softmax(S / temp, dim) * S.shape[dim]  (test.py:685-713)"""
    Nt, Nv = S.shape
    out = torch.empty_like(S)
    if dim == 0:
        _call("valor_dual_softmax", P(S), P(out), S.stride(0), 1, P(temp), Nt, Nv, ST())
    else:
        _call("valor_dual_softmax", P(S), P(out), 1, S.stride(0), P(temp), Nv, Nt, ST())
    return out


# ------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------
def grad_sumsq(g, out):
    _call("valor_grad_sumsq", P(g), g.numel(), P(out), ST())


def clip_coef(sumsq, max_norm, norm_out):
    _call("valor_clip_coef", P(sumsq), float(max_norm), P(norm_out), ST())


def adamw(p, g, m, v, p_lp, hyper, coef):
    _call("valor_adamw", P(p), P(g), P(m), P(v), P(p_lp), p.numel(), P(hyper), P(coef), ST())
