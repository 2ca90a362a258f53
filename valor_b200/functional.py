"""Autograd nodes of the hot path.  Each node is a thin composition of C-ABI kernels
(valor_b200.kernels); the backward formulas are written by hand and the parameter gradients
are accumulated straight into the flat fp32 gradient arena (`param.main_grad`, see
params.ParamStore) by the weight-gradient GEMMs — autograd only orders the nodes.
"""
import torch
from torch.autograd import Function

from . import kernels as K


class Lin:
    """A (possibly fused) linear layer's buffers: low-precision weight view [N,K], fp32
    gradient view, fp32 bias + its gradient."""
    __slots__ = ("w", "w_grad", "b", "b_grad")

    def __init__(self, w, w_grad, b=None, b_grad=None):
        self.w, self.w_grad, self.b, self.b_grad = w, w_grad, b, b_grad


def lin_of(weight, bias=None):
    """Lin for one nn.Parameter pair attached to a ParamStore."""
    return Lin(weight.lp, weight.main_grad, None if bias is None else bias.data,
               None if bias is None else bias.main_grad)


def fused_lin(weights, biases=None):
    """Lin over parameters laid out back-to-back in the arenas (q|k|v packs).  The reference keeps
    separate query/key/value Parameters (bert.py:233-235, transformer.py:109); the arena places
    them adjacently so one GEMM serves all three without renaming any state-dict key."""
    def cat_view(ts):
        t0 = ts[0]
        n = sum(t.shape[0] for t in ts)
        ptr = t0.data_ptr()
        for t in ts:
            assert t.data_ptr() == ptr and t.is_contiguous(), "parameters are not adjacent in the arena"
            ptr += t.numel() * t.element_size()
        return t0.as_strided((n,) + tuple(t0.shape[1:]), t0.stride())

    w = cat_view([p.lp for p in weights])
    wg = cat_view([p.main_grad for p in weights])
    if biases is None:
        return Lin(w, wg)
    return Lin(w, wg, cat_view([p.data for p in biases]), cat_view([p.main_grad for p in biases]))


def _wgrad(lin, dy, x):
    """dW += dy^T x and db += colsum(dy): one launch (the bias gradient rides the weight-gradient GEMM as a ones-column
    MMA on the dy tiles already in shared memory)."""
    if lin.w_grad is not None:
        K.gemm(dy, x, a_kmajor=False, b_kmajor=False, out=lin.w_grad, accumulate=True, bias_grad=lin.b_grad)
    elif lin.b_grad is not None:
        K.colsum(dy, lin.b_grad)


class LinearFn(Function):
    """y = act(x W^T + b) + residual    (nn.Linear + activation + residual add in one GEMM)."""

    @staticmethod
    def forward(ctx, x, lin, act, residual, out_dtype, anchor, rscale=None):
        """rscale = (scale [groups] fp32, rows_per_group): y = residual + scale[row // rpg] * (x W^T + b)  — DropPath
        applied in the epilogue of the GEMM that ends the residual branch (act must be none)."""
        x = x.contiguous()
        N = lin.w.shape[0]
        out = None
        if N % 8 and N > 64 and act == K.ACT_NONE:
            # odd-width outputs (the 30522-wide vocabulary) get a row pitch padded to 8 elements so
            # the 16-byte epilogue stores and the TMA descriptors of the backward GEMMs stay aligned
            out = torch.empty(x.shape[0], (N + 7) // 8 * 8, device=x.device, dtype=out_dtype or x.dtype)[:, :N]
        if act != K.ACT_NONE:
            y, h = K.gemm(x, lin.w, bias=lin.b, act=act, residual=residual, want_preact=True, out_dtype=out_dtype)
        else:
            kw = dict(row_scale=rscale[0], rows_per_group=rscale[1]) if rscale is not None else {}
            y, h = K.gemm(x, lin.w, bias=lin.b, residual=residual, out_dtype=out_dtype, out=out, **kw), None
        assert rscale is None or act == K.ACT_NONE
        ctx.save_for_backward(x, h)
        ctx.lin, ctx.act, ctx.has_res, ctx.rscale = lin, act, residual is not None, rscale
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h = ctx.saved_tensors
        lin = ctx.lin
        if not (dy.dim() == 2 and dy.stride(1) == 1 and dy.stride(0) % 8 == 0):
            dy = dy.contiguous()
        dres = dy if ctx.has_res else None
        if ctx.rscale is not None:   # gradient of the scaled branch
            dy = K.row_scale(dy.contiguous(), ctx.rscale[0], ctx.rscale[1])
        if dy.dtype != x.dtype:  # fp32 head outputs feeding low-precision operands
            t = torch.empty(dy.shape, device=dy.device, dtype=x.dtype)
            K.cast2d(dy, t)
            dy_lp = t
        else:
            dy_lp = dy
        dh = K.act_bwd(dy_lp, h, ctx.act) if h is not None else dy_lp
        _wgrad(lin, dh, x)
        dx = K.gemm(dh, lin.w, b_kmajor=False) if ctx.needs_input_grad[0] else None
        return dx, None, None, dres, None, None, None


def linear(x, lin, act=K.ACT_NONE, residual=None, out_dtype=None, anchor=None, row_scale=None):
    return LinearFn.apply(x, lin, act, residual, out_dtype, anchor, row_scale)


class MlpFn(Function):
    """y = W2 act(W1 x + b1) + b2 + residual — Swin Mlp (videoswin.py:67-73), AST FeedForward
    (transformer.py:141-142), BERT intermediate+output dense (bert.py:403-419).  The activation
    gradient rides the fc2-dgrad GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, lin1, lin2, act, residual, rscale=None):
        x = x.contiguous()
        a, h = K.gemm(x, lin1.w, bias=lin1.b, act=act, want_preact=True)
        kw = dict(row_scale=rscale[0], rows_per_group=rscale[1]) if rscale is not None else {}
        y = K.gemm(a, lin2.w, bias=lin2.b, residual=residual, **kw)
        ctx.save_for_backward(x, h, a)
        ctx.l1, ctx.l2, ctx.act, ctx.has_res, ctx.rscale = lin1, lin2, act, residual is not None, rscale
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h, a = ctx.saved_tensors
        dy = dy.contiguous()
        dres = dy if ctx.has_res else None
        if ctx.rscale is not None:   # gradient of the scaled branch (DropPath inside the fc2 epilogue)
            dy = K.row_scale(dy, ctx.rscale[0], ctx.rscale[1])
        _wgrad(ctx.l2, dy, a)
        dh = K.gemm(dy, ctx.l2.w, b_kmajor=False, act_aux=h, act=ctx.act)
        _wgrad(ctx.l1, dh, x)
        dx = K.gemm(dh, ctx.l1.w, b_kmajor=False) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, dres, None


def mlp(x, lin1, lin2, act, residual=None, row_scale=None):
    return MlpFn.apply(x, lin1, lin2, act, residual, row_scale)


class LN:
    __slots__ = ("g", "b", "g_grad", "b_grad", "eps")

    def __init__(self, weight, bias, eps):
        self.g, self.b, self.g_grad, self.b_grad, self.eps = weight.data, bias.data, weight.main_grad, bias.main_grad, eps


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, ln):
        x = x.contiguous()
        y, mean, rstd = K.layernorm_fwd(x, ln.g, ln.b, ln.eps)
        ctx.save_for_backward(x, mean, rstd)
        ctx.ln = ln
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        ln = ctx.ln
        return K.layernorm_bwd(dy, x, ln.g, mean, rstd, ln.g_grad, ln.b_grad), None


def layer_norm(x, ln):
    return LayerNormFn.apply(x, ln)


class LayerNormResidualFn(Function):
    """(y, x_res) = (LN(x), x): the pre-norm pattern x + f(LN(x)).  `x_res` feeds the residual add of the
    branch's last GEMM; in the backward the gradient of that residual path is added inside the LN-backward
    kernel (one pass) instead of a separate elementwise accumulation of two full [tokens, C] tensors."""

    @staticmethod
    def forward(ctx, x, ln):
        x = x.contiguous()
        y, mean, rstd = K.layernorm_fwd(x, ln.g, ln.b, ln.eps)
        ctx.save_for_backward(x, mean, rstd)
        ctx.ln = ln
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, mean, rstd = ctx.saved_tensors
        ln = ctx.ln
        return K.layernorm_bwd(dy, x, ln.g, mean, rstd, ln.g_grad, ln.b_grad, dres=dres), None


def layer_norm_residual(x, ln):
    return LayerNormResidualFn.apply(x, ln)


class WindowAttnFn(Function):
    """WindowAttention3D core on the natural token order (see include/valor_b200.h)."""

    @staticmethod
    def forward(ctx, qkv, table, dtable, geom):
        grid, win, shift, cfg_win, heads, hd, scale = geom
        o, lse = K.window_attn_fwd(qkv, table, grid, win, shift, cfg_win, heads, hd, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.table, ctx.dtable, ctx.geom = table, dtable, geom
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        grid, win, shift, cfg_win, heads, hd, scale = ctx.geom
        dqkv = K.window_attn_bwd(qkv, o, do, lse, ctx.table, ctx.dtable, grid, win, shift, cfg_win, heads, hd, scale)
        return dqkv, None, None, None


def window_attention(qkv, table_param, geom):
    return WindowAttnFn.apply(qkv, table_param.data, table_param.main_grad, geom)


def _attn_drop(spec):
    """(p, rng_state, site) of this attention call, or None: the probability and the generator ride in the spec dict
    (`attn_drop`, `rng`); the call takes the next site id so that the backward regenerates the same mask."""
    rng, p = spec.get("rng"), spec.get("attn_drop", 0.0)
    if rng is None or not rng.active or p <= 0.0 or spec.get("dtype_is_fp32", False):
        return None
    return (p, rng.state, rng.next_site())


class SelfAttnFn(Function):
    """softmax(q k^T * scale + mask) v over a fused [rows, 3*Hd] q|k|v buffer."""

    @staticmethod
    def forward(ctx, qkv, spec):
        Hd = spec["H"] * spec["hd"]
        ctx.drop = _attn_drop(spec)
        o, lse = K.mha_fwd(qkv[:, :Hd], qkv[:, Hd:2 * Hd], qkv[:, 2 * Hd:], spec["P"], spec["H"], spec["hd"],
                           spec["Nq"], spec["max_nk"], spec["scale"], key_valid=spec.get("key_valid"),
                           causal=spec.get("causal"), drop=ctx.drop)
        ctx.save_for_backward(qkv, o, lse)
        ctx.spec = spec
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        s = ctx.spec
        Hd = s["H"] * s["hd"]
        dqkv = torch.empty_like(qkv)
        K.mha_bwd(qkv[:, :Hd], qkv[:, Hd:2 * Hd], qkv[:, 2 * Hd:], o, do, lse, dqkv[:, :Hd], s["P"], s["H"], s["hd"],
                  s["Nq"], s["max_nk"], s["scale"], key_valid=s.get("key_valid"), causal=s.get("causal"),
                  dkv_out=(dqkv[:, Hd:2 * Hd], dqkv[:, 2 * Hd:]), drop=ctx.drop)
        return dqkv, None


class CrossAttnFn(Function):
    """BertCrossAttention core (bert.py:314-340): q [rows,Hd]; kv [B*S, 2*Hd] (k|v); no mask (bert.py:327).
    One problem per SAMPLE: its queries are the text rows of every caption pass (sample-major batch), its keys the
    sample's media tokens, and `q_key_range` gives each query the key subset of its pass (tva: all, tv: video,
    ta: audio).  Every K/V row then has exactly one owner, so dK/dV are written directly in the compute dtype."""

    @staticmethod
    def forward(ctx, q, kv, spec):
        Hd = spec["H"] * spec["hd"]
        q = q.contiguous()
        ctx.drop = _attn_drop(spec)
        o, lse = K.mha_fwd(q, kv[:, :Hd], kv[:, Hd:], spec["P"], spec["H"], spec["hd"], spec["Nq"], spec["max_nk"],
                           spec["scale"], kv_row0=spec["kv_row0"], kv_len=spec["kv_len"], q_key_range=spec.get("q_key_range"),
                           drop=ctx.drop)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.spec = spec
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        s = ctx.spec
        Hd = s["H"] * s["hd"]
        dq = torch.empty_like(q)
        if s.get("kv_exclusive"):
            dkv = torch.empty_like(kv) if s.get("kv_full") else torch.zeros_like(kv)   # rows outside every range get 0
            K.mha_bwd(q, kv[:, :Hd], kv[:, Hd:], o, do, lse, dq, s["P"], s["H"], s["hd"], s["Nq"], s["max_nk"], s["scale"],
                      kv_row0=s["kv_row0"], kv_len=s["kv_len"], q_key_range=s.get("q_key_range"),
                      dkv_out=(dkv[:, :Hd], dkv[:, Hd:]), drop=ctx.drop)
            return dq, dkv, None
        dkv32 = K.mha_bwd(q, kv[:, :Hd], kv[:, Hd:], o, do, lse, dq, s["P"], s["H"], s["hd"], s["Nq"], s["max_nk"],
                          s["scale"], kv_row0=s["kv_row0"], kv_len=s["kv_len"], q_key_range=s.get("q_key_range"), drop=ctx.drop)
        if kv.dtype == torch.float32:
            dkv = dkv32
        else:
            dkv = torch.empty_like(kv)
            K.cast2d(dkv32, dkv)
        return dq, dkv, None


class BertEmbedFn(Function):
    @staticmethod
    def forward(ctx, anchor, tokens, emb, dtype):
        ctx.tokens, ctx.emb = tokens, emb
        return K.bert_embed_fwd(tokens, emb["word"].data, emb["pos"].data, emb["type"].data[0], dtype)

    @staticmethod
    def backward(ctx, de):
        e = ctx.emb
        K.bert_embed_bwd(de, ctx.tokens, e["word"].main_grad, e["pos"].main_grad, e["type"].main_grad[0])
        return None, None, None, None


class AstAssembleFn(Function):
    @staticmethod
    def forward(ctx, tok, cls, pos, BA, Pn):
        ctx.cls, ctx.pos, ctx.BA, ctx.Pn = cls, pos, BA, Pn
        return K.ast_assemble_fwd(tok.contiguous(), cls.data.view(-1), pos.data, BA, Pn)

    @staticmethod
    def backward(ctx, dx):
        return K.ast_assemble_bwd(dx, ctx.cls.main_grad.view(-1), ctx.pos.main_grad, ctx.BA, ctx.Pn), None, None, None, None


class MediaInputFn(Function):
    """get_multimodal_forward_input_video/audio (modeling.py:485-502) writing the per-sample
    cross-attention source [B, Sv+Sa, Hd] (video rows first, audio rows after: bert.py:450)."""

    @staticmethod
    def forward(ctx, vx, ax, pv, pa, B):
        # vx: [B*nf*X, Hd] or None ; pv = (frame_emb param, type_emb param, nf, X)
        parts = [(vx, pv), (ax, pa)]
        S_total = sum(p[2] * p[3] for x, p in parts if x is not None)
        ref = vx if vx is not None else ax
        out = torch.empty(B * S_total, ref.shape[-1], device=ref.device, dtype=ref.dtype)
        row0 = 0
        rows = []
        for x, p in parts:
            if x is None:
                rows.append(None)
                continue
            K.media_input_fwd(x.contiguous(), p[0].data.view(-1, ref.shape[-1]), p[1].data.view(-1), out, B, p[2], p[3],
                              S_total, row0)
            rows.append(row0)
            row0 += p[2] * p[3]
        ctx.parts, ctx.rows, ctx.B, ctx.S_total = (pv, pa), rows, B, S_total
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        grads = []
        for p, row0 in zip(ctx.parts, ctx.rows):
            if row0 is None:
                grads.append(None)
                continue
            Hd = dout.shape[-1]
            grads.append(K.media_input_bwd(dout, p[0].main_grad.view(-1, Hd), p[1].main_grad.view(-1), ctx.B, p[2], p[3],
                                           ctx.S_total, row0))
        return grads[0], grads[1], None, None, None


class PatchMergeFn(Function):
    @staticmethod
    def forward(ctx, x, BD, H, W, C):
        ctx.dims = (BD, H, W, C)
        return K.patch_merge(x, BD, H, W, C, False)

    @staticmethod
    def backward(ctx, dy):
        return K.patch_merge(dy, *ctx.dims, True), None, None, None, None


class MeanPoolFn(Function):
    @staticmethod
    def forward(ctx, x, R, X):
        ctx.dims = (R, X)
        return K.mean_pool_fwd(x, R, X)

    @staticmethod
    def backward(ctx, dy):
        return K.mean_pool_bwd(dy, *ctx.dims), None, None


class SelectFirstFn(Function):
    """x [R, X, C] -> x[:, 0, :]  (cls token per clip, modeling.py:399)."""

    @staticmethod
    def forward(ctx, x, R, X):
        C = x.shape[-1]
        x = x.contiguous()
        y = torch.empty(R, C, device=x.device, dtype=x.dtype)
        K.strided_rows(x.view(R, X * C)[:, :C], y)
        ctx.dims = (R, X, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        R, X, C = ctx.dims
        dx = torch.zeros(R * X, C, device=dy.device, dtype=dy.dtype)
        K.strided_rows(dy.contiguous(), dx.view(R, X * C)[:, :C])
        return dx, None, None


class L2NormFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y, nrm = K.l2norm_fwd(x)
        ctx.save_for_backward(x, nrm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, nrm = ctx.saved_tensors
        return K.l2norm_bwd(dy, x, nrm)


class RngState:
    """{seed, step offset} in device memory + the host-side call-site counter of the current step.  Every stochastic
    layer takes the next site id in forward order and keeps it for its backward; `begin_step` advances the device
    offset (16-byte copy), so a captured CUDA graph replays with new masks."""
    STRIDE = 1 << 24

    def __init__(self, device, seed=0):
        self.state = torch.tensor([int(seed), 0], dtype=torch.int64, device=device)
        self._host = torch.tensor([int(seed), 0], dtype=torch.int64)
        self.site = 0
        self.active = True

    def begin_step(self):
        """start of a forward pass: restart the call-site numbering and move to a fresh counter range.  While a CUDA
        graph is being captured only the numbering restarts: the replaying loop calls `advance()` itself."""
        self.site = 0
        if not (self.state.is_cuda and torch.cuda.is_current_stream_capturing()):
            self.advance()

    def advance(self):
        self._host[1] += self.STRIDE
        self.state.copy_(self._host, non_blocking=True)

    def next_site(self):
        self.site += 1
        return self.site


class DropoutAddFn(Function):
    """out = residual + dropout(x, p): nn.Dropout before the residual add of BERT / AST sub-layers
    (bert.py:353-355,367-371,418-420; transformer.py:78,83) and on embeddings (residual None)."""

    @staticmethod
    def forward(ctx, x, residual, p, rng):
        ctx.p, ctx.rng, ctx.site, ctx.has_res = p, rng, rng.next_site(), residual is not None
        return K.dropout(x.contiguous(), None if residual is None else residual.contiguous(), p, rng.state, ctx.site)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return K.dropout(dy, None, ctx.p, ctx.rng.state, ctx.site), (dy if ctx.has_res else None), None, None


def residual_branch(y_fn, residual, rng, p, n_groups=None):
    """x + regularise(branch): `y_fn(res[, row_scale])` runs the branch's last GEMM with `res` fused into its epilogue.
    Without an active generator (eval / parity mode, or p == 0) only the residual add is fused; DropPath (n_groups =
    batch) adds its per-sample factor to the same epilogue; Dropout (n_groups None) runs the GEMM bare and applies the
    element mask together with the add."""
    if rng is None or not rng.active or p <= 0.0:
        return y_fn(residual)
    if n_groups is None:
        return DropoutAddFn.apply(y_fn(None), residual, p, rng)
    # DropPath: the per-sample keep / keep_prob factor rides the branch's last GEMM epilogue together with the residual
    # add (one pass over the activations instead of GEMM + scale-and-add); the backward scales the branch gradient
    scale = K.droppath_scale(n_groups, p, rng.state, rng.next_site())
    return y_fn(residual, (scale, residual.shape[0] // n_groups))


class CastFn(Function):
    """dtype change between the fp32 contrastive head and the low-precision GEMM operands"""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        if x.dtype == dtype:
            return x.view_as(x)
        y = torch.empty(x.shape, device=x.device, dtype=dtype)
        K.cast2d(x.contiguous(), y)
        return y

    @staticmethod
    def backward(ctx, dy):
        if dy.dtype == ctx.src:
            return dy, None
        dx = torch.empty(dy.shape, device=dy.device, dtype=ctx.src)
        K.cast2d(dy.contiguous(), dx)
        return dx, None


class FineSimFn(Function):
    """compute_fine_matrix for the three modality groups at once (pretrain.py:178-211,302-345).
    split_operands: features are fp32 and the similarity GEMM runs on the bf16 tensor cores through the two-term
    expansion of both operands (valor_split_bf16x3), i.e. with fp32-grade dot products.
    feat_t [Na*T, D], feat_va [Nb*Vt, D] (per sample: video slots then audio slots),
    w_t [Na,T], w_v [Nb,nV], w_a [Nb,nA] raw fine weights (fp32), maskA [Na,T] uint8.
    Returns scores [G, Na, Nb] for groups tva / tv / ta (those requested)."""

    @staticmethod
    def forward(ctx, feat_t, feat_va, w_t, w_v, w_a, maskA, dims, groups, split_operands=False):
        Na, Nb, T, nV, nA = dims
        Vt = nV + nA
        ctx_lp = None
        if split_operands:
            # perf mode: fp32 features, bf16 tensor cores: two-term expansion of both operands, ONE GEMM over K = 3D
            ta, vb = K.split_bf16x3(feat_t.contiguous(), 0), K.split_bf16x3(feat_va.contiguous(), 1)
            L = K.gemm(ta, vb, out_dtype=torch.float32)                                # [Na*T, Nb*Vt]
            ctx_lp = (ta, vb)
        else:
            L = K.gemm(feat_t.contiguous(), feat_va.contiguous(), out_dtype=torch.float32)
        ctx.lp = ctx_lp
        wsA = K.masked_softmax_fwd(w_t, maskA)
        scores, saved = [], []
        for g in groups:
            v0, nv = {"tva": (0, Vt), "tv": (0, nV), "ta": (nV, nA)}[g]
            wB = {"tva": torch.cat((w_v, w_a), dim=1), "tv": w_v, "ta": w_a}[g].contiguous()
            wsB = K.masked_softmax_fwd(wB, None)
            sc, av, at = K.fine_reduce_fwd(L, maskA, wsA, wsB, Na, Nb, T, Vt, v0, nv)
            scores.append(sc)
            saved += [wsB, av, at]
        ctx.save_for_backward(feat_t, feat_va, L, wsA, maskA, *saved)
        ctx.dims, ctx.groups = dims, groups
        return torch.stack(scores, 0)

    @staticmethod
    def backward(ctx, dscores):
        feat_t, feat_va, L, wsA, maskA, *saved = ctx.saved_tensors
        Na, Nb, T, nV, nA = ctx.dims
        Vt = nV + nA
        dL = torch.zeros_like(L)
        dwsA = torch.zeros_like(wsA)
        dw_v = torch.zeros(Nb, nV, device=L.device, dtype=torch.float32)
        dw_a = torch.zeros(Nb, nA, device=L.device, dtype=torch.float32)
        for i, g in enumerate(ctx.groups):
            v0, nv = {"tva": (0, Vt), "tv": (0, nV), "ta": (nV, nA)}[g]
            wsB, av, at = saved[3 * i: 3 * i + 3]
            dwsB = torch.zeros_like(wsB)
            K.fine_reduce_bwd(L, maskA, wsA, wsB, dscores[i], av, at, dL, dwsA, dwsB, Na, Nb, T, Vt, v0, nv)
            dwB = K.masked_softmax_bwd(wsB, dwsB)
            if g == "tva":
                dw_v += dwB[:, :nV]
                dw_a += dwB[:, nV:]
            elif g == "tv":
                dw_v += dwB
            else:
                dw_a += dwB
        dw_t = K.masked_softmax_bwd(wsA, dwsA)
        if ctx.lp is not None:      # gradients through the leading (hi) terms of the expansions, fp32 out
            ta, vb = ctx.lp
            D = feat_t.shape[1]
            dLl = torch.empty(dL.shape, device=dL.device, dtype=ta.dtype)
            K.cast2d(dL, dLl)
            dft = K.gemm(dLl, vb[:, :D], b_kmajor=False, out_dtype=torch.float32)
            dfva = K.gemm(dLl, ta[:, :D], a_kmajor=False, b_kmajor=False, out_dtype=torch.float32)
            return dft, dfva, dw_t, dw_v, dw_a, None, None, None, None
        if feat_t.dtype != torch.float32:
            dLl = torch.empty(dL.shape, device=dL.device, dtype=feat_t.dtype)
            K.cast2d(dL, dLl)
        else:
            dLl = dL
        dft = K.gemm(dLl, feat_va, b_kmajor=False)
        dfva = K.gemm(dLl, feat_t, a_kmajor=False, b_kmajor=False)
        return dft, dfva, dw_t, dw_v, dw_a, None, None, None, None


class ContrastiveFn(Function):
    """VALORModel.contrastive_loss (modeling.py:418-433) with the learnable temperature."""

    @staticmethod
    def forward(ctx, S, temp):
        S = S.contiguous()
        loss, rl, cl = K.contrastive_fwd(S, temp.data.view(1))
        ctx.save_for_backward(S, rl, cl)
        ctx.temp = temp
        return loss

    @staticmethod
    def backward(ctx, g):
        S, rl, cl = ctx.saved_tensors
        g = g.contiguous().view(1).float()
        return K.contrastive_bwd(S, ctx.temp.data.view(1), rl, cl, g, ctx.temp.main_grad.view(1)), None


class XentFn(Function):
    """F.cross_entropy(scores, labels) with labels == -1 ignored (pretrain.py:441-444)."""

    @staticmethod
    def forward(ctx, logits, labels):
        loss, lse, acc = K.xent_fwd(logits, labels)
        ctx.save_for_backward(logits, labels, lse, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, acc = ctx.saved_tensors
        # the gradient overwrites the saved logits in place (no second 187 MB tensor at the C2 size): a second
        # backward through this node would read gradients as logits, so it is refused instead of being silently wrong
        if getattr(ctx, "consumed", False):
            raise RuntimeError("XentFn.backward ran twice (retain_graph): the logits buffer was already overwritten")
        ctx.consumed = True
        g = g.contiguous().view(1).float()
        return K.xent_bwd(logits, labels, lse, acc, g), None
